// parse_number.cuh — decimal text -> binary64 exactly as CPython's float(str), plus the reference's
// "is it an integer" test: the arithmetic of the REAL cast of the /fieldTypes service,
//     values[field] = float(document[field]); if values[field].is_integer(): values[field] = int(...)
// (data_type_handler_image/data_type_update.py:40-43), which the reference runs one document at a time.
//
// Everything here is `__host__ __device__` and pure integer arithmetic so the very same code is unit-tested
// on the CPU against Python's own float() (tests/test_parse_cpu.py compiles it with g++) before it runs in
// the k_parse_number kernel.
//
// Grammar (CPython Objects/floatobject.c float_new -> _Py_string_to_number_with_underscores ->
// PyOS_string_to_double): ASCII whitespace stripped at both ends; '_' only between two digits; then
//     [+-] ( digits [ '.' digits* ] | '.' digits+ ) [ (e|E) [+-] digits+ ]   |   [+-] (inf | infinity | nan)
// Value: correctly rounded (nearest-even) binary64, overflow -> +-inf, underflow -> subnormals / +-0.
//
// Algorithm: Eisel-Lemire on the first 19 significant digits (exact for <= 19 digits: Mushtak & Lemire,
// "Fast number parsing without fallback", 2023).  With more digits the truncated significand w and w+1 are
// both converted; if they disagree the decision is made exactly by comparing S*10^q against the midpoint
// (2m+1)*2^(e-1) with a small big-integer (all digits up to kMaxDigits, the rest only as a sticky bit).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define LO_HD __host__ __device__ __forceinline__
#define LO_HD_NOINLINE __host__ __device__ __noinline__
#else
#define LO_HD inline
#define LO_HD_NOINLINE inline
#endif

namespace lo {
namespace num {

enum : uint8_t {
    kFloat       = 0,   // finite non-integer, or +-inf / nan                  -> stays a float
    kInteger     = 1,   // finite and integer valued (is_integer())           -> the adapter stores int(v)
    kEmpty       = 2,   // exactly ""                                          -> None (data_type_update.py:37-38)
    kInvalid     = 3,   // float() would raise ValueError
    kUnsupported = 4,   // byte >= 0x80 (the packer did not normalise the text, see below) or cell longer than kMaxLen
};

// CPython's float(str) first maps every non-ASCII character of the str: Unicode whitespace -> ' ', Unicode decimal
// digits (category Nd, e.g. fullwidth or Arabic-Indic) -> '0'..'9', anything else -> invalid
// (_PyUnicode_TransformDecimalAndSpaceToASCII).  That is a property lookup on code points, done by the host packer
// (columnar.ascii_number_text) while it encodes the column; the device sees the normalised ASCII text and does all of
// the arithmetic.  A raw byte >= 0x80 therefore means "not normalised" and is reported, never guessed.
constexpr int kMaxLen    = 1 << 20;   // bytes per cell handled on the device (the algorithm itself has no length limit:
                                      // digits beyond kMaxDigits only contribute a sticky bit)
constexpr int kMaxDigits = 800;    // significant digits kept exactly in the slow path (a midpoint has <= 767)
constexpr int kLimbs     = 168;    // 32-bit limbs of the slow path's big integers (5376 bits)

static
#ifdef __CUDACC__
__device__
#endif
const uint64_t kPow5Device[651 * 2] = {
#include "pow5_table.inc"
};
#ifdef __CUDACC__
static const uint64_t kPow5Host[651 * 2] = {
#include "pow5_table.inc"
};
#endif

LO_HD uint64_t pow5(int idx) {
#if defined(__CUDA_ARCH__)
    return kPow5Device[idx];
#elif defined(__CUDACC__)
    return kPow5Host[idx];
#else
    return kPow5Device[idx];
#endif
}

LO_HD void mul64(uint64_t a, uint64_t b, uint64_t &hi, uint64_t &lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umul64hi(a, b);
#else
    unsigned __int128 p = (unsigned __int128)a * b;
    lo = (uint64_t)p;
    hi = (uint64_t)(p >> 64);
#endif
}

LO_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

struct Binary {          // value = mantissa (52 explicit bits, hidden bit removed) , biased exponent power2
    uint64_t mantissa;
    int32_t  power2;     // 0 = subnormal / zero, 0x7FF = infinity
};

// Eisel-Lemire: nearest binary64 to w * 10^q, w != 0 handled by caller too (w == 0 -> zero)
LO_HD Binary eisel_lemire(int64_t q, uint64_t w) {
    Binary a;
    if (w == 0 || q < -342) { a.mantissa = 0; a.power2 = 0; return a; }
    if (q > 308) { a.mantissa = 0; a.power2 = 0x7FF; return a; }
    const int lz = clz64(w);
    w <<= lz;
    const int idx = 2 * (int)(q + 342);
    uint64_t hi, lo;
    mul64(w, pow5(idx), hi, lo);
    if ((hi & 0x1FFull) == 0x1FFull) {                 // 55 bits of precision wanted: refine with the low word
        uint64_t hi2, lo2;
        mul64(w, pow5(idx + 1), hi2, lo2);
        lo += hi2;
        if (hi2 > lo) hi++;
    }
    const int upperbit = (int)(hi >> 63);
    const int shift = upperbit + 9;
    a.mantissa = hi >> shift;
    a.power2 = (int32_t)((((152170 + 65536) * q) >> 16) + 63 + upperbit - lz + 1023);
    if (a.power2 <= 0) {                                // subnormal
        if (-a.power2 + 1 >= 64) { a.mantissa = 0; a.power2 = 0; return a; }
        a.mantissa >>= -a.power2 + 1;
        a.mantissa += (a.mantissa & 1);
        a.mantissa >>= 1;
        a.power2 = (a.mantissa < (1ull << 52)) ? 0 : 1;
        return a;
    }
    if (lo <= 1 && q >= -4 && q <= 23 && (a.mantissa & 3) == 1) {     // exactly halfway: round to even
        if ((a.mantissa << shift) == hi) a.mantissa &= ~1ull;
    }
    a.mantissa += (a.mantissa & 1);
    a.mantissa >>= 1;
    if (a.mantissa >= (2ull << 52)) { a.mantissa = 1ull << 52; a.power2++; }
    a.mantissa &= ~(1ull << 52);
    if (a.power2 >= 0x7FF) { a.power2 = 0x7FF; a.mantissa = 0; }
    return a;
}

LO_HD uint64_t to_bits(Binary a, bool negative) {
    return ((uint64_t)negative << 63) | ((uint64_t)a.power2 << 52) | a.mantissa;
}

// ---- slow path: exact comparison against the midpoint between two adjacent doubles -----------------
struct Big {
    uint32_t limb[kLimbs];
    int      n;          // used limbs (value 0 <=> n == 0)
};

LO_HD void big_set(Big &b, uint64_t v) {
    b.n = 0;
    if (v & 0xFFFFFFFFull || v >> 32) { b.limb[0] = (uint32_t)v; b.n = 1; }
    if (v >> 32) { b.limb[1] = (uint32_t)(v >> 32); b.n = 2; }
}

LO_HD bool big_mul_add_small(Big &b, uint32_t m, uint32_t add) {      // b = b*m + add ; false on overflow
    uint64_t carry = add;
    for (int i = 0; i < b.n; ++i) {
        uint64_t t = (uint64_t)b.limb[i] * m + carry;
        b.limb[i] = (uint32_t)t;
        carry = t >> 32;
    }
    if (carry) {
        if (b.n >= kLimbs) return false;
        b.limb[b.n++] = (uint32_t)carry;
    }
    return true;
}

LO_HD bool big_mul_pow5(Big &b, int e) {                              // b *= 5^e
    while (e >= 13) { if (!big_mul_add_small(b, 1220703125u, 0)) return false; e -= 13; }   // 5^13 < 2^32
    uint32_t m = 1;
    for (int i = 0; i < e; ++i) m *= 5;
    return e == 0 ? true : big_mul_add_small(b, m, 0);
}

LO_HD bool big_shl(Big &b, int bits) {                                // b <<= bits
    if (b.n == 0 || bits == 0) return true;
    const int words = bits >> 5, rem = bits & 31;
    if (b.n + words + 1 > kLimbs) return false;
    if (rem) {
        uint32_t carry = 0;
        for (int i = 0; i < b.n; ++i) {
            uint32_t v = b.limb[i];
            b.limb[i] = (v << rem) | carry;
            carry = v >> (32 - rem);
        }
        if (carry) b.limb[b.n++] = carry;
    }
    if (words) {
        for (int i = b.n - 1; i >= 0; --i) b.limb[i + words] = b.limb[i];
        for (int i = 0; i < words; ++i) b.limb[i] = 0;
        b.n += words;
    }
    return true;
}

LO_HD int big_cmp(const Big &a, const Big &b) {
    if (a.n != b.n) return a.n < b.n ? -1 : 1;
    for (int i = a.n - 1; i >= 0; --i)
        if (a.limb[i] != b.limb[i]) return a.limb[i] < b.limb[i] ? -1 : 1;
    return 0;
}

// ---- the scanner ------------------------------------------------------------------------------------
LO_HD bool is_space(uint8_t c) { return (c >= 0x09 && c <= 0x0D) || c == 0x20; }   // what float() strips in ASCII (0x1C-0x1F are NOT)
LO_HD bool is_digit(uint8_t c) { return (uint8_t)(c - '0') <= 9; }
LO_HD uint8_t lower(uint8_t c) { return (c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c; }

// decides between `down` (from w) and the next double up, exactly.  digits are re-read from the text.
// returns false if the big integers overflowed (cannot happen within kMaxLen / kMaxDigits; kept as a guard).
LO_HD_NOINLINE bool slow_path_round_up(const uint8_t *s, int b, int e, int64_t exp10_explicit, Binary down, bool &round_up) {
    // S = all significant digits (leading zeros skipped), up to kMaxDigits; q10 = decimal exponent of S's last digit
    Big S;
    S.n = 0;
    int nd = 0;
    bool sticky = false, seen_point = false, leading = true;
    int64_t frac_digits_kept = 0;      // digits after the point that went into S (or were skipped as leading zeros)
    int64_t int_digits_dropped = 0;    // digits before the point that were NOT put into S
    for (int i = b; i < e; ++i) {
        const uint8_t c = s[i];
        if (c == '_') continue;
        if (c == '.') { seen_point = true; continue; }
        if (!is_digit(c)) break;       // exponent marker
        const uint32_t d = c - '0';
        if (leading && d == 0) { if (seen_point) frac_digits_kept++; continue; }
        leading = false;
        if (nd < kMaxDigits) {
            if (!big_mul_add_small(S, 10, d)) return false;
            nd++;
            if (seen_point) frac_digits_kept++;
        } else {
            if (d) sticky = true;
            if (!seen_point) int_digits_dropped++;
        }
    }
    const int64_t q10 = exp10_explicit - frac_digits_kept + int_digits_dropped;
    // midpoint h = (2m+1) * 2^(eb-1) with down = m * 2^eb
    uint64_t m;
    int eb;
    if (down.power2 == 0) { m = down.mantissa; eb = -1074; }
    else { m = down.mantissa | (1ull << 52); eb = down.power2 - 1075; }
    Big H;
    big_set(H, 2 * m + 1);
    const int eh = eb - 1;
    // compare S * 10^q10  ?  H * 2^eh   ->   scale both to integers
    if (q10 >= 0) { if (!big_mul_pow5(S, (int)q10) || !big_shl(S, (int)q10)) return false; }
    else          { if (!big_mul_pow5(H, (int)-q10) || !big_shl(H, (int)-q10)) return false; }
    if (eh >= 0) { if (!big_shl(H, eh)) return false; }
    else         { if (!big_shl(S, -eh)) return false; }
    const int c = big_cmp(S, H);
    if (c > 0) round_up = true;
    else if (c < 0) round_up = false;
    else round_up = sticky || (m & 1);        // exact tie: to even, unless nonzero digits were dropped
    return true;
}

// Parses s[0..len).  Writes the IEEE-754 bit pattern of the value and returns the status.
LO_HD uint8_t parse_number(const uint8_t *s, int len, uint64_t &bits_out) {
    bits_out = 0;
    if (len == 0) return kEmpty;
    if (len > kMaxLen) return kUnsupported;
    for (int i = 0; i < len; ++i)
        if (s[i] >= 0x80) return kUnsupported;
    int b = 0, e = len;
    while (b < e && is_space(s[b])) ++b;
    while (e > b && is_space(s[e - 1])) --e;
    if (b == e) return kInvalid;
    // underscores: only between digits (CPython _Py_string_to_number_with_underscores)
    {
        uint8_t prev = 0;
        for (int i = b; i < e; ++i) {
            const uint8_t c = s[i];
            if (c == '_') { if (!is_digit(prev)) return kInvalid; }
            else if (prev == '_' && !is_digit(c)) return kInvalid;
            if (c == 0) return kInvalid;
            prev = c;
        }
        if (prev == '_') return kInvalid;
    }
    int p = b;
    bool negative = false;
    if (s[p] == '-') { negative = true; ++p; }
    else if (s[p] == '+') ++p;
    if (p == e) return kInvalid;
    // inf / infinity / nan
    if (!is_digit(s[p]) && s[p] != '.') {
        const int rem = e - p;
        auto match = [&](const char *w, int n) {
            if (rem != n) return false;
            for (int i = 0; i < n; ++i)
                if (lower(s[p + i]) != (uint8_t)w[i]) return false;
            return true;
        };
        if (match("inf", 3) || match("infinity", 8)) { bits_out = ((uint64_t)negative << 63) | 0x7FF0000000000000ull; return kFloat; }
        if (match("nan", 3)) { bits_out = ((uint64_t)negative << 63) | 0x7FF8000000000000ull; return kFloat; }
        return kInvalid;
    }
    const int digits_begin = p;
    uint64_t w = 0;
    int nsig = 0;                    // significant digits accumulated into w (<= 19)
    int64_t dropped_int = 0;         // integer-part digits not in w
    int64_t frac_in_w = 0;           // fraction digits that are in w or were leading zeros
    bool any_digit = false, too_many = false, seen_point = false;
    for (; p < e; ++p) {
        const uint8_t c = s[p];
        if (c == '_') continue;
        if (c == '.') { if (seen_point) return kInvalid; seen_point = true; continue; }
        if (!is_digit(c)) break;
        any_digit = true;
        const uint32_t d = c - '0';
        if (nsig == 0 && d == 0) { if (seen_point) frac_in_w++; continue; }     // leading zeros
        if (nsig < 19) { w = w * 10 + d; nsig++; if (seen_point) frac_in_w++; }
        else { if (d) too_many = true; if (!seen_point) dropped_int++; }
    }
    if (!any_digit) return kInvalid;
    int64_t exp10 = 0;
    if (p < e) {
        if (s[p] != 'e' && s[p] != 'E') return kInvalid;
        ++p;
        bool eneg = false;
        if (p < e && (s[p] == '-' || s[p] == '+')) { eneg = s[p] == '-'; ++p; }
        if (p == e) return kInvalid;
        bool edig = false;
        for (; p < e; ++p) {
            const uint8_t c = s[p];
            if (c == '_') continue;
            if (!is_digit(c)) return kInvalid;
            edig = true;
            if (exp10 < 100000000) exp10 = exp10 * 10 + (c - '0');
        }
        if (!edig) return kInvalid;
        if (eneg) exp10 = -exp10;
    }
    const int64_t q = exp10 - frac_in_w + dropped_int;
    Binary a = eisel_lemire(q, w);
    if (too_many && w != 0) {
        Binary up = eisel_lemire(q, w + 1);
        if (up.mantissa != a.mantissa || up.power2 != a.power2) {
            // the truncated value straddles a rounding boundary: decide exactly.  `a` may already be the
            // rounded-up neighbour of w's true position, so take the smaller of the two as `down`.
            bool round_up = false;
            if (!slow_path_round_up(s, digits_begin, e, exp10, a, round_up)) return kUnsupported;
            if (round_up) a = up;
        }
    }
    bits_out = to_bits(a, negative);
    if (a.power2 == 0x7FF) return kFloat;                               // inf
    // is_integer(): zero, or exponent large enough that no fraction bits remain
    if (a.power2 == 0) return a.mantissa == 0 ? kInteger : kFloat;       // +-0 -> int 0 ; subnormals are not integers
    const int e2 = a.power2 - 1023;                                     // value = 1.m * 2^e2
    if (e2 < 0) return kFloat;
    if (e2 >= 52) return kInteger;
    return (a.mantissa & ((1ull << (52 - e2)) - 1)) == 0 ? kInteger : kFloat;
}

}  // namespace num
}  // namespace lo
