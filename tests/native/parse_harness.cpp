// CPU harness for learningorchestra_b200/csrc/parse_number.cuh (the header is __host__ __device__; here it is
// compiled with g++ so the scanner / Eisel-Lemire / big-integer path can be checked against Python's float()).
#include <stdint.h>
#include "parse_number.cuh"

extern "C" void parse_batch(const uint8_t *chars, const int64_t *offsets, int64_t n, uint64_t *bits, uint8_t *status) {
    for (int64_t i = 0; i < n; ++i)
        status[i] = lo::num::parse_number(chars + offsets[i], (int)(offsets[i + 1] - offsets[i]), bits[i]);
}
