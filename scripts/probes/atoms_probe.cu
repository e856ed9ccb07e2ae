// Throughput probe (diagnostic, not product): how many conflict-free shared-memory atomics per clock an sm_100a SM
// retires when it does nothing else — the floor under k_hist_u8_cols, which issues one ATOMS.ADD per input byte on a
// lane-private bank.  Same launch shape as the kernel: 256 threads, 66 560 B dynamic smem, 3 CTAs/SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atoms_probe atoms_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// MODE 0: addresses and values precomputed in registers: the loop issues NOTHING but atomics (the pure ATOMS rate).
// MODE 1: a fresh pseudo-random address / value per atomic (LCG + mask + shift: ~6 integer ops per atomic, roughly the
//         arithmetic k_hist_u8_cols needs per byte) — shows what the integer pipes add.
template <int MODE>
__global__ void __launch_bounds__(256, 3) k_atoms(int iters, unsigned long long *cycles, unsigned *sink) {
    extern __shared__ uint32_t smem[];
    for (int w = 0; w < 64; ++w) smem[w * 256 + threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem) + 4 * threadIdx.x;
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x;
    uint32_t addr[8], val[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        x = x * 1664525u + 1013904223u;
        addr[u] = base + ((x >> 8) & 0xFC00u);
        val[u] = 1u << ((x >> 24) & 24u);
    }
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 1) {
                x = x * 1664525u + 1013904223u;
                addr[u] = base + ((x >> 8) & 0xFC00u);
                val[u] = 1u << ((x >> 24) & 24u);
            }
            asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr[u]), "r"(val[u]) : "memory");
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) atomicMax(cycles, (unsigned long long)(t1 - t0));
    unsigned s = 0;
    for (int w = 0; w < 64; ++w) s += smem[w * 256 + threadIdx.x];
    if (s == 0xdeadbeefu) *sink = s;
}

template <int MODE>
void run(const char *name, int sms, unsigned long long *cyc, unsigned *sink) {
    const int smem = (64 * 256 + 256) * 4, iters = 4096, grid = sms * 3;
    cudaFuncSetAttribute(k_atoms<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_atoms<MODE><<<grid, 256, smem>>>(16, cyc, sink);
    cudaMemset(cyc, 0, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_atoms<MODE><<<grid, 256, smem>>>(iters, cyc, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double warp_atoms_per_sm = 3.0 * 8 * iters * 8;          // CTAs x warps x iterations x unroll
    // "cycles" = the longest issue loop of any CTA (clock64): atomics are fire-and-forget, so it undercounts the time the
    // LSU needs to retire them; the wall-time figures (warp_atoms_per_s_per_sm, TBs_chip_from_ms) are the retirement rate.
    printf("{\"probe\":\"%s\",\"ms\":%.3f,\"cycles\":%llu,\"warp_atoms_per_issue_clk_per_sm\":%.4f,"
           "\"warp_atoms_per_s_per_sm\":%.4e,\"TBs_chip_from_ms\":%.3f}\n",
           name, ms, c, warp_atoms_per_sm / (double)c, warp_atoms_per_sm / (ms * 1e-3),
           32.0 * warp_atoms_per_sm * sms / (ms * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    unsigned long long *cyc; unsigned *sink;
    cudaMalloc(&cyc, 8); cudaMalloc(&sink, 4);
    run<0>("atoms_only_precomputed_operands", p.multiProcessorCount, cyc, sink);
    run<1>("atoms_plus_address_arithmetic", p.multiProcessorCount, cyc, sink);
    return 0;
}
