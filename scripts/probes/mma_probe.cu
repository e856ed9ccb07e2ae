// Throughput probe (diagnostic, not product): legacy mma.sync variants on sm_100a.
// Question it answers: can a one-hot nibble outer product (256 MAC per input byte)
// on the legacy tensor path outrun one shared-memory atomic per byte for the u8 column histogram?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int KIND>
__global__ void __launch_bounds__(256) k_mma(int iters, int *sink) {
    unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    int c[4][4] = {};
    float f[4][4] = {};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+r"(c[u][0]), "+r"(c[u][1]), "+r"(c[u][2]), "+r"(c[u][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+f"(f[u][0]), "+f"(f[u][1]), "+f"(f[u][2]), "+f"(f[u][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                    : "+r"(c[u][0]), "+r"(c[u][1]), "+r"(c[u][2]), "+r"(c[u][3]) : "r"(a0), "r"(a1), "r"(b0));
        }
    }
    int s = 0; float t = 0;
    for (int u = 0; u < 4; ++u) for (int q = 0; q < 4; ++q) { s += c[u][q]; t += f[u][q]; }
    sink[threadIdx.x + blockIdx.x * blockDim.x] = s + (int)t;
}
template <int KIND>
void run(const char *name, double macs_per_mma, int *sink) {
    const int iters = 4096, grid = 148 * 8;
    k_mma<KIND><<<grid, 256>>>(16, sink);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_mma<KIND><<<grid, 256>>>(iters, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)grid * 8 * iters * 4;
    printf("{\"probe\":\"%s\",\"ms\":%.3f,\"warp_mma_per_s\":%.4g,\"Tmac_per_s\":%.2f,\"one_hot_hist_TBs_at_256mac_per_byte\":%.3f}\n",
           name, ms, mmas / ms * 1e3, mmas * macs_per_mma / ms * 1e3 / 1e12, mmas * macs_per_mma / ms * 1e3 / 256 / 1e12);
}
int main() {
    int *sink; cudaMalloc(&sink, 148 * 8 * 256 * 4);
    run<0>("imma_m16n8k32_u8", 16.0 * 8 * 32, sink);
    run<1>("hmma_m16n8k16_bf16", 16.0 * 8 * 16, sink);
    run<2>("imma_m16n8k16_u8", 16.0 * 8 * 16, sink);
    return 0;
}
