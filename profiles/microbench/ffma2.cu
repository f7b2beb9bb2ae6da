// Microbenchmark: issue/pipe throughput of scalar FFMA vs packed FFMA2 on B200 (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float2 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    const float2 A = make_float2(a, a), B = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { x[i].x = __fmaf_rn(x[i].x, a, b); x[i].y = __fmaf_rn(x[i].y, a, b); }   // 2 scalar FFMA
            else x[i] = __ffma2_rn(x[i], A, B);                                                       // 1 FFMA2
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 1.0001f, 0.5f); else k<1><<<148 * 8, 256>>>(d, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double fmas = 148.0 * 8 * 256 * iters * 16;   // FMAs (each lane-FMA counted once)
        printf("%s: %.3f ms  %.2f T-FMA/s  (%.1f FMA/clk/SM at 1.9 GHz)\n", mode ? "FFMA2 " : "FFMA  ", ms, fmas / ms / 1e9, fmas / ms / 1e-3 / 148 / 1.9e9);
    }
    return 0;
}
