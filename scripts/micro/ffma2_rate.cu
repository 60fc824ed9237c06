// Issue-rate microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f2;
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
template <int MODE>
__global__ void k(float *out, float a, float b, int iters) {
    // 8 independent chains
    float x[8]; f2 y[8];
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x + i; y[i] = (f2)(threadIdx.x + i) * 0x100000001ull; }
    f2 a2 = ((f2)__float_as_uint(a) << 32) | __float_as_uint(a), b2 = ((f2)__float_as_uint(b) << 32) | __float_as_uint(b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) x[i] = fma1(x[i], a, b);
            else y[i] = fma2(y[i], a2, b2);
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i] + (float)(y[i] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 4, 512>>>(out, 1.0001f, 0.5f, iters); else k<1><<<148 * 4, 512>>>(out, 1.0001f, 0.5f, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double winstr = 148.0 * 4 * 16 * 8.0 * iters;  // warp instructions
        printf("%s: %.3f ms, %.2f warp-instr/ns total, %.3f per SM-subpartition per ns (clock ~1.9 GHz -> 1.9 = 1/cycle)\n", mode ? "FFMA2" : "FFMA ", ms, winstr / ms * 1e-6, winstr / ms * 1e-6 / (148 * 4));
    }
    return 0;
}
