// Issue-rate microbenchmark: F2I.U32.TRUNC, I2FP.F32.U32, FADD.RZ, FMNMX, VIMNMX3, IMAD, LOP3 on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cvt_rate cvt_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, float a, unsigned m, int iters) {
    float x[8]; unsigned u[8];
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 0.37f + i; u[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { asm volatile("cvt.rzi.u32.f32 %0, %1;" : "=r"(u[i]) : "f"(x[i])); x[i] = __uint_as_float(u[i] ^ m); }   // F2I + LOP
            if (MODE == 1) { asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(x[i]) : "r"(u[i])); u[i] = __float_as_uint(x[i]) ^ m; }   // I2F + LOP
            if (MODE == 2) { asm volatile("add.rz.f32 %0, %1, %2;" : "=f"(x[i]) : "f"(x[i]), "f"(a)); u[i] = __float_as_uint(x[i]) ^ m; x[i] = __uint_as_float(u[i]); }  // FADD.RZ + LOP
            if (MODE == 3) { asm volatile("xor.b32 %0, %1, %2;" : "=r"(u[i]) : "r"(u[i]), "r"(m)); asm volatile("xor.b32 %0, %1, %2;" : "=r"(u[i]) : "r"(u[i]), "r"(m + 1)); }  // 2 LOP
            if (MODE == 4) { asm volatile("cvt.rmi.f32.f32 %0, %1;" : "=f"(x[i]) : "f"(x[i])); u[i] = __float_as_uint(x[i]) ^ m; x[i] = __uint_as_float(u[i]); }  // FRND.FLOOR + LOP
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i] + (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 10000;
    const char *names[] = {"F2I+LOP", "I2F+LOP", "FADD.RZ+LOP", "LOP+LOP", "FRND+LOP"};
    for (int mode = 0; mode < 5; ++mode) for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        switch (mode) {
        case 0: k<0><<<148 * 4, 512>>>(out, 8388608.f, 3u, iters); break;
        case 1: k<1><<<148 * 4, 512>>>(out, 8388608.f, 3u, iters); break;
        case 2: k<2><<<148 * 4, 512>>>(out, 8388608.f, 3u, iters); break;
        case 3: k<3><<<148 * 4, 512>>>(out, 8388608.f, 3u, iters); break;
        case 4: k<4><<<148 * 4, 512>>>(out, 8388608.f, 3u, iters); break;
        }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double pairs = 148.0 * 4 * 16 * 8.0 * iters;  // warp-level (op + LOP) pairs
        if (rep) printf("%-12s: %.3f ms, %.2f cycles per pair per SM sub-partition at 1.9 GHz\n", names[mode], ms, ms * 1e-3 * 1.9e9 / (pairs / (148 * 4)));
    }
    return 0;
}
