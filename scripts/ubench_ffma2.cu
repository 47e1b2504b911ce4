// microbenchmark: FFMA vs FFMA2 (fma.rn.f32x2) throughput on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
template <int MODE>
__global__ void k(float* out, int iters, float s, float t) {
    float acc = 0;
    if (MODE == 0) {
        float a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = threadIdx.x + j;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = fmaf(a[j], s, t);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) acc += a[j];
    } else {
        u64 a[8];
        u64 S = pack(s, s), T = pack(t, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = pack(threadIdx.x + j, threadIdx.x - j);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fma2(a[j], S, T);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { float p, q; unpack(a[j], p, q); acc += p + q; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f); else k<1><<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma = 148.0 * 8 * 256 * iters * 16;
        printf("%s: %.3f ms  %.2f TFMA/s (%.1f FMA/clk/SM at 1.965 GHz)\n", mode ? "FFMA2" : "FFMA ", ms, fma / ms * 1e-9, fma / (ms * 1e-3) / 148 / 1.965e9);
    }
    return 0;
}
