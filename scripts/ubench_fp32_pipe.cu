// Microbenchmark (diagnostic, not product): issue cost of the FP32-pipe instruction forms the trace kernel uses on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_fp32_pipe scripts/ubench_fp32_pipe.cu
// Prints cycles per warp-instruction and SM sub-partition (SMSP) for each form, at 8 warps per SMSP with 8 independent
// accumulator chains per thread (latency hidden), assuming the SM clock given as argv[1] in MHz (default 1965).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;

#define DEV __device__ __forceinline__
DEV u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
DEV float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + 0.f * b; }

enum { FFMA_REUSE, FFMA_3REG, FFMA2_PK, FFMA2_BC, FADD_, FMUL_, FADD2_, FMUL2_BC, MIX_FFMA_FFMA2, MIX_FFMA_ALU, MIX_FFMA2_ALU,
       MIX_FFMA2_FADD, MIX_FFMA_MUFU, NMODES };
static const char* NAMES[] = {"FFMA  a=a*s+t (s,t shared)", "FFMA  a=a*b[j]+c[j] (3 distinct regs)", "FFMA2 a=a*S+T (packed S,T)",
                              "FFMA2 a=s.F32*a+c[j] (scalar-broadcast)", "FADD  a=a+b[j]", "FMUL  a=a*b[j]", "FADD2 a=a+b[j]",
                              "FMUL2 a=a*s.F32", "FFMA + FFMA2 alternating (per pair)", "FFMA + IADD3 alternating (per pair)",
                              "FFMA2 + IADD3 alternating (per pair)", "FFMA2 + FADD alternating (per pair)", "4 FFMA + 1 MUFU.RCP (per group)"};

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s, float t) {
    float a[8], b[8], c[8];
    u64 A[8], B[8], Cc[8];
    int ia[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        a[j] = threadIdx.x + j; b[j] = 0.999f + 1e-5f * (threadIdx.x + j); c[j] = 1e-3f * j + s + 1e-7f * threadIdx.x;
        A[j] = pk(a[j], -a[j]); B[j] = pk(b[j], b[j] * 1.0001f); Cc[j] = pk(c[j], -c[j]);
        ia[j] = threadIdx.x * j;
    }
    const u64 S = pk(s, s), T = pk(t, t);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == FFMA_REUSE) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(s), "f"(t));
            if (MODE == FFMA_3REG) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(b[j]), "f"(c[j]));
            if (MODE == FFMA2_PK) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[j]) : "l"(S), "l"(T));
            if (MODE == FFMA2_BC) { u64 sb; asm volatile("mov.b64 %0, {%1, %1};" : "=l"(sb) : "f"(b[j])); asm volatile("fma.rn.f32x2 %0, %1, %0, %2;" : "+l"(A[j]) : "l"(sb), "l"(Cc[j])); }
            if (MODE == FADD_) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[j]) : "f"(b[j]));
            if (MODE == FMUL_) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[j]) : "f"(b[j]));
            if (MODE == FADD2_) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[j]) : "l"(B[j]));
            if (MODE == FMUL2_BC) { u64 sb; asm volatile("mov.b64 %0, {%1, %1};" : "=l"(sb) : "f"(b[j])); asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(A[j]) : "l"(sb)); }
            if (MODE == MIX_FFMA_FFMA2) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(b[j]), "f"(c[j]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[j]) : "l"(B[j]), "l"(Cc[j]));
            }
            if (MODE == MIX_FFMA_ALU) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(b[j]), "f"(c[j]));
                asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[j]) : "r"(ia[(j + 1) & 7]));
            }
            if (MODE == MIX_FFMA2_ALU) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[j]) : "l"(B[j]), "l"(Cc[j]));
                asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[j]) : "r"(ia[(j + 1) & 7]));
            }
            if (MODE == MIX_FFMA2_FADD) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[j]) : "l"(B[j]), "l"(Cc[j]));
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[j]) : "f"(b[j]));
            }
            if (MODE == MIX_FFMA_MUFU) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(b[j]), "f"(c[j]));
                if ((j & 3) == 3) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(c[j]));
            }
        }
    }
    float acc = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += a[j] + lo(A[j]) + c[j] + (float)ia[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
static void run(float* out, double mhz) {
    const int iters = 4096, blocks = 148 * 4;          // 4 CTAs x 8 warps = 32 warps / SM = 8 / SMSP
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(out, iters, 0.999f, 0.001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double groups_per_smsp = 8.0 * iters * 8;   // warps per SMSP x iterations x unrolled chains
    printf("%-44s %7.3f ms  %.2f cycles per SMSP and %s\n", NAMES[MODE], best, best * 1e-3 * mhz * 1e6 / groups_per_smsp,
           MODE >= MIX_FFMA_FFMA2 ? "group" : "warp-instruction");
}

int main(int argc, char** argv) {
    const double mhz = argc > 1 ? atof(argv[1]) : 1965.0;
    float* out; cudaMalloc(&out, 148 * 4 * 256 * 4);
    run<FFMA_REUSE>(out, mhz); run<FFMA_3REG>(out, mhz); run<FFMA2_PK>(out, mhz); run<FFMA2_BC>(out, mhz);
    run<FADD_>(out, mhz); run<FMUL_>(out, mhz); run<FADD2_>(out, mhz); run<FMUL2_BC>(out, mhz);
    run<MIX_FFMA_FFMA2>(out, mhz); run<MIX_FFMA_ALU>(out, mhz); run<MIX_FFMA2_ALU>(out, mhz); run<MIX_FFMA2_FADD>(out, mhz);
    run<MIX_FFMA_MUFU>(out, mhz);
    return cudaDeviceSynchronize() != cudaSuccess;
}
