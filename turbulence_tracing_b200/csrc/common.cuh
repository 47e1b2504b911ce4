// Shared helpers for libtt_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "tt_b200.h"

namespace tt {

// thread-local error message returned by tt_last_error()
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define TT_CUDA(call)                                              \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return ::tt::cuda_fail(e_, #call);  \
    } while (0)

#define TT_REQUIRE(cond, ...)            \
    do {                                 \
        if (!(cond)) {                   \
            ::tt::set_error(__VA_ARGS__); \
            return TT_ERR_INVALID;       \
        }                                \
    } while (0)

// 4-byte stream-ordered scratch word, zeroed (nullptr if it cannot be served), from a memory pool owned by this
// library (api.cu) that keeps what it has been given: with the default pool's release threshold of 0 the pool hands
// its memory back to the driver at every synchronisation, and the next cudaMallocAsync has to map fresh physical
// memory -- milliseconds up to (measured on a B200 VM) hundreds of milliseconds per trace launch.
unsigned int* scratch_flag(cudaStream_t s);
void* scratch_alloc(size_t bytes, cudaStream_t s);       // stream-ordered scratch from the same pool (nullptr on failure)

void count_launch();             // api.cu: process-wide counter behind tt_launch_count()
inline int launch_check(const char* name) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, name);
    count_launch();
    return TT_OK;
}

// TT_HD: helpers that also compile for the host, so that the per-ray bodies built from them can be run on
// the CPU by the test harness (tests/host/: the same source, no GPU needed).  No effect on the device code.
#define TT_HD __host__ __device__ __forceinline__

// read-only scalar load (node tables, histogram edges)
TT_HD double ldg_f64(const double* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
// one IEEE operation, never contracted into an FMA (numpy rounds every product and sum)
TT_HD double mul_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
TT_HD double add_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}

// frame permutation: (u, v, w) = (t1, t2, par) -> index into (x, y, z)
struct Frame {
    int a[3];
};
__host__ __device__ inline Frame frame_of(int par) {
    Frame f;
    if (par == 2) { f.a[0] = 0; f.a[1] = 1; f.a[2] = 2; }
    else if (par == 1) { f.a[0] = 0; f.a[1] = 2; f.a[2] = 1; }
    else { f.a[0] = 1; f.a[1] = 2; f.a[2] = 0; }
    return f;
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: any ray / voxel is addressable -----
struct Philox {
    uint32_t c[4];
};
__host__ __device__ inline Philox philox4x32_10(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t key) {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32);
    uint32_t c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox o;
    o.c[0] = c0; o.c[1] = c1; o.c[2] = c2; o.c[3] = c3;
    return o;
}
// uniform in (0, 1) with 53 random bits
__host__ __device__ inline double u01(uint32_t hi, uint32_t lo) {
    uint64_t b = (((uint64_t)hi << 32) | lo) >> 11;
    return ((double)b + 0.5) * (1.0 / 9007199254740992.0);
}

}  // namespace tt
