// Event-marching trace kernels (tt_trace variants 3 and 4, and the FP32 fast path of tt_trace_aux).
#include "trace_event_ray.cuh"     // event_ray<T, SPC1>: the per-ray body of the scalar kernels (host + device)

#pragma nv_diag_suppress 550   // lo2()/hi2() unpack a register pair through asm and use one half each

namespace tt {

// ---- variant 3: event marching ---------------------------------------------------------------------
// Every RK4 step lies inside ONE grid cell: a step ends on the next (sub-)plane of the probing axis
// or, if the stage-1 slope predicts that the ray leaves its (u, v) cell column first, on that cell
// face (chord fraction lambda of the remaining interval); the ray is then relabelled into the
// neighbouring cell and continues.  Inside a cell the field is one trilinear polynomial
//      g(tu, tv, fw) = (A + tu B + tv (C + tu D)) + fw (A' + tu B' + tv (C' + tu D'))
// held in 24 registers per ray (7 FMA per component, no loads, no per-stage cell tests, no
// divergence inside a step); the next plane's corners are prefetched one step ahead.  All lanes of a
// warp stay converged in one loop (a lane with more cell crossings simply iterates a few more times),
// and because no step straddles a kink of the piecewise-trilinear field the integrator keeps its 4th
// order.  A predicted crossing lands within ~1e-4 cells of the face (chord vs arc); the ray keeps its
// true position (fractions may be ~1e-4 outside [0, 1], where the polynomial is simply extrapolated).
// Anything unusual -- launched outside the cube, steep or backward, side exit, possible time cap,
// non-finite state -- is flagged TT_RAY_DEFERRED and integrated by the general kernel in a second
// launch (none of the rays of a beam that fits the cube).
#ifndef TT_EVENT_MIN_BLOCKS
#define TT_EVENT_MIN_BLOCKS 5        // FP32: 95 registers, no spills -> 20 warps / SM
#endif
#ifndef TT_EVENT_BLOCK
#define TT_EVENT_BLOCK 128           // threads per CTA of the packed kernel (A/B: 64 with TT_EVENT_MIN_BLOCKS=10, 256 with 2)
#endif
#ifndef TT_EVENT_MIN_BLOCKS_AUX
#define TT_EVENT_MIN_BLOCKS_AUX 3    // passive quantities on board: four packed polynomials, 168 registers
#endif
#ifndef TT_EVENT_MIN_BLOCKS_F64
#define TT_EVENT_MIN_BLOCKS_F64 3
#endif
// Tri<T> / Bil<T> (the trilinear polynomial of one cell and its evaluation) live in trace_common.cuh.

template <typename T, bool SPC1>
__global__ void __launch_bounds__(128, sizeof(T) == 8 ? TT_EVENT_MIN_BLOCKS_F64 : TT_EVENT_MIN_BLOCKS)
trace_event_kernel(const typename GridT<T>::V4* __restrict__ grid, const double* __restrict__ s0,
                   const uint32_t* __restrict__ perm, double* __restrict__ rf, double* __restrict__ sf,
                   unsigned long long* __restrict__ ray_steps, uint8_t* __restrict__ status, TraceArgs A) {
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned steps = 0;
    if (tid < A.np) {
        const long ray = perm ? (long)perm[tid] : tid;
        bool deferred = false;
        steps = event_ray<T, SPC1>(grid, s0, ray, rf, sf, status, A, deferred);    // trace_event_ray.cuh
        if (deferred && A.any_deferred) *A.any_deferred = 1u;                       // (benign race: everybody stores 1)
    }
    if (ray_steps) {
        unsigned v = steps;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(ray_steps, (unsigned long long)v);
    }
}

// ---- variant 3 in FP32: the same event-marching kernel written with Blackwell's packed FP32x2 math ----
// sm_100 executes two FP32 FMAs per instruction on a 64-bit register pair (SASS FFMA2 / FADD2 / FMUL2,
// PTX fma.rn.f32x2, with a scalar-broadcast operand form).  The FP32 lanes are not faster, but the
// kernel is bound by instruction ISSUE (profiles/: issue slots 75 %, FMA pipe 57 %), so halving the
// instruction count of the FP work pays.  Pairs: the (u, v) fractions, the (u, v) direction
// components, and the float4 grid lanes as (g_u, g_v) and (g_w, ne/nc) -- the 4th lane rides along for
// free.  Every lane performs exactly the operations of the scalar kernel above, in the same order, so
// the two produce bit-identical rays (tested).
// (f32x2 helpers, Tri2 / Bil2 and the per-ray body event_ray_f32x2 live in trace_event_ray.cuh: host + device)

// CUBIC: h_u = h_v = h_w, so the index-space slopes need no rescaling (x * 1.0f is exact: same bits).
// TRACK_S: sf (state at time T) wanted -> the path time is accumulated; false drops that bookkeeping from the loop.
template <bool SPC1, bool AUX, bool CUBIC, bool TRACK_S>
__global__ void __launch_bounds__(TT_EVENT_BLOCK, AUX ? TT_EVENT_MIN_BLOCKS_AUX : TT_EVENT_MIN_BLOCKS)
trace_event_kernel_f32x2(const float4* __restrict__ grid, const double* __restrict__ s0,
                         const uint32_t* __restrict__ perm, double* __restrict__ rf, double* __restrict__ sf,
                         unsigned long long* __restrict__ ray_steps, uint8_t* __restrict__ status, TraceArgs A,
                         const float4* __restrict__ aux4 = nullptr, double* __restrict__ aux_out = nullptr,
                         AuxArgs AX = AuxArgs()) {
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned steps = 0;
    if (tid < A.np) {
        const long ray = perm ? (long)perm[tid] : tid;
        bool deferred = false;
        steps = event_ray_f32x2<SPC1, AUX, CUBIC, TRACK_S>(grid, s0, ray, rf, sf, status, A, aux4, aux_out, AX, deferred);
        if (deferred && A.any_deferred) *A.any_deferred = 1u;       // (benign race: everybody stores 1)
    }
    if (ray_steps) {
        unsigned v = steps;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(ray_steps, (unsigned long long)v);
    }
}


int launch_trace_event(int dtype, bool packed, int steps_per_cell, const void* grid4, const double* s0,
                       const uint32_t* perm, double* rf, double* sf, unsigned long long* ray_steps, uint8_t* status,
                       const TraceArgs& A, const void* aux4, double* aux_out, const AuxArgs& AX, cudaStream_t s) {
    const int block = TT_EVENT_BLOCK;      // packed kernel
    const unsigned blocks = (unsigned)((A.np + block - 1) / block);
    const int block_s = 128;               // scalar kernels (their launch bounds)
    const unsigned blocks_s = (unsigned)((A.np + block_s - 1) / block_s);
    const bool spc1 = steps_per_cell == 1;
    const bool cubic = A.ruf == 1.0f && A.rvf == 1.0f;
#define TT_EV2(S1, AX_, CU, ...)                                                                          \
    do {                                                                                                  \
        if (AX_ || sf) trace_event_kernel_f32x2<S1, AX_, CU, true><<<blocks, block, 0, s>>>(__VA_ARGS__);  \
        else trace_event_kernel_f32x2<S1, AX_, CU, AX_><<<blocks, block, 0, s>>>(__VA_ARGS__);             \
    } while (0)
#define TT_EV2_DISPATCH(AX_, ...)                                                       \
    do {                                                                                \
        if (spc1) { if (cubic) TT_EV2(true, AX_, true, __VA_ARGS__); else TT_EV2(true, AX_, false, __VA_ARGS__); }   \
        else { if (cubic) TT_EV2(false, AX_, true, __VA_ARGS__); else TT_EV2(false, AX_, false, __VA_ARGS__); }      \
    } while (0)
    if (dtype == TT_F32 && aux_out) {
        TT_EV2_DISPATCH(true, (const float4*)grid4, s0, perm, rf, sf, ray_steps, status, A, (const float4*)aux4, aux_out, AX);
        return launch_check("trace_event_kernel_f32x2<aux>");
    }
    if (dtype == TT_F32 && packed) {
        TT_EV2_DISPATCH(false, (const float4*)grid4, s0, perm, rf, sf, ray_steps, status, A);
        return launch_check("trace_event_kernel_f32x2");
    }
#undef TT_EV2_DISPATCH
#undef TT_EV2
    if (dtype == TT_F32) {
        if (spc1) trace_event_kernel<float, true><<<blocks_s, block_s, 0, s>>>((const float4*)grid4, s0, perm, rf, sf, ray_steps, status, A);
        else trace_event_kernel<float, false><<<blocks_s, block_s, 0, s>>>((const float4*)grid4, s0, perm, rf, sf, ray_steps, status, A);
    } else {
        if (spc1) trace_event_kernel<double, true><<<blocks_s, block_s, 0, s>>>((const double4*)grid4, s0, perm, rf, sf, ray_steps, status, A);
        else trace_event_kernel<double, false><<<blocks_s, block_s, 0, s>>>((const double4*)grid4, s0, perm, rf, sf, ray_steps, status, A);
    }
    return launch_check("trace_event_kernel");
}

}  // namespace tt
