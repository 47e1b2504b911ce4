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
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo2(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)b; return a; }
__device__ __forceinline__ float hi2(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)a; return b; }
__device__ __forceinline__ f32x2 bc2(float s) { return pk2(s, s); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

struct Tri2 {          // two trilinear polynomials (one per lane) of one cell
    f32x2 a, b, c, d, a1, b1, c1, d1;
};
struct Bil2 {
    f32x2 a, b, c, d;
};
__device__ __forceinline__ Bil2 tri2_at(const Tri2& q, f32x2 FW) {
    Bil2 r;
    r.a = fma2(FW, q.a1, q.a); r.b = fma2(FW, q.b1, q.b); r.c = fma2(FW, q.c1, q.c); r.d = fma2(FW, q.d1, q.d);
    return r;
}
__device__ __forceinline__ f32x2 bil2_eval(const Bil2& q, f32x2 TU, f32x2 TV) {
    return fma2(TV, fma2(TU, q.d, q.c), fma2(TU, q.b, q.a));
}
__device__ __forceinline__ void tri2_set(Tri2& q, f32x2 c00, f32x2 c10, f32x2 c01, f32x2 c11, f32x2 e00, f32x2 e10,
                                         f32x2 e01, f32x2 e11) {
    q.a = c00; q.b = sub2(c10, c00); q.c = sub2(c01, c00); q.d = sub2(sub2(c11, c01), q.b);
    f32x2 eb = sub2(e10, e00), ec = sub2(e01, e00), ed = sub2(sub2(e11, e01), eb);
    q.a1 = sub2(e00, q.a); q.b1 = sub2(eb, q.b); q.c1 = sub2(ec, q.c); q.d1 = sub2(ed, q.d);
}
__device__ __forceinline__ void tri2_advance(Tri2& q, f32x2 n00, f32x2 n10, f32x2 n01, f32x2 n11) {
    q.a = add2(q.a, q.a1); q.b = add2(q.b, q.b1); q.c = add2(q.c, q.c1); q.d = add2(q.d, q.d1);
    f32x2 eb = sub2(n10, n00);
    q.a1 = sub2(n00, q.a); q.b1 = sub2(eb, q.b); q.c1 = sub2(sub2(n01, n00), q.c);
    q.d1 = sub2(sub2(sub2(n11, n01), eb), q.d);
}
#define TT_XY(v) pk2((v).x, (v).y)
#define TT_ZW(v) pk2((v).z, (v).w)

// AUX = true additionally carries the passive quantities of tt_trace_aux (phase, Faraday rotation,
// absorption): the ne/nc lane rides with g_w as a packed pair, a second grid (B_u, B_v | B_w, kappa) gets
// two more packed polynomials, and the RK4 stages double as Simpson nodes of the three line integrals.
__device__ __forceinline__ void aux_integrands(float nn, f32x2 bxy, f32x2 bzk, f32x2 duv, float dw, float hq, bool has_b,
                                               float& fp, float& ff, float& fa) {
    const float r = sqrtf(fmaxf(1.f - nn, 0.f));
    fp = -nn / (1.f + r) * hq;                      // (sqrt(1 - ne/nc) - 1) ds, without cancellation
    ff = 0.f; fa = 0.f;
    if (has_b) {
        const float bd = fmaf(lo2(bxy), lo2(duv), fmaf(hi2(bxy), hi2(duv), lo2(bzk) * dw));
        ff = nn * bd * hq;                          // (ne/nc) (B . d) ds
        fa = hi2(bzk) * hq;                         // kappa ds
    }
}

// CUBIC: h_u = h_v = h_w, so the index-space slopes need no rescaling (x * 1.0f is exact: same bits).
template <bool SPC1, bool AUX, bool CUBIC>
__global__ void __launch_bounds__(128, AUX ? TT_EVENT_MIN_BLOCKS_AUX : TT_EVENT_MIN_BLOCKS)
trace_event_kernel_f32x2(const float4* __restrict__ grid, const double* __restrict__ s0,
                         const uint32_t* __restrict__ perm, double* __restrict__ rf, double* __restrict__ sf,
                         unsigned long long* __restrict__ ray_steps, uint8_t* __restrict__ status, TraceArgs A,
                         const float4* __restrict__ aux4 = nullptr, double* __restrict__ aux_out = nullptr,
                         AuxArgs AX = AuxArgs()) {
    typedef float T;
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned steps = 0;
    if (tid < A.np) {
        const long ray = perm ? (long)perm[tid] : tid;
        const int nu = A.n[0], nv = A.n[1], nw = A.n[2];
        const long long plane = A.plane_elems;
        // ---- prologue (identical to the scalar kernel) ----------------------------------------------
        double X[3], D[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            X[k] = (s0[(size_t)A.fa[k] * A.np + ray] - A.o[k]) / A.h[k];
            D[k] = s0[(size_t)(3 + A.fa[k]) * A.np + ray] * (1.0 / kC);
        }
        // a ray launched in front of the cube (asymmetric axes: the reference launches at -extent whatever
        // the axis starts at) flies freely to the entry face first; the field is zero out there
        double s_pre = 0.0;
        if (X[2] < 0.0 && D[2] > TT_MARCH_MIN_DW) {
            s_pre = -X[2] * A.h[2] / D[2];
            X[0] += D[0] / A.h[0] * s_pre;
            X[1] += D[1] / A.h[1] * s_pre;
            X[2] = 0.0;
        }
        bool fast = X[0] >= 0.0 && X[0] <= (double)(nu - 1) && X[1] >= 0.0 && X[1] <= (double)(nv - 1) &&
                    X[2] >= 0.0 && X[2] <= (double)(nw - 1) && D[2] > TT_MARCH_MIN_DW;
        fast = fast && ((double)(nw - 1) - X[2]) * A.h[2] <= TT_MARCH_MIN_DW * (A.s_max - s_pre);
        int cu = 0, cv = 0, k = 0;
        T tu0 = 0.f, tv0 = 0.f, fw = 0.f;
        if (fast) {
            double fl;
            fl = fmin(floor(X[0]), (double)(nu - 2)); cu = (int)fl; tu0 = (T)(X[0] - fl);
            fl = fmin(floor(X[1]), (double)(nv - 2)); cv = (int)fl; tv0 = (T)(X[1] - fl);
            fl = floor(X[2]); k = (int)fl; fw = (T)(X[2] - fl);
        }
        f32x2 tuv = pk2(tu0, tv0), duv = pk2((T)D[0], (T)D[1]);
        T dw = (T)D[2], s = 0.f;
        const T hw = A.hwf;
        const f32x2 RUV = pk2(A.ruf, A.rvf);
        const bool track_s = sf != nullptr;
        const int spc = A.spc;
        const T hsub = SPC1 ? 1.f : 1.f / (T)spc;
        int j = SPC1 ? 0 : (int)(fw * (T)spc);
        double acc_p = 0.0, acc_f = 0.0, acc_a = 0.0;     // AUX: line integrals of (n-1), (ne/nc)(B.d), kappa

        if (fast && k < nw - 1) {
            const float4* p = grid + ((size_t)k * plane + (size_t)cv * nu + cu);
            Tri2 qxy;                 // (g_u, g_v) lanes, packed
            Tri<float> qz;            // g_w, scalar (packing it with the unused ne/nc lane would only
                                      // add work to the FP32 pipe, which is what bounds this kernel)
            Tri2 qzw, bxy, bzk;       // AUX: (g_w, ne/nc), (B_u, B_v), (B_w, kappa)
            const bool has_b = AUX && aux4 != nullptr;
            const float4* pa = has_b ? aux4 + (p - grid) : nullptr;
            float4 n00, n10, n01, n11;
            // (re)build the polynomials of the current cell from planes k and k+1
            auto load_cell = [&]() {
                float4 c00 = __ldg(p), c10 = __ldg(p + 1), c01 = __ldg(p + nu), c11 = __ldg(p + nu + 1);
                const float4* p1 = p + plane;
                float4 e00 = __ldg(p1), e10 = __ldg(p1 + 1), e01 = __ldg(p1 + nu), e11 = __ldg(p1 + nu + 1);
                tri2_set(qxy, TT_XY(c00), TT_XY(c10), TT_XY(c01), TT_XY(c11), TT_XY(e00), TT_XY(e10), TT_XY(e01), TT_XY(e11));
                if (AUX) tri2_set(qzw, TT_ZW(c00), TT_ZW(c10), TT_ZW(c01), TT_ZW(c11), TT_ZW(e00), TT_ZW(e10), TT_ZW(e01), TT_ZW(e11));
                else tri_set<float>(qz, c00.z, c10.z, c01.z, c11.z, e00.z, e10.z, e01.z, e11.z);
                if (has_b) {
                    c00 = __ldg(pa); c10 = __ldg(pa + 1); c01 = __ldg(pa + nu); c11 = __ldg(pa + nu + 1);
                    const float4* q1 = pa + plane;
                    e00 = __ldg(q1); e10 = __ldg(q1 + 1); e01 = __ldg(q1 + nu); e11 = __ldg(q1 + nu + 1);
                    tri2_set(bxy, TT_XY(c00), TT_XY(c10), TT_XY(c01), TT_XY(c11), TT_XY(e00), TT_XY(e10), TT_XY(e01), TT_XY(e11));
                    tri2_set(bzk, TT_ZW(c00), TT_ZW(c10), TT_ZW(c01), TT_ZW(c11), TT_ZW(e00), TT_ZW(e10), TT_ZW(e01), TT_ZW(e11));
                }
            };
            load_cell();
            bool have_next = false;
            while (true) {
                if (!have_next && k + 2 <= nw - 1) {
                    const float4* p2 = p + 2 * plane;
                    n00 = __ldg(p2); n10 = __ldg(p2 + 1); n01 = __ldg(p2 + nu); n11 = __ldg(p2 + nu + 1);
                    have_next = true;
                }
                // ---- stage 1 and the length of this step -------------------------------------------
                T q = trcp<T>(dw), hq = hw * q;
                bool ok = dw > T(TT_MARCH_MIN_DW);
                f32x2 TU = bc2(lo2(tuv)), TV = bc2(hi2(tuv));
                const f32x2 aUV = CUBIC ? mul2(duv, bc2(q)) : mul2(mul2(RUV, duv), bc2(q));
                const f32x2 aduv = mul2(bil2_eval(tri2_at(qxy, bc2(fw)), TU, TV), bc2(hq));
                T adw, as = hq;
                float fp1 = 0.f, ff1 = 0.f, fa1 = 0.f, fp2 = 0.f, ff2 = 0.f, fa2 = 0.f, fp3 = 0.f, ff3 = 0.f, fa3 = 0.f,
                      fp4 = 0.f, ff4 = 0.f, fa4 = 0.f;
                if (AUX) {
                    const f32x2 FW = bc2(fw);
                    const f32x2 gzw = bil2_eval(tri2_at(qzw, FW), TU, TV);
                    adw = lo2(gzw) * hq;
                    f32x2 b1 = 0, b2 = 0;
                    if (has_b) { b1 = bil2_eval(tri2_at(bxy, FW), TU, TV); b2 = bil2_eval(tri2_at(bzk, FW), TU, TV); }
                    aux_integrands(hi2(gzw), b1, b2, duv, dw, hq, has_b, fp1, ff1, fa1);
                } else {
                    adw = bil_eval<float>(tri_at<float>(qz, fw), lo2(tuv), hi2(tuv)) * hq;
                }
                const T fw_t = SPC1 ? 1.f : ((j + 1 == spc) ? 1.f : (T)(j + 1) * hsub);
                T h = fw_t - fw;
                int cross = 0;
                {
                    const f32x2 puv = fma2(bc2(h), aUV, tuv);
                    const T pu = lo2(puv), pv = hi2(puv);
                    if (pu > 1.f || pu < 0.f || pv > 1.f || pv < 0.f) {
                        const T aU = lo2(aUV), aV = hi2(aUV), tu = lo2(tuv), tv = hi2(tuv);
                        // (a branch-free variant with approximate divisions was measured slower: 459.6 vs 451.4 ms)
                        T lu = 2.f, lv = 2.f;
                        if (aU > 0.f) lu = (1.f - tu) / (h * aU); else if (aU < 0.f) lu = -tu / (h * aU);
                        if (aV > 0.f) lv = (1.f - tv) / (h * aV); else if (aV < 0.f) lv = -tv / (h * aV);
                        T lam = fminf(lu, lv);
                        if (lam < 1.f) {
                            cross = lu <= lv ? (aU > 0.f ? 1 : -1) : (aV > 0.f ? 2 : -2);
                            h *= lam > 0.f ? lam : 0.f;
                        }
                    }
                }
                const T half = 0.5f * h;
                const f32x2 HALF = bc2(half), H = bc2(h);
                // ---- stages 2-4 ---------------------------------------------------------------------
                f32x2 suv = fma2(HALF, aUV, tuv), duv2 = fma2(HALF, aduv, duv);
                T sw = fw + half, dw2 = fmaf(half, adw, dw);
                q = trcp<T>(dw2); hq = hw * q; ok = ok && dw2 > 0.f;
                TU = bc2(lo2(suv)); TV = bc2(hi2(suv));
                const Bil2 mxy = tri2_at(qxy, bc2(sw));             // stages 2 and 3 share their w-fraction
                Bil<float> mz;
                Bil2 mzw, mb1, mb2;
                if (AUX) {
                    mzw = tri2_at(qzw, bc2(sw));
                    if (has_b) { mb1 = tri2_at(bxy, bc2(sw)); mb2 = tri2_at(bzk, bc2(sw)); }
                } else {
                    mz = tri_at<float>(qz, sw);
                }
                const f32x2 bUV = CUBIC ? mul2(duv2, bc2(q)) : mul2(mul2(RUV, duv2), bc2(q));
                const f32x2 bduv = mul2(bil2_eval(mxy, TU, TV), bc2(hq));
                T bdw, bs = hq;
                if (AUX) {
                    const f32x2 gzw = bil2_eval(mzw, TU, TV);
                    bdw = lo2(gzw) * hq;
                    f32x2 b1 = 0, b2 = 0;
                    if (has_b) { b1 = bil2_eval(mb1, TU, TV); b2 = bil2_eval(mb2, TU, TV); }
                    aux_integrands(hi2(gzw), b1, b2, duv2, dw2, hq, has_b, fp2, ff2, fa2);
                } else {
                    bdw = bil_eval<float>(mz, lo2(suv), hi2(suv)) * hq;
                }
                suv = fma2(HALF, bUV, tuv); duv2 = fma2(HALF, bduv, duv); dw2 = fmaf(half, bdw, dw);
                q = trcp<T>(dw2); hq = hw * q; ok = ok && dw2 > 0.f;
                TU = bc2(lo2(suv)); TV = bc2(hi2(suv));
                const f32x2 cUV = CUBIC ? mul2(duv2, bc2(q)) : mul2(mul2(RUV, duv2), bc2(q));
                const f32x2 cduv = mul2(bil2_eval(mxy, TU, TV), bc2(hq));
                T cdw, cs = hq;
                if (AUX) {
                    const f32x2 gzw = bil2_eval(mzw, TU, TV);
                    cdw = lo2(gzw) * hq;
                    f32x2 b1 = 0, b2 = 0;
                    if (has_b) { b1 = bil2_eval(mb1, TU, TV); b2 = bil2_eval(mb2, TU, TV); }
                    aux_integrands(hi2(gzw), b1, b2, duv2, dw2, hq, has_b, fp3, ff3, fa3);
                } else {
                    cdw = bil_eval<float>(mz, lo2(suv), hi2(suv)) * hq;
                }
                suv = fma2(H, cUV, tuv); duv2 = fma2(H, cduv, duv); dw2 = fmaf(h, cdw, dw); sw = fw + h;
                q = trcp<T>(dw2); hq = hw * q; ok = ok && dw2 > 0.f;
                TU = bc2(lo2(suv)); TV = bc2(hi2(suv));
                const f32x2 eUV = CUBIC ? mul2(duv2, bc2(q)) : mul2(mul2(RUV, duv2), bc2(q));
                const f32x2 eduv = mul2(bil2_eval(tri2_at(qxy, bc2(sw)), TU, TV), bc2(hq));
                T edw, es = hq;
                if (AUX) {
                    const f32x2 FW = bc2(sw);
                    const f32x2 gzw = bil2_eval(tri2_at(qzw, FW), TU, TV);
                    edw = lo2(gzw) * hq;
                    f32x2 b1 = 0, b2 = 0;
                    if (has_b) { b1 = bil2_eval(tri2_at(bxy, FW), TU, TV); b2 = bil2_eval(tri2_at(bzk, FW), TU, TV); }
                    aux_integrands(hi2(gzw), b1, b2, duv2, dw2, hq, has_b, fp4, ff4, fa4);
                } else {
                    edw = bil_eval<float>(tri_at<float>(qz, sw), lo2(suv), hi2(suv)) * hq;
                }
                const T h6 = h * T(1.0 / 6.0);
                const f32x2 H6 = bc2(h6), TWO = bc2(2.f);
                tuv = fma2(H6, add2(add2(aUV, mul2(TWO, add2(bUV, cUV))), eUV), tuv);
                duv = fma2(H6, add2(add2(aduv, mul2(TWO, add2(bduv, cduv))), eduv), duv);
                dw = fmaf(h6, adw + 2.f * (bdw + cdw) + edw, dw);
                if (track_s) s = fmaf(h6, as + 2.f * (bs + cs) + es, s);
                if (AUX) {                      // Simpson over the four stages, summed in FP64
                    acc_p += (double)(h6 * (fp1 + 2.f * (fp2 + fp3) + fp4));
                    if (has_b) {
                        acc_f += (double)(h6 * (ff1 + 2.f * (ff2 + ff3) + ff4));
                        acc_a += (double)(h6 * (fa1 + 2.f * (fa2 + fa3) + fa4));
                    }
                }
                if (!(ok && dw > T(TT_MARCH_MIN_DW))) { fast = false; break; }
                if (cross == 0) {
                    ++steps;
                    fw = fw_t;
                    if (SPC1 || ++j == spc) {
                        j = 0; fw = 0.f;
                        if (++k >= nw - 1) break;
                        p += plane;
                        if (TT_EVENT_PREFETCH && k + TT_EVENT_PREFETCH <= nw - 1) {   // register-free L1 prefetch
                            prefetch_l1(p + TT_EVENT_PREFETCH * plane);
                            prefetch_l1(p + TT_EVENT_PREFETCH * plane + nu);
                        }
                        tri2_advance(qxy, TT_XY(n00), TT_XY(n10), TT_XY(n01), TT_XY(n11));
                        if (AUX) tri2_advance(qzw, TT_ZW(n00), TT_ZW(n10), TT_ZW(n01), TT_ZW(n11));
                        else tri_advance<float>(qz, n00.z, n10.z, n01.z, n11.z);
                        if (has_b) {
                            pa += plane;
                            const float4* q1 = pa + plane;
                            const float4 b00 = __ldg(q1), b10 = __ldg(q1 + 1), b01 = __ldg(q1 + nu), b11 = __ldg(q1 + nu + 1);
                            tri2_advance(bxy, TT_XY(b00), TT_XY(b10), TT_XY(b01), TT_XY(b11));
                            tri2_advance(bzk, TT_ZW(b00), TT_ZW(b10), TT_ZW(b01), TT_ZW(b11));
                        }
                        have_next = false;
                    }
                } else {
                    fw += h;
                    T tu = lo2(tuv), tv = hi2(tuv);
                    int dp = 0;
                    if (cross == 1) { ++cu; tu -= 1.f; dp = 1; } else if (cross == -1) { --cu; tu += 1.f; dp = -1; }
                    else if (cross == 2) { ++cv; tv -= 1.f; dp = nu; } else { --cv; tv += 1.f; dp = -nu; }
                    p += dp;
                    if (has_b) pa += dp;
                    tuv = pk2(tu, tv);
                    if (cu < 0 || cu > nu - 2 || cv < 0 || cv > nv - 2) { fast = false; break; }
                    load_cell();
                    have_next = false;
                }
            }
        }
        if (!fast) {
            status[ray] = TT_RAY_DEFERRED;
            if (A.any_deferred) *A.any_deferred = 1u;
            steps = 0;
        } else {
            const double Pu = A.o[0] + ((double)cu + (double)lo2(tuv)) * A.h[0];
            const double Pv = A.o[1] + ((double)cv + (double)hi2(tuv)) * A.h[1];
            const double Pw = A.o[2] + (double)(nw - 1) * A.h[2];
            const double Vu = (double)lo2(duv) * kC, Vv = (double)hi2(duv) * kC, Vw = (double)dw * kC;
            const double tb = (Pw - A.extent) / Vw;
            rf[0 * A.np + ray] = Pu - Vu * tb;
            rf[1 * A.np + ray] = atan(Vu / Vw);
            rf[2 * A.np + ray] = Pv - Vv * tb;
            rf[3 * A.np + ray] = atan(Vv / Vw);
            if (sf) {
                const double t_rest = (A.s_max - s_pre - (double)s) / kC;
                const double Pf[3] = {Pu, Pv, Pw}, Vf[3] = {Vu, Vv, Vw};
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    sf[(size_t)A.fa[m] * A.np + ray] = Pf[m] + Vf[m] * t_rest;
                    sf[(size_t)(3 + A.fa[m]) * A.np + ray] = Vf[m];
                }
            }
            if (AUX) {
                aux_out[0 * A.np + ray] = exp(-0.5 * acc_a);
                aux_out[1 * A.np + ray] = AX.omega_over_c * acc_p;
                aux_out[2 * A.np + ray] = AX.verdet_nc * acc_f;
            }
            status[ray] = (uint8_t)TT_RAY_EXIT_FACE;
        }
    }
    if (ray_steps) {
        unsigned v = steps;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(ray_steps, (unsigned long long)v);
    }
}
#undef TT_XY
#undef TT_ZW


int launch_trace_event(int dtype, bool packed, int steps_per_cell, const void* grid4, const double* s0,
                       const uint32_t* perm, double* rf, double* sf, unsigned long long* ray_steps, uint8_t* status,
                       const TraceArgs& A, const void* aux4, double* aux_out, const AuxArgs& AX, cudaStream_t s) {
    const int block = 128;
    const unsigned blocks = (unsigned)((A.np + block - 1) / block);
    const bool spc1 = steps_per_cell == 1;
    const bool cubic = A.ruf == 1.0f && A.rvf == 1.0f;
#define TT_EV2(S1, AX_, CU, ...) trace_event_kernel_f32x2<S1, AX_, CU><<<blocks, block, 0, s>>>(__VA_ARGS__)
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
        if (spc1) trace_event_kernel<float, true><<<blocks, block, 0, s>>>((const float4*)grid4, s0, perm, rf, sf, ray_steps, status, A);
        else trace_event_kernel<float, false><<<blocks, block, 0, s>>>((const float4*)grid4, s0, perm, rf, sf, ray_steps, status, A);
    } else {
        if (spc1) trace_event_kernel<double, true><<<blocks, block, 0, s>>>((const double4*)grid4, s0, perm, rf, sf, ray_steps, status, A);
        else trace_event_kernel<double, false><<<blocks, block, 0, s>>>((const double4*)grid4, s0, perm, rf, sf, ray_steps, status, A);
    }
    return launch_check("trace_event_kernel");
}

}  // namespace tt
