// Event marching over the FACE-COEFFICIENT grid: the production FP32 trace kernel of tt_trace_faces (1 step per cell).
//
// The event kernels of trace_event_ray.cuh rebuild the bilinear coefficients (A, B, C, D) of every cell face from its
// 4 corners for EVERY ray at EVERY plane: 8 loads and ~36 FP32-pipe slots of differences per ray-step, the same
// numbers for the ~10^3 rays that share a cell of the benchmark beam.  Here they are formed ONCE per face by
// face_grid_cell() (tt_build_face_grid, a streaming kernel over the float4 node grid) and stored as three 16-byte
// words per face cell, in the register-pair order the packed arithmetic wants:
//      word 0 = (A_u, A_v, B_u, B_v)      word 1 = (C_u, C_v, D_u, D_v)      word 2 = (A_w, C_w, B_w, D_w)
//      g_m(tu, tv) = A_m + tu B_m + tv (C_m + tu D_m)          on face k of the cell column (cu, cv)
// in CENTRED cell coordinates tu, tv in [-1/2, 1/2] (A = mean of the 4 corners, ...): the test "does the predicted end
// of the step lie outside the cell" is then |tu| > 1/2 || |tv| > 1/2, two compares with free |.| operand modifiers
// instead of four, and the polynomial is evaluated around the middle of the cell.
// with the step-size factors folded in (index-space march, independent variable = the w-fraction):
//      (A..D)_u = (h_w^2 / h_u) g_u,   (A..D)_v = (h_w^2 / h_v) g_v,   (A..D)_w = h_w g_w,    g = grad(ne/nc) * (-1/2)
// so that with the SCALED direction  e = (d_u h_w/h_u, d_v h_w/h_v, d_w)  the ray equations per unit w-fraction are
//      d(tu, tv)/dw = (e_u, e_v) / e_w        d e/dw = G(tu, tv, fw) / e_w
// -- no h_w / d_w product per stage, no slope rescaling for non-cubic cells.  Per plane the kernel loads 3 words
// (48 B, one cell face) instead of 4 corners, and renews its polynomial with 12 packed operations instead of 19.
// Same integrator, same event logic (a step ends on the next plane or on the predicted (u, v) cell face), same
// hand-over of unusual rays (TT_RAY_DEFERRED) as event_ray_f32x2; results agree with it to FP32 rounding (the
// coefficients are rounded once from FP64 differences instead of being differenced in FP32).
// Compiled for the device (trace_face.cu) and for the host (tests/host/trace_face_host.cu).
#pragma once
#include "trace_event_ray.cuh"

namespace tt {

struct FaceArgs {
    long long face_plane3;     // float4 words per face plane: 3 (nu-1)(nv-1)
    int row3;                  // float4 words per row of face cells: 3 (nu-1)
    float ruf, rvf;            // h_w/h_u, h_w/h_v: scaling of the transverse direction components
    double inv_ru, inv_rv;     // and back
};

inline void fill_face_args(FaceArgs& FA, const TraceArgs& A) {
    FA.row3 = 3 * (A.n[0] - 1);
    FA.face_plane3 = 3LL * (A.n[0] - 1) * (A.n[1] - 1);
    FA.ruf = (float)(A.h[2] / A.h[0]); FA.rvf = (float)(A.h[2] / A.h[1]);
    FA.inv_ru = 1.0 / (double)FA.ruf; FA.inv_rv = 1.0 / (double)FA.rvf;
}
// the factors tt_build_face_grid folds into the (u, v, w) coefficients: exactly the FP32 ratios the kernel scales with
inline void face_scales(const double h[3], double& su, double& sv, double& sw) {
    su = (double)(float)(h[2] / h[0]) * h[2]; sv = (double)(float)(h[2] / h[1]) * h[2]; sw = h[2];
}

// ---- builder: one face cell (k, cv, cu) of the coefficient grid from the 4 corners of the node grid --------------
TT_HD void face_grid_cell(const float4* __restrict__ grid, int nu, long long plane, int cu, int cv, int k, double su,
                          double sv, double sw, float4* __restrict__ out) {
    const float4* p = grid + ((size_t)k * plane + (size_t)cv * nu + cu);
    const float4 c00 = GridT<float>::ld(p), c10 = GridT<float>::ld(p + 1), c01 = GridT<float>::ld(p + nu), c11 = GridT<float>::ld(p + nu + 1);
    // centred: A = mean of the corners, B = mean u-difference, C = mean v-difference, D = the mixed difference
#define TT_FC(m, s, a, b, c, d)                                                                                          \
    const float a = (float)(s * (0.25 * (((double)c00.m + (double)c10.m) + ((double)c01.m + (double)c11.m)))),           \
                b = (float)(s * (0.5 * (((double)c10.m - (double)c00.m) + ((double)c11.m - (double)c01.m)))),            \
                c = (float)(s * (0.5 * (((double)c01.m - (double)c00.m) + ((double)c11.m - (double)c10.m)))),            \
                d = (float)(s * ((((double)c11.m - (double)c01.m) - (double)c10.m) + (double)c00.m));
    TT_FC(x, su, au, bu, cu_, du)
    TT_FC(y, sv, av, bv, cv_, dv)
    TT_FC(z, sw, aw, bw, cw_, dw)
#undef TT_FC
    out[0] = make_float4(au, av, bu, bv);
    out[1] = make_float4(cu_, cv_, du, dv);
    out[2] = make_float4(aw, cw_, bw, dw);
}

// ---- the polynomial of one face / one cell in packed registers -----------------------------------------------------
struct FaceQ {
    f32x2 a, b, c, d;      // (u, v) lanes
    f32x2 zac, zbd;        // w component: (A, C) and (B, D)
};
TT_HD FaceQ face_ld(const float4* __restrict__ p) {
    const float4 v0 = GridT<float>::ld(p), v1 = GridT<float>::ld(p + 1), v2 = GridT<float>::ld(p + 2);
    FaceQ q;
    q.a = pk2(v0.x, v0.y); q.b = pk2(v0.z, v0.w); q.c = pk2(v1.x, v1.y); q.d = pk2(v1.z, v1.w);
    q.zac = pk2(v2.x, v2.y); q.zbd = pk2(v2.z, v2.w);
    return q;
}
// the three words of a face as loaded: packed into register pairs only where they are consumed (a mov.b64 right behind
// the load would make the warp wait for it there, and the prefetch a whole step ahead would hide nothing)
struct FaceW {
    float4 v0, v1, v2;
};
TT_HD FaceW face_ldw(const float4* __restrict__ p) {
    FaceW w;
    w.v0 = GridT<float>::ld(p); w.v1 = GridT<float>::ld(p + 1); w.v2 = GridT<float>::ld(p + 2);
    return w;
}
TT_HD FaceQ face_pack(const FaceW& w) {
    FaceQ q;
    q.a = pk2(w.v0.x, w.v0.y); q.b = pk2(w.v0.z, w.v0.w); q.c = pk2(w.v1.x, w.v1.y); q.d = pk2(w.v1.z, w.v1.w);
    q.zac = pk2(w.v2.x, w.v2.y); q.zbd = pk2(w.v2.z, w.v2.w);
    return q;
}
TT_HD FaceQ face_sub(const FaceQ& f, const FaceQ& b) {
    FaceQ r;
    r.a = sub2(f.a, b.a); r.b = sub2(f.b, b.b); r.c = sub2(f.c, b.c); r.d = sub2(f.d, b.d);
    r.zac = sub2(f.zac, b.zac); r.zbd = sub2(f.zbd, b.zbd);
    return r;
}
TT_HD FaceQ face_add(const FaceQ& f, const FaceQ& b) {
    FaceQ r;
    r.a = add2(f.a, b.a); r.b = add2(f.b, b.b); r.c = add2(f.c, b.c); r.d = add2(f.d, b.d);
    r.zac = add2(f.zac, b.zac); r.zbd = add2(f.zbd, b.zbd);
    return r;
}
TT_HD FaceQ face_scale(const FaceQ& f, f32x2 C) {
    FaceQ r;
    r.a = mul2(f.a, C); r.b = mul2(f.b, C); r.c = mul2(f.c, C); r.d = mul2(f.d, C); r.zac = mul2(f.zac, C); r.zbd = mul2(f.zbd, C);
    return r;
}
TT_HD FaceQ face_at(const FaceQ& base, const FaceQ& primed, f32x2 FW) {
    FaceQ r;
    r.a = fma2(FW, primed.a, base.a); r.b = fma2(FW, primed.b, base.b); r.c = fma2(FW, primed.c, base.c);
    r.d = fma2(FW, primed.d, base.d); r.zac = fma2(FW, primed.zac, base.zac); r.zbd = fma2(FW, primed.zbd, base.zbd);
    return r;
}
// G(tu, tv): (u, v) lanes packed, w scalar
TT_HD void face_eval(const FaceQ& q, f32x2 tuv, f32x2& guv, float& gw) {
    const f32x2 TU = bc2(lo2(tuv)), TV = bc2(hi2(tuv));
    guv = fma2(TV, fma2(TU, q.d, q.c), fma2(TU, q.b, q.a));
    const f32x2 r = fma2(TU, q.zbd, q.zac);                 // (A + tu B, C + tu D)
    gw = fmaf(hi2(tuv), hi2(r), lo2(r));
}

#ifndef TT_FACE_REBASE
#define TT_FACE_REBASE 0       // (measured, not adopted: 305.3 vs 302.9 ms on 513^3 / 1e8 rays -- see below) every step runs the whole-cell arithmetic on the polynomial of ITS OWN w-interval: a lane
                               // that starts inside a cell (after a side crossing, or launched inside the cube) or stops at
                               // a side face rebases (B, P) once -- B' = B + fw P, P' = h P -- instead of the whole warp
                               // evaluating B + (fw + c h) P at every stage of every step in which any lane does.  No
                               // warp vote, one instantiation of the step: a ray's arithmetic never depends on its warp.
                               // ncu (r02_trace_face_v7_c3): the loop body is 112 instructions for every step instead of
                               // ~100 / ~150 (2/3 fast, 1/3 general), but a quarter of the warp's steps hold a side crossing
                               // of some lane and the prediction + column change with the rebase arithmetic cost ~112
                               // instructions there (1.6 lanes active): 139.6 vs 134.8 warp-instructions per warp-step.
                               // Host-tested (tests/test_host_kernels.py builds either form), GPU suite green with it.
#endif
#ifndef TT_FACE_PREFETCH
#define TT_FACE_PREFETCH 0     // experiment: planes beyond the register prefetch (k+2) whose face is pulled into L2
                               // (prefetch.global.L2): at 1025^3 every fourth face load is a first touch served by DRAM
#endif
#if TT_FACE_PREFETCH && defined(__CUDA_ARCH__)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif
#ifndef TT_FACE_FASTPATH
#define TT_FACE_FASTPATH 1     // warp-uniform shortcut for whole-cell steps: when every lane of the warp starts ON its
                               // plane and no lane predicts a side crossing, the stage coefficients are the base face,
                               // base + primed/2 and the far face (which IS the next base): 12 packed operations less
#endif

// One RK4 step of the ray inside its cell, from w-fraction fw over h (FULL: fw = 0 and h = 1 for every lane of the warp).
// In: base / primed polynomial, state (tuv, duv, dw), stage-1 slopes; out: the new state; FULL: B becomes the far face
// (in place: the base face is dead once the mid-step coefficients exist, and the far face is the next cell's base).
//
// A ray's bits must not depend on which of the two instantiations its warp happened to take (a chunked solve and a
// single launch sort the rays into different warps).  ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2
// in spite of the explicit rounding modifiers (and folds an fma2 by a literal 1 into that), so the step is written in
// already-fused form -- no packed product is ever the operand of a packed addition -- and h stays a run-time value in
// both instantiations; what differs is only how the stage coefficients are obtained, and those agree exactly
// (fma(0, P, B) = B, fma(1, P, B) = rn(B + P)).  Tested on the device: sorted == unsorted == chunked, bit for bit.
template <bool FULL, bool TRACK_S>
TT_HD bool face_step(FaceQ& B, const FaceQ& P, f32x2& tuv, f32x2& duv, float& dw, float& s, float fw, float h,
                     float q1, f32x2 aUV, f32x2 aduv, float adw) {
    const float half = 0.5f * h;
    const f32x2 HALF = bc2(half), H = bc2(h);
    // ---- stages 2 and 3 share their w-fraction ------------------------------------------------------------
    f32x2 suv = fma2(HALF, aUV, tuv), duv2 = fma2(HALF, aduv, duv);
    float dw2 = fmaf(half, adw, dw);
    const float q2 = trcp<float>(dw2);
    bool ok = dw2 > 0.f;
    const FaceQ M = face_at(B, P, bc2(FULL ? 0.5f : fw + half));
    const f32x2 bUV = mul2(duv2, bc2(q2));
    f32x2 g; float gw;
    face_eval(M, suv, g, gw);
    const f32x2 bduv = mul2(g, bc2(q2));
    const float bdw = gw * q2;
    suv = fma2(HALF, bUV, tuv); duv2 = fma2(HALF, bduv, duv); dw2 = fmaf(half, bdw, dw);
    const float q3 = trcp<float>(dw2);
    ok = ok && dw2 > 0.f;
    const f32x2 Q3 = bc2(q3), cUV = mul2(duv2, Q3), sUV = fma2(duv2, Q3, bUV);              // sUV = b + c
    face_eval(M, suv, g, gw);
    const f32x2 cduv = mul2(g, Q3), sduv = fma2(g, Q3, bduv);
    const float cdw = gw * q3, sdw = fmaf(gw, q3, bdw);
    // ---- stage 4 at the end of the step -------------------------------------------------------------------
    suv = fma2(H, cUV, tuv); duv2 = fma2(H, cduv, duv); dw2 = fmaf(h, cdw, dw);
    const float q4 = trcp<float>(dw2);
    ok = ok && dw2 > 0.f;
    if (FULL) { B = face_add(B, P); face_eval(B, suv, g, gw); }
    else face_eval(face_at(B, P, bc2(fw + h)), suv, g, gw);
    const float h6 = h * (float)(1.0 / 6.0);
    const f32x2 H6 = bc2(h6), TWO = bc2(2.f), Q4 = bc2(q4);
    tuv = fma2(H6, fma2(duv2, Q4, fma2(TWO, sUV, aUV)), tuv);                                // a + 2 (b + c) + e
    duv = fma2(H6, fma2(g, Q4, fma2(TWO, sduv, aduv)), duv);
    dw = fmaf(h6, fmaf(gw, q4, fmaf(2.f, sdw, adw)), dw);
    if (TRACK_S) s = fmaf(h6, q4 + fmaf(2.f, q2 + q3, q1), s);
    return ok;
}

// Returns the plane arrivals of this ray (0 if it is deferred to the general kernel).  faces: the coefficient grid of
// tt_build_face_grid for the same node grid and probing direction.  active: this thread holds a ray (the warp-uniform
// votes of the fast path need every lane of the warp in the loop).
template <bool TRACK_S>
TT_HD unsigned face_ray_f32x2(const float4* __restrict__ faces, const double* __restrict__ s0, long ray, double* __restrict__ rf,
                              double* __restrict__ sf, uint8_t* __restrict__ status, const TraceArgs& A, const FaceArgs& FA,
                              bool& deferred) {
    unsigned steps = 0;
    const int nu = A.n[0], nv = A.n[1], nw = A.n[2];
    const long long fplane = FA.face_plane3;
    const int row3 = FA.row3;
    // ---- prologue (the same entry conditions as event_ray_f32x2) ---------------------------------------------
    double X[3], D[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        X[k] = (s0[(size_t)A.fa[k] * A.np + ray] - A.o[k]) / A.h[k];
        D[k] = s0[(size_t)(3 + A.fa[k]) * A.np + ray] * (1.0 / kC);
    }
    double s_pre = 0.0;
    if (X[2] < 0.0 && D[2] > TT_MARCH_MIN_DW) {           // launched in front of the cube: free flight to the entry face
        s_pre = -X[2] * A.h[2] / D[2];
        X[0] += D[0] / A.h[0] * s_pre;
        X[1] += D[1] / A.h[1] * s_pre;
        X[2] = 0.0;
    }
    bool fast = X[0] >= 0.0 && X[0] <= (double)(nu - 1) && X[1] >= 0.0 && X[1] <= (double)(nv - 1) &&
                X[2] >= 0.0 && X[2] <= (double)(nw - 1) && D[2] > TT_MARCH_MIN_DW;
    fast = fast && ((double)(nw - 1) - X[2]) * A.h[2] <= TT_MARCH_MIN_DW * (A.s_max - s_pre);
    int cu = 0, cv = 0, k = 0;
    float tu0 = 0.f, tv0 = 0.f, fw = 0.f;
    if (fast) {
        double fl;
        fl = fmin(floor(X[0]), (double)(nu - 2)); cu = (int)fl; tu0 = (float)((X[0] - fl) - 0.5);     // centred
        fl = fmin(floor(X[1]), (double)(nv - 2)); cv = (int)fl; tv0 = (float)((X[1] - fl) - 0.5);
        fl = floor(X[2]); k = (int)fl; fw = (float)(X[2] - fl);
    }
    f32x2 tuv = pk2(tu0, tv0), duv = pk2((float)D[0] * FA.ruf, (float)D[1] * FA.rvf);     // scaled transverse direction
    float dw = (float)D[2], s = 0.f;
    const bool track_s = TRACK_S && sf != nullptr;

    if (fast && k < nw - 1) {
        const float4* p = faces + ((size_t)k * fplane + (size_t)cv * row3 + (size_t)cu * 3);
        FaceQ B = face_ld(p), P;
        FaceW N;
        P = face_sub(face_ld(p + fplane), B);
#if TT_FACE_REBASE
        // (B, P) = the polynomial at the start of the step and its change over the step's w-interval [fw, fw + h]
        if (fw != 0.f) { B = face_at(B, P, bc2(fw)); P = face_scale(P, bc2(1.f - fw)); }       // launched inside a cell
        while (true) {
            // face k+2, consumed when the ray arrives at plane k+1.  Unconditional (a predicated load goes through
            // temporaries and 12 moves): the grid carries one spare plane behind the last face
            N = face_ldw(p + 2 * fplane);
            const float q = trcp<float>(dw);        // (dw > TT_MARCH_MIN_DW: checked at the entry and after every step)
            const f32x2 aUV = mul2(duv, bc2(q));
            float h = 1.f - fw;
            int cross = 0;
            {
                const f32x2 puv = fma2(bc2(h), aUV, tuv);
                const float pu = lo2(puv), pv = hi2(puv);
                if (fabsf(pu) > 0.5f || fabsf(pv) > 0.5f) {
                    const float aU = lo2(aUV), aV = hi2(aUV), tu = lo2(tuv), tv = hi2(tuv);
                    float lu = 2.f, lv = 2.f;
                    if (aU > 0.f) lu = chord_fraction<float>(0.5f - tu, h * aU); else if (aU < 0.f) lu = chord_fraction<float>(-0.5f - tu, h * aU);
                    if (aV > 0.f) lv = chord_fraction<float>(0.5f - tv, h * aV); else if (aV < 0.f) lv = chord_fraction<float>(-0.5f - tv, h * aV);
                    float lam = fminf(lu, lv);
                    if (lam < 1.f) {
                        cross = lu <= lv ? (aU > 0.f ? 1 : -1) : (aV > 0.f ? 2 : -2);
                        lam = lam > 0.f ? lam : 0.f;
                        h *= lam;
                        P = face_scale(P, bc2(lam));                          // the step stops at the side face
                    }
                }
            }
            f32x2 g; float gw;
            face_eval(B, tuv, g, gw);
            const bool ok = face_step<true, TRACK_S>(B, P, tuv, duv, dw, s, 0.f, h, q, aUV, mul2(g, bc2(q)), gw * q);
            if (!(ok && dw > (float)TT_MARCH_MIN_DW)) { fast = false; break; }
            if (cross == 0) {
                ++steps;
                fw = 0.f;
                if (++k >= nw - 1) break;
                p += fplane;
                P = face_sub(face_pack(N), B);                                // B is the far face by now: the next base
            } else {
                fw += h;
                float tu = lo2(tuv), tv = hi2(tuv);
                int dp = 0;
                if (cross == 1) { ++cu; tu -= 1.f; dp = 3; } else if (cross == -1) { --cu; tu += 1.f; dp = -3; }
                else if (cross == 2) { ++cv; tv -= 1.f; dp = row3; } else { --cv; tv += 1.f; dp = -row3; }
                p += dp;
                tuv = pk2(tu, tv);
                if (cu < 0 || cu > nu - 2 || cv < 0 || cv > nv - 2) { fast = false; break; }      // side exit
                const FaceQ Bn = face_ld(p);                                  // both faces of the new cell column
                const FaceQ Pn = face_sub(face_ld(p + fplane), Bn);
                B = face_at(Bn, Pn, bc2(fw));                                 // rebased to the rest of the cell
                P = face_scale(Pn, bc2(1.f - fw));
            }
        }
#else
        while (true) {
            // face k+2, consumed at the end of the step.  Unconditional (a predicated load goes through temporaries and
            // 12 moves): the grid carries one spare plane behind the last face for the load issued in the last cell
            N = face_ldw(p + 2 * fplane);
#if TT_FACE_PREFETCH && defined(__CUDA_ARCH__)
            if (k + 2 + TT_FACE_PREFETCH <= nw) {
                prefetch_l2(p + (2 + TT_FACE_PREFETCH) * fplane);
                prefetch_l2(p + (2 + TT_FACE_PREFETCH) * fplane + 2);
            }
#endif
            // ---- stage 1 and the length of this step -------------------------------------------------------
            const float q = trcp<float>(dw);        // (dw > TT_MARCH_MIN_DW: checked at the entry and after every step)
            bool ok = true;
            const f32x2 aUV = mul2(duv, bc2(q));
            float h = 1.f - fw;
            int cross = 0;
            {
                const f32x2 puv = fma2(bc2(h), aUV, tuv);
                const float pu = lo2(puv), pv = hi2(puv);
                if (fabsf(pu) > 0.5f || fabsf(pv) > 0.5f) {
                    const float aU = lo2(aUV), aV = hi2(aUV), tu = lo2(tuv), tv = hi2(tuv);
                    float lu = 2.f, lv = 2.f;
                    if (aU > 0.f) lu = chord_fraction<float>(0.5f - tu, h * aU); else if (aU < 0.f) lu = chord_fraction<float>(-0.5f - tu, h * aU);
                    if (aV > 0.f) lv = chord_fraction<float>(0.5f - tv, h * aV); else if (aV < 0.f) lv = chord_fraction<float>(-0.5f - tv, h * aV);
                    const float lam = fminf(lu, lv);
                    if (lam < 1.f) {
                        cross = lu <= lv ? (aU > 0.f ? 1 : -1) : (aV > 0.f ? 2 : -2);
                        h *= lam > 0.f ? lam : 0.f;
                    }
                }
            }
            f32x2 g; float gw;
            bool full = false;
#if TT_FACE_FASTPATH && defined(__CUDA_ARCH__)
            full = __all_sync(__activemask(), fw == 0.f && cross == 0);
#elif TT_FACE_FASTPATH
            full = fw == 0.f && cross == 0;
#endif
            if (full) {
                face_eval(B, tuv, g, gw);
                ok = face_step<true, TRACK_S>(B, P, tuv, duv, dw, s, 0.f, h, q, aUV, mul2(g, bc2(q)), gw * q) && ok;
                if (!(ok && dw > (float)TT_MARCH_MIN_DW)) { fast = false; break; }
                ++steps;
                if (++k >= nw - 1) break;
                p += fplane;
                P = face_sub(face_pack(N), B);                                // B is the far face by now: the next base
                continue;
            }
            face_eval(face_at(B, P, bc2(fw)), tuv, g, gw);
            ok = face_step<false, TRACK_S>(B, P, tuv, duv, dw, s, fw, h, q, aUV, mul2(g, bc2(q)), gw * q) && ok;
            if (!(ok && dw > (float)TT_MARCH_MIN_DW)) { fast = false; break; }
            if (cross == 0) {
                ++steps;
                fw = 0.f;
                if (++k >= nw - 1) break;
                p += fplane;
                B = face_add(B, P);                                           // plane k+1 becomes the base face
            } else {
                fw += h;
                float tu = lo2(tuv), tv = hi2(tuv);
                int dp = 0;
                if (cross == 1) { ++cu; tu -= 1.f; dp = 3; } else if (cross == -1) { --cu; tu += 1.f; dp = -3; }
                else if (cross == 2) { ++cv; tv -= 1.f; dp = row3; } else { --cv; tv += 1.f; dp = -row3; }
                p += dp;
                tuv = pk2(tu, tv);
                if (cu < 0 || cu > nu - 2 || cv < 0 || cv > nv - 2) { fast = false; break; }      // side exit
                B = face_ld(p);                                               // both faces of the new cell column
                N = face_ldw(p + fplane);
            }
            P = face_sub(face_pack(N), B);
        }
#endif
    }
    if (!fast) {
        status[ray] = TT_RAY_DEFERRED;
        deferred = true;
        steps = 0;
    } else {
        const double Pu = A.o[0] + (((double)cu + 0.5) + (double)lo2(tuv)) * A.h[0];
        const double Pv = A.o[1] + (((double)cv + 0.5) + (double)hi2(tuv)) * A.h[1];
        const double Pw = A.o[2] + (double)(nw - 1) * A.h[2];
        const double Vu = (double)lo2(duv) * FA.inv_ru * kC, Vv = (double)hi2(duv) * FA.inv_rv * kC, Vw = (double)dw * kC;
        const double tb = (Pw - A.extent) / Vw;
        rf[0 * A.np + ray] = Pu - Vu * tb;
        rf[1 * A.np + ray] = atan(Vu / Vw);
        rf[2 * A.np + ray] = Pv - Vv * tb;
        rf[3 * A.np + ray] = atan(Vv / Vw);
        if (track_s) {
            const double t_rest = (A.s_max - s_pre - (double)s * A.h[2]) / kC;
            const double Pf[3] = {Pu, Pv, Pw}, Vf[3] = {Vu, Vv, Vw};
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                sf[(size_t)A.fa[m] * A.np + ray] = Pf[m] + Vf[m] * t_rest;
                sf[(size_t)(3 + A.fa[m]) * A.np + ray] = Vf[m];
            }
        }
        status[ray] = (uint8_t)TT_RAY_EXIT_FACE;
    }
    return steps;
}

}  // namespace tt
