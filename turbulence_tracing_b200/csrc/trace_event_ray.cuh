// Event marching on uniform grids: the per-ray body of the scalar kernels (tt_trace variant 4, and variant 3 in
// FP64), compiled for the device (trace_event.cu) AND for the host (tests/host/trace_event_host.cu runs the very
// same source on the CPU against the C oracle).  The packed FP32x2 kernel of trace_event.cu performs the same
// operations in the same order (bit-identical rays, tested on the GPU), written with inline PTX.
#pragma once
#pragma nv_diag_suppress 550   // lo2()/hi2() unpack a register pair through asm and use one half each
#include <string.h>
#include "trace_common.cuh"

namespace tt {

#ifndef TT_EVENT_PREFETCH
#define TT_EVENT_PREFETCH 0        // planes ahead whose corner rows are prefetched into L1 (prefetch.global.L1).
                                   // Measured on B200 (513^3, 1e8 rays): 0 -> 470.7 ms, 2/4/8 -> 485-487 ms: the
                                   // register prefetch one plane ahead already covers L1 hits; off by default.
#endif
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// (Measured and removed in round 2: issuing the prefetch of plane k+2 where the cell changes instead of behind the `have_next`
// flag at the loop top -- same loads, bit-identical rays, a few control instructions less: 391.5 vs 383.0 ms on 513^3 / 1e8
// rays through tt_trace, 62.1 vs 62.0 ms on configs[3]; profiles/r02_ab_lean_axesrcp.txt.)
#ifndef TT_EVENT_MERGE
#define TT_EVENT_MERGE 1           // a cell change (next plane OR a u / v face) renews the polynomial in two halves: the base
                                   // plane (advance: base += primed; face: 4 corners of the new column) and, in code COMMON
                                   // to both, the primed half from the 4 corners of the far plane.  Same operations on the
                                   // same operands as tri_set / tri_advance (bit-identical rays), but the part of the
                                   // relabelling that a warp executes for one or two lanes only shrinks by ~50 instructions
#endif
#ifndef TT_EVENT_FASTDIV
#define TT_EVENT_FASTDIV 1         // chord fractions of a predicted face crossing with MUFU.RCP instead of IEEE divisions
                                   // (they only decide where the step is split; the ray keeps its true position)
#endif
// chord fraction num / den of a predicted face crossing (den != 0); NaN (0 * inf) and overflow -> 2 = "not reached"
template <typename T>
TT_HD T chord_fraction(T num, T den) {
#if TT_EVENT_FASTDIV
    const T l = num * trcp<T>(den);
    return l < T(2) ? l : T(2);                 // false for NaN
#else
    return num / den;
#endif
}

// Returns the (sub-)plane arrivals of this ray (0 if it is deferred to the general kernel).
template <typename T, bool SPC1>
TT_HD unsigned event_ray(const typename GridT<T>::V4* __restrict__ grid, const double* __restrict__ s0, long ray,
                         double* __restrict__ rf, double* __restrict__ sf, uint8_t* __restrict__ status,
                         const TraceArgs& A, bool& deferred) {
    typedef typename GridT<T>::V4 V4;
    unsigned steps = 0;
    const int nu = A.n[0], nv = A.n[1], nw = A.n[2];
    const long long plane = A.plane_elems;
    // ---- prologue ---------------------------------------------------------------------------
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
    // the path-time cap must be out of reach while marching (d_w > 0.75 throughout)
    fast = fast && ((double)(nw - 1) - X[2]) * A.h[2] <= TT_MARCH_MIN_DW * (A.s_max - s_pre);
    int cu = 0, cv = 0, k = 0;
    T tu = T(0), tv = T(0), fw = T(0);
    if (fast) {
        double fl;
        fl = fmin(floor(X[0]), (double)(nu - 2)); cu = (int)fl; tu = (T)(X[0] - fl);
        fl = fmin(floor(X[1]), (double)(nv - 2)); cv = (int)fl; tv = (T)(X[1] - fl);
        fl = floor(X[2]); k = (int)fl; fw = (T)(X[2] - fl);
    }
    T du = (T)D[0], dv = (T)D[1], dw = (T)D[2], s = T(0);
    const T hw = (T)A.h[2], ru = (T)(A.h[2] / A.h[0]), rv = (T)(A.h[2] / A.h[1]);
    const bool track_s = sf != nullptr;
    const int spc = A.spc;
    const T hsub = SPC1 ? T(1) : T(1) / (T)spc;
    int j = SPC1 ? 0 : (int)(fw * (T)spc);       // current sub-plane interval of the w-cell

    if (fast && k < nw - 1) {
        const V4* p = grid + ((size_t)k * plane + (size_t)cv * nu + cu);
        Tri<T> qx, qy, qz;
        V4 n00, n10, n01, n11;                    // corners of plane k+2 (prefetch)
        {
            V4 c00 = GridT<T>::ld3(p), c10 = GridT<T>::ld3(p + 1), c01 = GridT<T>::ld3(p + nu), c11 = GridT<T>::ld3(p + nu + 1);
            const V4* p1 = p + plane;
            V4 e00 = GridT<T>::ld3(p1), e10 = GridT<T>::ld3(p1 + 1), e01 = GridT<T>::ld3(p1 + nu), e11 = GridT<T>::ld3(p1 + nu + 1);
            tri_set<T>(qx, c00.x, c10.x, c01.x, c11.x, e00.x, e10.x, e01.x, e11.x);
            tri_set<T>(qy, c00.y, c10.y, c01.y, c11.y, e00.y, e10.y, e01.y, e11.y);
            tri_set<T>(qz, c00.z, c10.z, c01.z, c11.z, e00.z, e10.z, e01.z, e11.z);
        }
        bool have_next = false;
        while (true) {
            if (!have_next && k + 2 <= nw - 1) {
                const V4* p2 = p + 2 * plane;
                n00 = GridT<T>::ld3(p2); n10 = GridT<T>::ld3(p2 + 1); n01 = GridT<T>::ld3(p2 + nu); n11 = GridT<T>::ld3(p2 + nu + 1);
                have_next = true;
            }
            // ---- stage 1 and the length of this step -------------------------------------------
            T q = trcp<T>(dw), hq = hw * q;
            bool ok = dw > T(TT_MARCH_MIN_DW);
            const T aU = ru * du * q, aV = rv * dv * q;
            const T adu = bil_eval<T>(tri_at<T>(qx, fw), tu, tv) * hq, adv = bil_eval<T>(tri_at<T>(qy, fw), tu, tv) * hq,
                    adw = bil_eval<T>(tri_at<T>(qz, fw), tu, tv) * hq, as = hq;
            const T fw_t = SPC1 ? T(1) : ((j + 1 == spc) ? T(1) : (T)(j + 1) * hsub);
            T h = fw_t - fw;
            int cross = 0;                         // +-1: u face, +-2: v face
            {
                const T pu = tfma(h, aU, tu), pv = tfma(h, aV, tv);
                if (pu > T(1) || pu < T(0) || pv > T(1) || pv < T(0)) {
                    // fraction of the remaining interval at which the chord reaches the face the ray
                    // is heading for (a zero slope never reaches a face)
                    T lu = T(2), lv = T(2);
                    if (aU > T(0)) lu = chord_fraction<T>(T(1) - tu, h * aU); else if (aU < T(0)) lu = chord_fraction<T>(-tu, h * aU);
                    if (aV > T(0)) lv = chord_fraction<T>(T(1) - tv, h * aV); else if (aV < T(0)) lv = chord_fraction<T>(-tv, h * aV);
                    T lam = fmin(lu, lv);
                    if (lam < T(1)) {
                        cross = lu <= lv ? (aU > T(0) ? 1 : -1) : (aV > T(0) ? 2 : -2);
                        h *= lam > T(0) ? lam : T(0);
                    }
                }
            }
            const T half = T(0.5) * h;
            // ---- stages 2-4 ---------------------------------------------------------------------
            T su = tfma(half, aU, tu), sv = tfma(half, aV, tv), sw = fw + half;
            T du2 = tfma(half, adu, du), dv2 = tfma(half, adv, dv), dw2 = tfma(half, adw, dw);
            q = trcp<T>(dw2); hq = hw * q; ok = ok && dw2 > T(0);
            const T bU = ru * du2 * q, bV = rv * dv2 * q;
            const Bil<T> mx = tri_at<T>(qx, sw), my = tri_at<T>(qy, sw), mz = tri_at<T>(qz, sw);   // stages 2 and 3
            const T bdu = bil_eval<T>(mx, su, sv) * hq, bdv = bil_eval<T>(my, su, sv) * hq,
                    bdw = bil_eval<T>(mz, su, sv) * hq, bs = hq;
            su = tfma(half, bU, tu); sv = tfma(half, bV, tv);
            du2 = tfma(half, bdu, du); dv2 = tfma(half, bdv, dv); dw2 = tfma(half, bdw, dw);
            q = trcp<T>(dw2); hq = hw * q; ok = ok && dw2 > T(0);
            const T cU = ru * du2 * q, cV = rv * dv2 * q;
            const T cdu = bil_eval<T>(mx, su, sv) * hq, cdv = bil_eval<T>(my, su, sv) * hq,
                    cdw = bil_eval<T>(mz, su, sv) * hq, cs = hq;
            su = tfma(h, cU, tu); sv = tfma(h, cV, tv); sw = fw + h;
            du2 = tfma(h, cdu, du); dv2 = tfma(h, cdv, dv); dw2 = tfma(h, cdw, dw);
            q = trcp<T>(dw2); hq = hw * q; ok = ok && dw2 > T(0);
            const T eU = ru * du2 * q, eV = rv * dv2 * q;
            const T edu = bil_eval<T>(tri_at<T>(qx, sw), su, sv) * hq, edv = bil_eval<T>(tri_at<T>(qy, sw), su, sv) * hq,
                    edw = bil_eval<T>(tri_at<T>(qz, sw), su, sv) * hq, es = hq;
            const T h6 = h * T(1.0 / 6.0);
            tu = tfma(h6, aU + T(2) * (bU + cU) + eU, tu);
            tv = tfma(h6, aV + T(2) * (bV + cV) + eV, tv);
            du = tfma(h6, adu + T(2) * (bdu + cdu) + edu, du);
            dv = tfma(h6, adv + T(2) * (bdv + cdv) + edv, dv);
            dw = tfma(h6, adw + T(2) * (bdw + cdw) + edw, dw);
            if (track_s) s = tfma(h6, as + T(2) * (bs + cs) + es, s);
            if (!(ok && dw > T(TT_MARCH_MIN_DW))) { fast = false; break; }   // steep / turning / NaN
#if TT_EVENT_MERGE
            bool renew = true;
            if (cross == 0) {
                // ---- reached the next (sub-)plane: plane k+1 becomes the base plane -----------
                ++steps;
                fw = fw_t;
                if (SPC1 || ++j == spc) {
                    j = 0; fw = T(0);
                    if (++k >= nw - 1) break;                                     // far face: done
                    p += plane;
                    tri_base_step<T>(qx); tri_base_step<T>(qy); tri_base_step<T>(qz);
                } else {
                    renew = false;
                }
            } else {
                // ---- reached a u / v cell face inside the w-cell: relabel, base plane of the new column ----
                fw += h;
                if (cross == 1) { ++cu; tu -= T(1); p += 1; } else if (cross == -1) { --cu; tu += T(1); p -= 1; }
                else if (cross == 2) { ++cv; tv -= T(1); p += nu; } else { --cv; tv += T(1); p -= nu; }
                if (cu < 0 || cu > nu - 2 || cv < 0 || cv > nv - 2) { fast = false; break; }   // side exit
                const V4 c00 = GridT<T>::ld3(p), c10 = GridT<T>::ld3(p + 1), c01 = GridT<T>::ld3(p + nu), c11 = GridT<T>::ld3(p + nu + 1);
                tri_base<T>(qx, c00.x, c10.x, c01.x, c11.x);
                tri_base<T>(qy, c00.y, c10.y, c01.y, c11.y);
                tri_base<T>(qz, c00.z, c10.z, c01.z, c11.z);
                const V4* p1 = p + plane;                                         // its far plane, into the prefetch registers
                n00 = GridT<T>::ld3(p1); n10 = GridT<T>::ld3(p1 + 1); n01 = GridT<T>::ld3(p1 + nu); n11 = GridT<T>::ld3(p1 + nu + 1);
            }
            if (SPC1 || renew) {                  // common to both: the primed half from the 4 corners of the far plane
                tri_primed<T>(qx, n00.x, n10.x, n01.x, n11.x);
                tri_primed<T>(qy, n00.y, n10.y, n01.y, n11.y);
                tri_primed<T>(qz, n00.z, n10.z, n01.z, n11.z);
                have_next = false;
            }
#else
            if (cross == 0) {
                // ---- reached the next (sub-)plane ---------------------------------------------
                ++steps;
                fw = fw_t;
                if (SPC1 || ++j == spc) {
                    j = 0; fw = T(0);
                    if (++k >= nw - 1) break;                                     // far face: done
                    p += plane;
#if TT_EVENT_PREFETCH && defined(__CUDA_ARCH__)
                    if (k + TT_EVENT_PREFETCH <= nw - 1) {                        // register-free L1 prefetch
                        prefetch_l1(p + TT_EVENT_PREFETCH * plane);
                        prefetch_l1(p + TT_EVENT_PREFETCH * plane + nu);
                    }
#endif
                    if (k + 1 <= nw - 1) {
                        tri_advance<T>(qx, n00.x, n10.x, n01.x, n11.x);
                        tri_advance<T>(qy, n00.y, n10.y, n01.y, n11.y);
                        tri_advance<T>(qz, n00.z, n10.z, n01.z, n11.z);
                    }
                    have_next = false;
                }
            } else {
                // ---- reached a u / v cell face inside the w-cell: relabel and reload ---------
                fw += h;
                if (cross == 1) { ++cu; tu -= T(1); p += 1; } else if (cross == -1) { --cu; tu += T(1); p -= 1; }
                else if (cross == 2) { ++cv; tv -= T(1); p += nu; } else { --cv; tv += T(1); p -= nu; }
                if (cu < 0 || cu > nu - 2 || cv < 0 || cv > nv - 2) { fast = false; break; }   // side exit
                V4 c00 = GridT<T>::ld3(p), c10 = GridT<T>::ld3(p + 1), c01 = GridT<T>::ld3(p + nu), c11 = GridT<T>::ld3(p + nu + 1);
                const V4* p1 = p + plane;
                V4 e00 = GridT<T>::ld3(p1), e10 = GridT<T>::ld3(p1 + 1), e01 = GridT<T>::ld3(p1 + nu), e11 = GridT<T>::ld3(p1 + nu + 1);
                tri_set<T>(qx, c00.x, c10.x, c01.x, c11.x, e00.x, e10.x, e01.x, e11.x);
                tri_set<T>(qy, c00.y, c10.y, c01.y, c11.y, e00.y, e10.y, e01.y, e11.y);
                tri_set<T>(qz, c00.z, c10.z, c01.z, c11.z, e00.z, e10.z, e01.z, e11.z);
                have_next = false;
            }
#endif
        }
    }
    if (!fast) {
        status[ray] = TT_RAY_DEFERRED;          // the general kernel redoes this ray from s0
        deferred = true;
        steps = 0;
    } else {
        // ---- epilogue: ray_at_exit (particle_tracker.py:345-380) and state at time T ---------
        const double Pu = A.o[0] + ((double)cu + (double)tu) * A.h[0];
        const double Pv = A.o[1] + ((double)cv + (double)tv) * A.h[1];
        const double Pw = A.o[2] + (double)(nw - 1) * A.h[2];
        const double Vu = (double)du * kC, Vv = (double)dv * kC, Vw = (double)dw * kC;
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
        status[ray] = (uint8_t)TT_RAY_EXIT_FACE;
    }
    return steps;
}

// ---- the same body with Blackwell's packed FP32x2 arithmetic (tt_trace variant 3 in FP32: the production kernel) ----
typedef unsigned long long f32x2;
#ifdef __CUDA_ARCH__
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo2(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)b; return a; }
__device__ __forceinline__ float hi2(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)a; return b; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
#else
// host emulation of the packed instructions, lane by lane with the same IEEE operations (round to nearest, fused
// multiply-add): the host run of event_ray_f32x2 reproduces the device's arithmetic except for MUFU.RCP
inline f32x2 pk2(float lo, float hi) { unsigned int a, b; memcpy(&a, &lo, 4); memcpy(&b, &hi, 4); return ((f32x2)b << 32) | a; }
inline float lo2(f32x2 v) { unsigned int a = (unsigned int)v; float f; memcpy(&f, &a, 4); return f; }
inline float hi2(f32x2 v) { unsigned int a = (unsigned int)(v >> 32); float f; memcpy(&f, &a, 4); return f; }
inline f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { return pk2(fmaf(lo2(a), lo2(b), lo2(c)), fmaf(hi2(a), hi2(b), hi2(c))); }
inline f32x2 mul2(f32x2 a, f32x2 b) { volatile float l = lo2(a) * lo2(b), h = hi2(a) * hi2(b); return pk2(l, h); }
inline f32x2 add2(f32x2 a, f32x2 b) { volatile float l = lo2(a) + lo2(b), h = hi2(a) + hi2(b); return pk2(l, h); }
inline f32x2 sub2(f32x2 a, f32x2 b) { volatile float l = lo2(a) - lo2(b), h = hi2(a) - hi2(b); return pk2(l, h); }
#endif
TT_HD f32x2 bc2(float s) { return pk2(s, s); }

struct Tri2 {          // two trilinear polynomials (one per lane) of one cell
    f32x2 a, b, c, d, a1, b1, c1, d1;
};
struct Bil2 {
    f32x2 a, b, c, d;
};
TT_HD Bil2 tri2_at(const Tri2& q, f32x2 FW) {
    Bil2 r;
    r.a = fma2(FW, q.a1, q.a); r.b = fma2(FW, q.b1, q.b); r.c = fma2(FW, q.c1, q.c); r.d = fma2(FW, q.d1, q.d);
    return r;
}
TT_HD f32x2 bil2_eval(const Bil2& q, f32x2 TU, f32x2 TV) {
    return fma2(TV, fma2(TU, q.d, q.c), fma2(TU, q.b, q.a));
}
TT_HD void tri2_set(Tri2& q, f32x2 c00, f32x2 c10, f32x2 c01, f32x2 c11, f32x2 e00, f32x2 e10,
                                         f32x2 e01, f32x2 e11) {
    q.a = c00; q.b = sub2(c10, c00); q.c = sub2(c01, c00); q.d = sub2(sub2(c11, c01), q.b);
    f32x2 eb = sub2(e10, e00), ec = sub2(e01, e00), ed = sub2(sub2(e11, e01), eb);
    q.a1 = sub2(e00, q.a); q.b1 = sub2(eb, q.b); q.c1 = sub2(ec, q.c); q.d1 = sub2(ed, q.d);
}
TT_HD void tri2_advance(Tri2& q, f32x2 n00, f32x2 n10, f32x2 n01, f32x2 n11) {
    q.a = add2(q.a, q.a1); q.b = add2(q.b, q.b1); q.c = add2(q.c, q.c1); q.d = add2(q.d, q.d1);
    f32x2 eb = sub2(n10, n00);
    q.a1 = sub2(n00, q.a); q.b1 = sub2(eb, q.b); q.c1 = sub2(sub2(n01, n00), q.c);
    q.d1 = sub2(sub2(sub2(n11, n01), eb), q.d);
}
// the two halves of tri2_set / tri2_advance (TT_EVENT_MERGE): same operations on the same operands
TT_HD void tri2_base(Tri2& q, f32x2 c00, f32x2 c10, f32x2 c01, f32x2 c11) {
    q.a = c00; q.b = sub2(c10, c00); q.c = sub2(c01, c00); q.d = sub2(sub2(c11, c01), q.b);
}
TT_HD void tri2_base_step(Tri2& q) {
    q.a = add2(q.a, q.a1); q.b = add2(q.b, q.b1); q.c = add2(q.c, q.c1); q.d = add2(q.d, q.d1);
}
TT_HD void tri2_primed(Tri2& q, f32x2 n00, f32x2 n10, f32x2 n01, f32x2 n11) {
    f32x2 eb = sub2(n10, n00);
    q.a1 = sub2(n00, q.a); q.b1 = sub2(eb, q.b); q.c1 = sub2(sub2(n01, n00), q.c);
    q.d1 = sub2(sub2(sub2(n11, n01), eb), q.d);
}
// ONE trilinear polynomial (the scalar g_w lane) evaluated with packed operations: its bilinear coefficients ride as the
// pairs (a, c) and (b, d), so that  a + tu b  and  c + tu d  are one FFMA2.  Per lane these are the operations of
// Tri<float> / tri_at / bil_eval in the same order (bit-identical); 7 scalar FFMA per evaluation become 3 FFMA2 + 1 FFMA,
// and on sm_100 an FFMA costs almost as much operand bandwidth as an FFMA2 (scripts/ubench_fp32_pipe.cu: 1.7 vs 2.9
// cycles per sub-partition with three register operands)
#ifndef TT_EVENT_PACK_W
#define TT_EVENT_PACK_W 1
#endif
struct TriP {
    f32x2 ac, bd, ac1, bd1;
};
struct BilP {
    f32x2 ac, bd;
};
TT_HD BilP trip_at(const TriP& q, f32x2 FW) {
    BilP r;
    r.ac = fma2(FW, q.ac1, q.ac); r.bd = fma2(FW, q.bd1, q.bd);
    return r;
}
TT_HD float bilp_eval(const BilP& q, float tu, float tv) {
    const f32x2 r = fma2(bc2(tu), q.bd, q.ac);          // (a + tu b, c + tu d)
    return fmaf(tv, hi2(r), lo2(r));
}
TT_HD void trip_base(TriP& q, float c00, float c10, float c01, float c11) {
    const f32x2 bt = sub2(pk2(c10, c11), pk2(c00, c01));          // (b, c11 - c01)
    q.ac = pk2(c00, c01 - c00);
    q.bd = pk2(lo2(bt), hi2(bt) - lo2(bt));
}
TT_HD void trip_base_step(TriP& q) {
    q.ac = add2(q.ac, q.ac1); q.bd = add2(q.bd, q.bd1);
}
TT_HD void trip_primed(TriP& q, float n00, float n10, float n01, float n11) {
    const f32x2 e = sub2(pk2(n10, n11), pk2(n00, n01));           // (eb, n11 - n01)
    q.ac1 = sub2(pk2(n00, n01 - n00), q.ac);
    q.bd1 = sub2(pk2(lo2(e), hi2(e) - lo2(e)), q.bd);
}
#define TT_XY(v) pk2((v).x, (v).y)
#define TT_ZW(v) pk2((v).z, (v).w)

// AUX = true additionally carries the passive quantities of tt_trace_aux (phase, Faraday rotation,
// absorption): the ne/nc lane rides with g_w as a packed pair, a second grid (B_u, B_v | B_w, kappa) gets
// two more packed polynomials, and the RK4 stages double as Simpson nodes of the three line integrals.
TT_HD void aux_integrands(float nn, f32x2 bxy, f32x2 bzk, f32x2 duv, float dw, float hq, bool has_b,
                                               float& fp, float& ff, float& fa) {
    const float r = tsqrt01(fmaxf(1.f - nn, 1e-30f));             // (MUFU.RSQ / MUFU.RCP: the IEEE square root and division
    fp = -nn * trcp<float>(1.f + r) * hq;                          //  were ~20 instructions per stage)  (sqrt(1 - ne/nc) - 1) ds
    ff = 0.f; fa = 0.f;
    if (has_b) {
        const float bd = fmaf(lo2(bxy), lo2(duv), fmaf(hi2(bxy), hi2(duv), lo2(bzk) * dw));
        ff = nn * bd * hq;                          // (ne/nc) (B . d) ds
        fa = hi2(bzk) * hq;                         // kappa ds
    }
}

// Returns the (sub-)plane arrivals of this ray (0 if it is deferred to the general kernel).
// TRACK_S = false: the caller passes sf == nullptr (no state at time T wanted), so the path time needs no bookkeeping
template <bool SPC1, bool AUX, bool CUBIC, bool TRACK_S = true>
TT_HD unsigned event_ray_f32x2(const float4* __restrict__ grid, const double* __restrict__ s0, long ray,
                               double* __restrict__ rf, double* __restrict__ sf, uint8_t* __restrict__ status,
                               const TraceArgs& A, const float4* __restrict__ aux4, double* __restrict__ aux_out,
                               const AuxArgs& AX, bool& deferred) {
    typedef float T;
#ifndef TT_EVENT_LD3
#define TT_EVENT_LD3 1
#endif
#define TT_LDN(q) ((AUX || !TT_EVENT_LD3) ? GridT<float>::ld(q) : GridT<float>::ld3(q))     // 4th lane (ne/nc) only when it is used
    unsigned steps = 0;
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
    const bool track_s = TRACK_S && sf != nullptr;
    const int spc = A.spc;
    const T hsub = SPC1 ? 1.f : 1.f / (T)spc;
    int j = SPC1 ? 0 : (int)(fw * (T)spc);
    double acc_p = 0.0, acc_f = 0.0, acc_a = 0.0;     // AUX: line integrals of (n-1), (ne/nc)(B.d), kappa

    if (fast && k < nw - 1) {
        const float4* p = grid + ((size_t)k * plane + (size_t)cv * nu + cu);
        Tri2 qxy;                 // (g_u, g_v) lanes, packed
#if TT_EVENT_PACK_W
        TriP qz;                  // g_w: one polynomial, its coefficient pairs (a, c), (b, d) packed
#define TT_QZ_AT(fwv) trip_at(qz, bc2(fwv))
#define TT_QZ_EVAL(bil, tuv_) bilp_eval(bil, lo2(tuv_), hi2(tuv_))
#define TT_QZ_BASE(c00, c10, c01, c11) trip_base(qz, c00, c10, c01, c11)
#define TT_QZ_STEP() trip_base_step(qz)
#define TT_QZ_PRIMED(n00, n10, n01, n11) trip_primed(qz, n00, n10, n01, n11)
        typedef BilP BilZ;
#else
        Tri<float> qz;            // g_w, scalar
#define TT_QZ_AT(fwv) tri_at<float>(qz, fwv)
#define TT_QZ_EVAL(bil, tuv_) bil_eval<float>(bil, lo2(tuv_), hi2(tuv_))
#define TT_QZ_BASE(c00, c10, c01, c11) tri_base<float>(qz, c00, c10, c01, c11)
#define TT_QZ_STEP() tri_base_step<float>(qz)
#define TT_QZ_PRIMED(n00, n10, n01, n11) tri_primed<float>(qz, n00, n10, n01, n11)
        typedef Bil<float> BilZ;
#endif
        Tri2 qzw, bxy, bzk;       // AUX: (g_w, ne/nc), (B_u, B_v), (B_w, kappa)
        const bool has_b = AUX && aux4 != nullptr;
        const float4* pa = has_b ? aux4 + (p - grid) : nullptr;
        float4 n00, n10, n01, n11;
        // (re)build the polynomials of the current cell from planes k and k+1
        auto load_cell = [&]() {
            float4 c00 = TT_LDN(p), c10 = TT_LDN(p + 1), c01 = TT_LDN(p + nu), c11 = TT_LDN(p + nu + 1);
            const float4* p1 = p + plane;
            float4 e00 = TT_LDN(p1), e10 = TT_LDN(p1 + 1), e01 = TT_LDN(p1 + nu), e11 = TT_LDN(p1 + nu + 1);
            tri2_set(qxy, TT_XY(c00), TT_XY(c10), TT_XY(c01), TT_XY(c11), TT_XY(e00), TT_XY(e10), TT_XY(e01), TT_XY(e11));
            if (AUX) tri2_set(qzw, TT_ZW(c00), TT_ZW(c10), TT_ZW(c01), TT_ZW(c11), TT_ZW(e00), TT_ZW(e10), TT_ZW(e01), TT_ZW(e11));
            else { TT_QZ_BASE(c00.z, c10.z, c01.z, c11.z); TT_QZ_PRIMED(e00.z, e10.z, e01.z, e11.z); }
            if (has_b) {
                c00 = GridT<float>::ld(pa); c10 = GridT<float>::ld(pa + 1); c01 = GridT<float>::ld(pa + nu); c11 = GridT<float>::ld(pa + nu + 1);
                const float4* q1 = pa + plane;
                e00 = GridT<float>::ld(q1); e10 = GridT<float>::ld(q1 + 1); e01 = GridT<float>::ld(q1 + nu); e11 = GridT<float>::ld(q1 + nu + 1);
                tri2_set(bxy, TT_XY(c00), TT_XY(c10), TT_XY(c01), TT_XY(c11), TT_XY(e00), TT_XY(e10), TT_XY(e01), TT_XY(e11));
                tri2_set(bzk, TT_ZW(c00), TT_ZW(c10), TT_ZW(c01), TT_ZW(c11), TT_ZW(e00), TT_ZW(e10), TT_ZW(e01), TT_ZW(e11));
            }
        };
        load_cell();
        bool have_next = false;
        while (true) {
            if (!have_next && k + 2 <= nw - 1) {
                const float4* p2 = p + 2 * plane;
                n00 = TT_LDN(p2); n10 = TT_LDN(p2 + 1); n01 = TT_LDN(p2 + nu); n11 = TT_LDN(p2 + nu + 1);
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
                adw = TT_QZ_EVAL(TT_QZ_AT(fw), tuv) * hq;
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
                    if (aU > 0.f) lu = chord_fraction<T>(1.f - tu, h * aU); else if (aU < 0.f) lu = chord_fraction<T>(-tu, h * aU);
                    if (aV > 0.f) lv = chord_fraction<T>(1.f - tv, h * aV); else if (aV < 0.f) lv = chord_fraction<T>(-tv, h * aV);
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
            BilZ mz;
            Bil2 mzw, mb1, mb2;
            if (AUX) {
                mzw = tri2_at(qzw, bc2(sw));
                if (has_b) { mb1 = tri2_at(bxy, bc2(sw)); mb2 = tri2_at(bzk, bc2(sw)); }
            } else {
                mz = TT_QZ_AT(sw);
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
                bdw = TT_QZ_EVAL(mz, suv) * hq;
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
                cdw = TT_QZ_EVAL(mz, suv) * hq;
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
                edw = TT_QZ_EVAL(TT_QZ_AT(sw), suv) * hq;
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
#if TT_EVENT_MERGE
            bool renew = true;
            if (cross == 0) {
                ++steps;
                fw = fw_t;
                if (SPC1 || ++j == spc) {
                    j = 0; fw = 0.f;
                    if (++k >= nw - 1) break;
                    p += plane;
                    tri2_base_step(qxy);                              // plane k+1 becomes the base plane
                    if (AUX) tri2_base_step(qzw); else TT_QZ_STEP();
                    if (has_b) { pa += plane; tri2_base_step(bxy); tri2_base_step(bzk); }
                } else {
                    renew = false;
                }
            } else {
                fw += h;
                T tu = lo2(tuv), tv = hi2(tuv);
                int dp = 0;
                if (cross == 1) { ++cu; tu -= 1.f; dp = 1; } else if (cross == -1) { --cu; tu += 1.f; dp = -1; }
                else if (cross == 2) { ++cv; tv -= 1.f; dp = nu; } else { --cv; tv += 1.f; dp = -nu; }
                p += dp;
                tuv = pk2(tu, tv);
                if (cu < 0 || cu > nu - 2 || cv < 0 || cv > nv - 2) { fast = false; break; }
                {                                                     // base plane of the new cell column
                    const float4 c00 = TT_LDN(p), c10 = TT_LDN(p + 1), c01 = TT_LDN(p + nu), c11 = TT_LDN(p + nu + 1);
                    tri2_base(qxy, TT_XY(c00), TT_XY(c10), TT_XY(c01), TT_XY(c11));
                    if (AUX) tri2_base(qzw, TT_ZW(c00), TT_ZW(c10), TT_ZW(c01), TT_ZW(c11));
                    else TT_QZ_BASE(c00.z, c10.z, c01.z, c11.z);
                }
                if (has_b) {
                    pa += dp;
                    const float4 c00 = GridT<float>::ld(pa), c10 = GridT<float>::ld(pa + 1), c01 = GridT<float>::ld(pa + nu), c11 = GridT<float>::ld(pa + nu + 1);
                    tri2_base(bxy, TT_XY(c00), TT_XY(c10), TT_XY(c01), TT_XY(c11));
                    tri2_base(bzk, TT_ZW(c00), TT_ZW(c10), TT_ZW(c01), TT_ZW(c11));
                }
                const float4* p1 = p + plane;                         // its far plane, into the prefetch registers
                n00 = TT_LDN(p1); n10 = TT_LDN(p1 + 1); n01 = TT_LDN(p1 + nu); n11 = TT_LDN(p1 + nu + 1);
            }
            if (SPC1 || renew) {                                      // common to both: the primed half from the far plane
                tri2_primed(qxy, TT_XY(n00), TT_XY(n10), TT_XY(n01), TT_XY(n11));
                if (AUX) tri2_primed(qzw, TT_ZW(n00), TT_ZW(n10), TT_ZW(n01), TT_ZW(n11));
                else TT_QZ_PRIMED(n00.z, n10.z, n01.z, n11.z);
                if (has_b) {
                    const float4* q1 = pa + plane;
                    const float4 b00 = GridT<float>::ld(q1), b10 = GridT<float>::ld(q1 + 1), b01 = GridT<float>::ld(q1 + nu), b11 = GridT<float>::ld(q1 + nu + 1);
                    tri2_primed(bxy, TT_XY(b00), TT_XY(b10), TT_XY(b01), TT_XY(b11));
                    tri2_primed(bzk, TT_ZW(b00), TT_ZW(b10), TT_ZW(b01), TT_ZW(b11));
                }
                have_next = false;
            }
#else
            if (cross == 0) {
                ++steps;
                fw = fw_t;
                if (SPC1 || ++j == spc) {
                    j = 0; fw = 0.f;
                    if (++k >= nw - 1) break;
                    p += plane;
#if TT_EVENT_PREFETCH && defined(__CUDA_ARCH__)
                    if (k + TT_EVENT_PREFETCH <= nw - 1) {                        // register-free L1 prefetch
                        prefetch_l1(p + TT_EVENT_PREFETCH * plane);
                        prefetch_l1(p + TT_EVENT_PREFETCH * plane + nu);
                    }
#endif
                    tri2_advance(qxy, TT_XY(n00), TT_XY(n10), TT_XY(n01), TT_XY(n11));
                    if (AUX) tri2_advance(qzw, TT_ZW(n00), TT_ZW(n10), TT_ZW(n01), TT_ZW(n11));
                    else { TT_QZ_STEP(); TT_QZ_PRIMED(n00.z, n10.z, n01.z, n11.z); }
                    if (has_b) {
                        pa += plane;
                        const float4* q1 = pa + plane;
                        const float4 b00 = GridT<float>::ld(q1), b10 = GridT<float>::ld(q1 + 1), b01 = GridT<float>::ld(q1 + nu), b11 = GridT<float>::ld(q1 + nu + 1);
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
#endif
        }
    }
    if (!fast) {
        status[ray] = TT_RAY_DEFERRED;
        deferred = true;
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
        if (TRACK_S && sf) {
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
    return steps;
}
#undef TT_LDN
#undef TT_QZ_AT
#undef TT_QZ_EVAL
#undef TT_QZ_BASE
#undef TT_QZ_STEP
#undef TT_QZ_PRIMED
#undef TT_XY
#undef TT_ZW

}  // namespace tt
