// Event marching on uniform grids: the per-ray body of the scalar kernels (tt_trace variant 4, and variant 3 in
// FP64), compiled for the device (trace_event.cu) AND for the host (tests/host/trace_event_host.cu runs the very
// same source on the CPU against the C oracle).  The packed FP32x2 kernel of trace_event.cu performs the same
// operations in the same order (bit-identical rays, tested on the GPU), written with inline PTX.
#pragma once
#include "trace_common.cuh"

namespace tt {

#ifndef TT_EVENT_PREFETCH
#define TT_EVENT_PREFETCH 0        // planes ahead whose corner rows are prefetched into L1 (prefetch.global.L1).
                                   // Measured on B200 (513^3, 1e8 rays): 0 -> 470.7 ms, 2/4/8 -> 485-487 ms: the
                                   // register prefetch one plane ahead already covers L1 hits; off by default.
#endif
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

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
            V4 c00 = GridT<T>::ld(p), c10 = GridT<T>::ld(p + 1), c01 = GridT<T>::ld(p + nu), c11 = GridT<T>::ld(p + nu + 1);
            const V4* p1 = p + plane;
            V4 e00 = GridT<T>::ld(p1), e10 = GridT<T>::ld(p1 + 1), e01 = GridT<T>::ld(p1 + nu), e11 = GridT<T>::ld(p1 + nu + 1);
            tri_set<T>(qx, c00.x, c10.x, c01.x, c11.x, e00.x, e10.x, e01.x, e11.x);
            tri_set<T>(qy, c00.y, c10.y, c01.y, c11.y, e00.y, e10.y, e01.y, e11.y);
            tri_set<T>(qz, c00.z, c10.z, c01.z, c11.z, e00.z, e10.z, e01.z, e11.z);
        }
        bool have_next = false;
        while (true) {
            if (!have_next && k + 2 <= nw - 1) {
                const V4* p2 = p + 2 * plane;
                n00 = GridT<T>::ld(p2); n10 = GridT<T>::ld(p2 + 1); n01 = GridT<T>::ld(p2 + nu); n11 = GridT<T>::ld(p2 + nu + 1);
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
                    if (aU > T(0)) lu = (T(1) - tu) / (h * aU); else if (aU < T(0)) lu = -tu / (h * aU);
                    if (aV > T(0)) lv = (T(1) - tv) / (h * aV); else if (aV < T(0)) lv = -tv / (h * aV);
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
                V4 c00 = GridT<T>::ld(p), c10 = GridT<T>::ld(p + 1), c01 = GridT<T>::ld(p + nu), c11 = GridT<T>::ld(p + nu + 1);
                const V4* p1 = p + plane;
                V4 e00 = GridT<T>::ld(p1), e10 = GridT<T>::ld(p1 + 1), e01 = GridT<T>::ld(p1 + nu), e11 = GridT<T>::ld(p1 + nu + 1);
                tri_set<T>(qx, c00.x, c10.x, c01.x, c11.x, e00.x, e10.x, e01.x, e11.x);
                tri_set<T>(qy, c00.y, c10.y, c01.y, c11.y, e00.y, e10.y, e01.y, e11.y);
                tri_set<T>(qz, c00.z, c10.z, c01.z, c11.z, e00.z, e10.z, e01.z, e11.z);
                have_next = false;
            }
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

}  // namespace tt
