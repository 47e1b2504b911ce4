// Plane-marching ("gather") ray integrator of tt_trace variants 1 / 2 and of the second pass behind the event
// kernels: the marching loops and the per-ray body, compiled for the device (trace.cu) AND for the host
// (tests/host/trace_event_host.cu runs whole bundles -- event marching plus this second pass -- on the CPU).
// See trace.cu for the formulation.
#pragma once
#include "trace_common.cuh"

namespace tt {

// bookkeeping shared by the marching loops
template <typename T>
struct March {
    int& st;
    unsigned& steps;
    bool& alive;
    bool& general;
};

// Plane marching with a full 8-corner gather at every stage (the straightforward kernel; also
// used for the partial first cell of rays that enter through a side face and for side exits).
// Marches from (r.iw, r.fw) to the far face, or only to the next integer plane (until_plane).
template <typename T, bool AUX = false>
__host__ __device__ __noinline__ void march_generic(const typename GridT<T>::V4* __restrict__ grid, const Consts<T>& C,
                                           Ray<T>& r, March<T>& m, T s_left0, int spc, bool until_plane,
                                           AuxCtx<T>* ctx = nullptr) {
    const T hsub = T(1) / (T)spc;
    int k = r.iw;
    T fw = r.fw;
    int j = (int)(fw * (T)spc);          // sub-plane interval index containing fw
    while (k < C.nw - 1) {
        T fwb = (j + 1 == spc) ? T(1) : (T)(j + 1) * hsub;
        T h = fwb - fw;
        Ray<T> old = r;
        AuxCtx<T> old_ctx;
        if (AUX) old_ctx = *ctx;
        bool ok = zstep<T, AUX>(grid, C, r, k, fw, h, ctx);
        ++m.steps;
        bool bad = !ok || !(r.dw > T(TT_MARCH_MIN_DW)) || !(r.s <= s_left0);
        if (bad) {          // hand the step to the general integrator from the old state
            r = old; r.iw = k; r.fw = fw; --m.steps;
            if (AUX) *ctx = old_ctx;
            m.general = true;
            return;
        }
        T lu = leave_fraction(old.iu, old.fu, r.fu, C.nu);
        T lv = leave_fraction(old.iv, old.fv, r.fv, C.nv);
        T lam = fmin(lu, lv);
        if (lam <= T(1)) {   // side exit: re-step to the face, freeze
            r = old;
            if (AUX) *ctx = old_ctx;
            lam = lam < T(0) ? T(0) : lam;
            zstep<T, AUX>(grid, C, r, k, fw, lam * h, ctx);
            r.iw = k; r.fw = fw + lam * h;
            clamp_in(r.iu, r.fu, C.nu); clamp_in(r.iv, r.fv, C.nv);
            m.st |= TT_RAY_EXIT_SIDE;
            m.alive = false;
            return;
        }
        renorm(r.iu, r.fu); renorm(r.iv, r.fv);
        fw = fwb;
        if (++j == spc) {
            j = 0; ++k; fw = T(0);
            if (until_plane) { r.iw = k; r.fw = T(0); return; }
        }
    }
    r.iw = C.nw - 1; r.fw = T(0);
    m.st |= TT_RAY_EXIT_FACE;
    m.alive = false;
}

// ---- the fast path: plane marching with the ray's cell cached in registers ----------------------
// A ray moves ~1e-2 cells sideways per plane, so consecutive steps (and all four stages of a step)
// almost always sit in the same (u, v) cell column.  The 2 x 4 corners of the current cell are kept
// in registers as bilinear coefficient sets  g(tu, tv) = A + tu B + tv (C + tu D)  per plane (3 FMA
// per component), the next plane's 4 corners are prefetched while the step is computed, and a step
// then issues 4 instead of 32 corner loads.  A stage that leaves the cell gathers its 8 corners
// from memory (same arithmetic as march_generic); a ray that changes cell reloads its cache.
template <typename T>
struct PlaneC {
    T ax, bx, cx, dx, ay, by, cy, dy, az, bz, cz, dz;
};
template <typename T>
TT_HD void make_plane(PlaneC<T>& P, const typename GridT<T>::V4& c00,
                                           const typename GridT<T>::V4& c10, const typename GridT<T>::V4& c01,
                                           const typename GridT<T>::V4& c11) {
    P.ax = c00.x; P.bx = c10.x - c00.x; P.cx = c01.x - c00.x; P.dx = (c11.x - c01.x) - P.bx;
    P.ay = c00.y; P.by = c10.y - c00.y; P.cy = c01.y - c00.y; P.dy = (c11.y - c01.y) - P.by;
    P.az = c00.z; P.bz = c10.z - c00.z; P.cz = c01.z - c00.z; P.dz = (c11.z - c01.z) - P.bz;
}
template <typename T>
TT_HD void load_plane(PlaneC<T>& P, const typename GridT<T>::V4* __restrict__ p, int nu) {
    typedef typename GridT<T>::V4 V4;
    V4 c00 = GridT<T>::ld(p), c10 = GridT<T>::ld(p + 1), c01 = GridT<T>::ld(p + nu), c11 = GridT<T>::ld(p + nu + 1);
    make_plane<T>(P, c00, c10, c01, c11);
}
template <typename T>
TT_HD G3<T> eval_plane(const PlaneC<T>& P, T tu, T tv) {
    G3<T> g;
    g.x = tfma(tv, tfma(tu, P.dx, P.cx), tfma(tu, P.bx, P.ax));
    g.y = tfma(tv, tfma(tu, P.dy, P.cy), tfma(tu, P.by, P.ay));
    g.z = tfma(tv, tfma(tu, P.dz, P.cz), tfma(tu, P.bz, P.az));
    return g;
}
template <typename T, bool SPC1>
TT_HD void march_cached(const typename GridT<T>::V4* __restrict__ grid, const Consts<T>& C,
                                             Ray<T>& r, March<T>& m, T s_left0, int spc, bool track_s) {
    typedef typename GridT<T>::V4 V4;
    int k = r.iw;                                   // r.fw == 0: the ray sits on plane k
    if (k >= C.nw - 1) { r.iw = C.nw - 1; r.fw = T(0); m.st |= TT_RAY_EXIT_FACE; m.alive = false; return; }
    int cu, cv; T tu, tv;
    cell_of(r.iu, r.fu, C.nu, cu, tu); cell_of(r.iv, r.fv, C.nv, cv, tv);
    T du = r.du, dv = r.dv, dw = r.dw, s = r.s;
    const V4* p = grid + ((size_t)k * C.plane + (size_t)cv * C.nu + cu);
    PlaneC<T> P0, P1;
    load_plane<T>(P0, p, C.nu);
    load_plane<T>(P1, p + C.plane, C.nu);
    const T hsub = SPC1 ? T(1) : T(1) / (T)spc;

    // field at stage position (su, sv) of the current cell, w-fraction fwq (wsel 0: plane k, 1: plane
    // k+1, 2: general); outside the cached cell: full gather
    auto field = [&](T su, T sv, T fwq, int wsel) -> G3<T> {
        if (su >= T(0) && su <= T(1) && sv >= T(0) && sv <= T(1)) {
            if (wsel == 0) return eval_plane<T>(P0, su, sv);
            if (wsel == 1) return eval_plane<T>(P1, su, sv);
            G3<T> a = eval_plane<T>(P0, su, sv), b = eval_plane<T>(P1, su, sv), g;
            g.x = tfma(fwq, b.x - a.x, a.x); g.y = tfma(fwq, b.y - a.y, a.y); g.z = tfma(fwq, b.z - a.z, a.z);
            return g;
        }
        int c1, c2; T t1, t2;
        cell_of(cu, su, C.nu, c1, t1); cell_of(cv, sv, C.nv, c2, t2);
        return trilinear<T>(grid, C.nu, C.plane, c1, c2, k, t1, t2, fwq);
    };

    while (true) {
        // prefetch the 4 corners of plane k+2 for this cell; consumed when the step is done
        const bool has_next = k + 2 <= C.nw - 1;
        V4 n00, n10, n01, n11;
        if (has_next) {
            const V4* q = p + 2 * C.plane;
            n00 = GridT<T>::ld(q); n10 = GridT<T>::ld(q + 1); n01 = GridT<T>::ld(q + C.nu); n11 = GridT<T>::ld(q + C.nu + 1);
        }
        for (int j = 0; j < (SPC1 ? 1 : spc); ++j) {
            const T fwa = SPC1 ? T(0) : (T)j * hsub;
            const T fwb = SPC1 ? T(1) : ((j + 1 == spc) ? T(1) : (T)(j + 1) * hsub);
            const T h = fwb - fwa, half = T(0.5) * h;
            // ---- RK4 in W -----------------------------------------------------------------------
            G3<T> g = field(tu, tv, fwa, SPC1 ? 0 : 2);
            T q = trcp<T>(dw), hq = C.hw * q;
            bool ok = dw > T(0);
            const T aU = C.ru * du * q, aV = C.rv * dv * q, adu = g.x * hq, adv = g.y * hq, adw = g.z * hq, as = hq;
            T du2 = tfma(half, adu, du), dv2 = tfma(half, adv, dv), dw2 = tfma(half, adw, dw);
            g = field(tfma(half, aU, tu), tfma(half, aV, tv), fwa + half, 2);
            q = trcp<T>(dw2); hq = C.hw * q; ok = ok && dw2 > T(0);
            const T bU = C.ru * du2 * q, bV = C.rv * dv2 * q, bdu = g.x * hq, bdv = g.y * hq, bdw = g.z * hq, bs = hq;
            du2 = tfma(half, bdu, du); dv2 = tfma(half, bdv, dv); dw2 = tfma(half, bdw, dw);
            g = field(tfma(half, bU, tu), tfma(half, bV, tv), fwa + half, 2);
            q = trcp<T>(dw2); hq = C.hw * q; ok = ok && dw2 > T(0);
            const T cU = C.ru * du2 * q, cV = C.rv * dv2 * q, cdu = g.x * hq, cdv = g.y * hq, cdw = g.z * hq, cs = hq;
            du2 = tfma(h, cdu, du); dv2 = tfma(h, cdv, dv); dw2 = tfma(h, cdw, dw);
            g = field(tfma(h, cU, tu), tfma(h, cV, tv), fwb, SPC1 ? 1 : 2);
            q = trcp<T>(dw2); hq = C.hw * q; ok = ok && dw2 > T(0);
            const T eU = C.ru * du2 * q, eV = C.rv * dv2 * q, edu = g.x * hq, edv = g.y * hq, edw = g.z * hq, es = hq;
            const T h6 = h * T(1.0 / 6.0);
            const T tu_n = tfma(h6, aU + T(2) * (bU + cU) + eU, tu);
            const T tv_n = tfma(h6, aV + T(2) * (bV + cV) + eV, tv);
            const T du_n = tfma(h6, adu + T(2) * (bdu + cdu) + edu, du);
            const T dv_n = tfma(h6, adv + T(2) * (bdv + cdv) + edv, dv);
            const T dw_n = tfma(h6, adw + T(2) * (bdw + cdw) + edw, dw);
            const T s_n = track_s ? tfma(h6, as + T(2) * (bs + cs) + es, s) : s;
            ++m.steps;
            if (!ok || !(dw_n > T(TT_MARCH_MIN_DW)) || !(s_n <= s_left0)) {
                // steep / turning ray: the general integrator redoes this step from the old state
                r.iu = cu; r.fu = tu; r.iv = cv; r.fv = tv; r.iw = k; r.fw = fwa;
                r.du = du; r.dv = dv; r.dw = dw; r.s = s;
                --m.steps;
                m.general = true;
                return;
            }
            if (!(tu_n >= T(0) && tu_n <= T(1) && tv_n >= T(0) && tv_n <= T(1))) {
                // left the cached cell: side exit of the cube, or a new cell whose corners are reloaded
                T lam = fmin(leave_fraction(cu, tu, tu_n, C.nu), leave_fraction(cv, tv, tv_n, C.nv));
                if (lam <= T(1)) {
                    r.iu = cu; r.fu = tu; r.iv = cv; r.fv = tv; r.iw = k; r.fw = fwa;
                    r.du = du; r.dv = dv; r.dw = dw; r.s = s;
                    lam = lam < T(0) ? T(0) : lam;
                    zstep<T>(grid, C, r, k, fwa, lam * h);
                    r.fw = fwa + lam * h;
                    clamp_in(r.iu, r.fu, C.nu); clamp_in(r.iv, r.fv, C.nv);
                    m.st |= TT_RAY_EXIT_SIDE;
                    m.alive = false;
                    return;
                }
                int c1, c2; T t1, t2;
                cell_of(cu, tu_n, C.nu, c1, t1); cell_of(cv, tv_n, C.nv, c2, t2);
                cu = c1; cv = c2; tu = t1; tv = t2;
                p = grid + ((size_t)k * C.plane + (size_t)cv * C.nu + cu);
                if (!SPC1 && j + 1 < spc) load_plane<T>(P0, p, C.nu);
                load_plane<T>(P1, p + C.plane, C.nu);
                if (has_next) {
                    const V4* q2 = p + 2 * C.plane;
                    n00 = GridT<T>::ld(q2); n10 = GridT<T>::ld(q2 + 1); n01 = GridT<T>::ld(q2 + C.nu); n11 = GridT<T>::ld(q2 + C.nu + 1);
                }
            } else {
                tu = tu_n; tv = tv_n;
            }
            du = du_n; dv = dv_n; dw = dw_n; s = s_n;
        }
        ++k;
        if (k >= C.nw - 1) break;
        p += C.plane;
        P0 = P1;
        make_plane<T>(P1, n00, n10, n01, n11);
    }
    r.iu = cu; r.fu = tu; r.iv = cv; r.fv = tv; r.iw = C.nw - 1; r.fw = T(0);
    r.du = du; r.dv = dv; r.dw = dw; r.s = s;
    m.st |= TT_RAY_EXIT_FACE;
    m.alive = false;
}

// One ray from launch to exit plane; returns the RK4 steps taken (plane arrivals + arc-length steps).
template <typename T, int VARIANT, bool AUX>
TT_HD unsigned gather_ray(const typename GridT<T>::V4* __restrict__ grid, const double* __restrict__ s0, long ray,
                          double* __restrict__ rf, double* __restrict__ sf, uint8_t* __restrict__ status,
                          const TraceArgs& A, const typename GridT<T>::V4* __restrict__ aux4,
                          double* __restrict__ aux_out, const AuxArgs& AX) {
    unsigned steps = 0;
    double P[3], D[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        P[k] = s0[(size_t)A.fa[k] * A.np + ray];
        D[k] = s0[(size_t)(3 + A.fa[k]) * A.np + ray] * (1.0 / kC);
    }
    double X[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) X[k] = (P[k] - A.o[k]) / A.h[k];

    int st = 0;
    double s_acc = 0.0;     // path time spent before/inside the cube
    // ---- prologue: free flight to the cube if launched outside ------------------------------
    // (a NaN / inf launch state is "a ray that misses": fmin / fmax below would drop the NaN and turn it
    //  into a plausible ray; as MISSED its non-finite values reach rf untouched)
    bool finite = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) finite = finite && isfinite(P[k]) && isfinite(D[k]);
    bool inside = finite;
#pragma unroll
    for (int k = 0; k < 3; ++k) inside = inside && X[k] >= 0.0 && X[k] <= (double)(A.n[k] - 1);
    if (!inside) {
        double t_in = 0.0, t_out = 1e300;
        bool hit = finite;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double rate = D[k] / A.h[k], hi = (double)(A.n[k] - 1);
            if (rate == 0.0) {
                hit = hit && X[k] >= 0.0 && X[k] <= hi;
            } else {
                double ta = (0.0 - X[k]) / rate, tb = (hi - X[k]) / rate;
                t_in = fmax(t_in, fmin(ta, tb));
                t_out = fmin(t_out, fmax(ta, tb));
            }
        }
        hit = hit && t_in <= t_out && t_in < A.s_max;
        if (hit) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                X[k] += D[k] / A.h[k] * t_in;
                X[k] = fmin(fmax(X[k], 0.0), (double)(A.n[k] - 1));
            }
            s_acc = t_in;
        } else {
            st = TT_RAY_MISSED;
        }
    }

    Consts<T> C;
    C.nu = A.n[0]; C.nv = A.n[1]; C.nw = A.n[2];
    C.plane = (size_t)A.n[0] * A.n[1];
    C.ru = (T)(A.h[2] / A.h[0]); C.rv = (T)(A.h[2] / A.h[1]); C.hw = (T)A.h[2];
    C.iu_ = (T)(1.0 / A.h[0]); C.iv_ = (T)(1.0 / A.h[1]); C.iw_ = (T)(1.0 / A.h[2]);

    Ray<T> r;
    {
        double fl;
        fl = floor(X[0]); r.iu = (int)fl; r.fu = (T)(X[0] - fl);
        fl = floor(X[1]); r.iv = (int)fl; r.fv = (T)(X[1] - fl);
        fl = floor(X[2]); r.iw = (int)fl; r.fw = (T)(X[2] - fl);
    }
    r.du = (T)D[0]; r.dv = (T)D[1]; r.dw = (T)D[2];
    r.s = T(0);
    const T s_left0 = (T)(A.s_max - s_acc);    // path time available inside the cube

    bool alive = (st == 0);
    bool general = false;
    const T hsub = T(1) / (T)A.spc;

    // ---- plane marching ---------------------------------------------------------------------
    if (alive && !(r.dw > T(TT_MARCH_MIN_DW))) general = true;
    // rays that could run into the path-time cap c*T while marching are integrated by the general
    // loop, which stops exactly at the cap (never the case for the reference's symmetric cubes)
    if (alive && !general && (T)(C.nw - 1 - r.iw) * C.hw > T(TT_MARCH_MIN_DW) * s_left0) general = true;
    AuxCtx<T> ctx;
    ctx.aux4 = aux4; ctx.phase = 0.0; ctx.farad = 0.0; ctx.absorb = 0.0;
    if (alive && !general) {
        March<T> m{st, steps, alive, general};
        if (VARIANT == 1 || AUX) {
            march_generic<T, AUX>(grid, C, r, m, s_left0, A.spc, false, &ctx);
        } else {
            if (r.fw != T(0)) march_generic<T>(grid, C, r, m, s_left0, A.spc, true);   // entry through a side face
            if (alive && !general) {
                if (A.spc == 1) march_cached<T, true>(grid, C, r, m, s_left0, 1, sf != nullptr);
                else march_cached<T, false>(grid, C, r, m, s_left0, A.spc, sf != nullptr);
            }
        }
    }
    // ---- general arc-length integrator (steep / backward / time-capped rays) ----------------
    if (alive && general) {
        st |= TT_RAY_GENERAL;
        const T hmin = (T)fmin(A.h[0], fmin(A.h[1], A.h[2]));
        const T ds0 = hmin * hsub;
        // a ray sitting on a face and heading out leaves immediately (e.g. probing 'x' launch)
        for (long it = 0; it < (1L << 40); ++it) {
            T left = s_left0 - r.s;
            if (!(left > T(0))) { st |= TT_RAY_TIME_CAP; break; }
            T ds = ds0 < left ? ds0 : left;
            Ray<T> old = r;
            AuxCtx<T> old_ctx;
            if (AUX) old_ctx = ctx;
            sstep<T, AUX>(grid, C, r, ds, &ctx);
            ++steps;
            T lu = leave_fraction(old.iu, old.fu, r.fu, C.nu);
            T lv = leave_fraction(old.iv, old.fv, r.fv, C.nv);
            T lw = leave_fraction(old.iw, old.fw, r.fw, C.nw);
            T lam = fmin(lu, fmin(lv, lw));
            if (lam <= T(1)) {
                bool far_face = (lw <= lu && lw <= lv) && r.fw > old.fw;
                r = old;
                if (AUX) ctx = old_ctx;
                lam = lam < T(0) ? T(0) : lam;
                if (lam > T(0)) sstep<T, AUX>(grid, C, r, lam * ds, &ctx);
                clamp_in(r.iu, r.fu, C.nu); clamp_in(r.iv, r.fv, C.nv); clamp_in(r.iw, r.fw, C.nw);
                st |= far_face ? TT_RAY_EXIT_FACE : TT_RAY_EXIT_SIDE;
                break;
            }
            renorm(r.iu, r.fu); renorm(r.iv, r.fv); renorm(r.iw, r.fw);
            if (ds < ds0) { st |= TT_RAY_TIME_CAP; break; }
        }
        alive = false;
    }

    // ---- epilogue: ray_at_exit (particle_tracker.py:345-380) and state at time T -------------
    double Pf[3], Vf[3];
    if (st & TT_RAY_MISSED) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { Pf[k] = P[k]; Vf[k] = D[k] * kC; }
        s_acc = 0.0;
    } else {
        Pf[0] = A.o[0] + ((double)r.iu + (double)r.fu) * A.h[0];
        Pf[1] = A.o[1] + ((double)r.iv + (double)r.fv) * A.h[1];
        Pf[2] = A.o[2] + ((double)r.iw + (double)r.fw) * A.h[2];
        Vf[0] = (double)r.du * kC; Vf[1] = (double)r.dv * kC; Vf[2] = (double)r.dw * kC;
        s_acc += (double)r.s;
    }
    const double tb = (Pf[2] - A.extent) / Vf[2];
    rf[0 * A.np + ray] = Pf[0] - Vf[0] * tb;
    rf[1 * A.np + ray] = atan(Vf[0] / Vf[2]);
    rf[2 * A.np + ray] = Pf[1] - Vf[1] * tb;
    rf[3 * A.np + ray] = atan(Vf[1] / Vf[2]);
    if (sf) {
        const double t_rest = (A.s_max - s_acc) / kC;   // remaining free flight up to T
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            sf[(size_t)A.fa[k] * A.np + ray] = Pf[k] + Vf[k] * t_rest;
            sf[(size_t)(3 + A.fa[k]) * A.np + ray] = Vf[k];
        }
    }
    if (AUX) {
        aux_out[0 * A.np + ray] = exp(-0.5 * ctx.absorb);          // amplitude factor
        aux_out[1 * A.np + ray] = AX.omega_over_c * ctx.phase;     // phase (rad)
        aux_out[2 * A.np + ray] = AX.verdet_nc * ctx.farad;        // polarisation rotation (rad)
    }
    if (status) status[ray] = (uint8_t)st;
    return steps;
}

// ElectronCube.dndr (particle_tracker.py:243-256): trilinear gradient at point i of pos[3][npts], zero outside,
// faces inclusive (scipy _rgi.py:635-642); out[3][npts] in the reference's units (c^2 times the grid's 1/m)
template <typename T>
TT_HD void dndr_point(const typename GridT<T>::V4* __restrict__ grid, const TraceArgs& A, const double* __restrict__ pos,
                      long npts, long i, double* __restrict__ out) {
    double g3[3] = {0.0, 0.0, 0.0};
    double X[3];
    bool inside = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double p = pos[(size_t)A.fa[k] * npts + i];
        // inclusive faces tested on the physical coordinate, like the reference
        const double hi = A.o[k] + A.h[k] * (A.n[k] - 1);
        inside = inside && !(p < A.o[k]) && !(p > hi) && p == p;
        X[k] = (p - A.o[k]) / A.h[k];
    }
    if (inside) {
        int c[3]; T t[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double fl = floor(X[k]);
            cell_of<T>((int)fl, (T)(X[k] - fl), A.n[k], c[k], t[k]);
        }
        G3<T> g = trilinear<T>(grid, A.n[0], (size_t)A.n[0] * A.n[1], c[0], c[1], c[2], t[0], t[1], t[2]);
        g3[0] = (double)g.x; g3[1] = (double)g.y; g3[2] = (double)g.z;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) out[(size_t)A.fa[k] * npts + i] = g3[k] * (kC * kC);
}

}  // namespace tt
