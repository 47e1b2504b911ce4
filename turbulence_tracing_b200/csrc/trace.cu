// K3 + K4: the ray integrator.  Replaces ElectronCube.solve / dsdt / dndr / ray_at_exit
// (particle_tracker.py:312-331, 398-419, 243-256, 333-380), i.e. scipy's global-step RK45 over
// three RegularGridInterpolator objects, by ONE kernel: one thread per ray, fixed-step RK4.
//
// Formulation (not a translation of the reference):
//   * non-dimensional state: position in grid-index units held as (int cell, fraction) per
//     axis, direction d = v/c, path time s = c*t.  Equations of motion
//         dX/ds = d / h,   dd/ds = g(X),   g = -1/2 grad(ne/nc)   (the reference's dnd?/c^2).
//   * plane marching: the probing-axis coordinate W is the independent variable,
//         dU/dW = (hw/hu) du/dw,  d(d)/dW = hw g/dw,  ds/dW = hw/dw,
//     so every RK4 step starts and ends exactly on a grid (sub-)plane of the probing axis and
//     never straddles the kink of the piecewise-trilinear field at a cell face in W.
//     steps_per_cell sub-planes per cell.  Stage coordinates are clamped into the cube; a ray
//     that would leave through a side face is re-stepped to the face and frozen there.
//   * rays that are steep or move backwards (d_w <= kMarchMinDw), and rays about to hit the
//     path-time cap s_max = c*T, continue in a general arc-length RK4 with the same field.
//   * outside the cube the field is exactly zero (fill_value=0.0), so rays are straight lines:
//     the prologue moves a ray launched outside to its entry point, the epilogue projects the
//     frozen exit state to the plane par-axis = +extent (ray_at_exit) and, if asked, to time T
//     (the reference's cube.sf).
// Template parameter T is the arithmetic/grid type: float (16 B corners) or double (32 B).
#include "common.cuh"

namespace tt {

static constexpr double kC = 299792458.0;      // scipy.constants.c (particle_tracker.py:119)
#define TT_MARCH_MIN_DW 0.75
#define TT_RAY_DEFERRED 0xFF       // internal status: ray left to the second-pass kernel
#ifndef TT_TRACE_MIN_BLOCKS
#define TT_TRACE_MIN_BLOCKS 3      // CTAs of 128 threads per SM the register allocation aims for
#endif

struct TraceArgs {
    int n[3];        // nu, nv, nw
    double o[3];     // origin per frame axis
    double h[3];     // spacing per frame axis
    int fa[3];       // frame axis -> xyz row
    double extent, s_max;
    int spc;
    long np;
    // precomputed for the event kernels (operands straight from the constant bank: no in-loop 64-bit
    // multiplies, no double->float conversions)
    long long plane_elems;   // nu * nv
    float hwf, ruf, rvf;     // (float) h_w, h_w/h_u, h_w/h_v
};

template <typename T> struct GridT;
template <> struct GridT<float> {
    typedef float4 V4;
    static __device__ __forceinline__ float4 ld(const float4* p) { return __ldg(p); }
};
template <> struct GridT<double> {
    typedef double4 V4;
    static __device__ __forceinline__ double4 ld(const double4* p) {
        const double2* q = reinterpret_cast<const double2*>(p);
        double2 a = __ldg(q), b = __ldg(q + 1);
        return make_double4(a.x, a.y, b.x, b.y);
    }
};

template <typename T> __device__ __forceinline__ T tfloor(T x);
template <> __device__ __forceinline__ float tfloor<float>(float x) { return floorf(x); }
template <> __device__ __forceinline__ double tfloor<double>(double x) { return floor(x); }
template <typename T> __device__ __forceinline__ T tfma(T a, T b, T c);
template <> __device__ __forceinline__ float tfma<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> __device__ __forceinline__ double tfma<double>(double a, double b, double c) { return fma(a, b, c); }

// cell/fraction of coordinate (i + f) clamped into [0, n-1]; the upper face is cell n-2, t = 1
// (as scipy's find_indices does for x == grid[-1]).
template <typename T>
__device__ __forceinline__ void cell_of(int i, T f, int n, int& c, T& t) {
    T fl = tfloor(f);
    c = i + (int)fl;
    t = f - fl;
    if (c < 0) { c = 0; t = T(0); }
    if (c > n - 2) { c = n - 2; t = T(1); }
}

template <typename T> struct G3 { T x, y, z; };

// trilinear gradient in cell (cu, cv, cw) at fractions (tu, tv, tw)
template <typename T>
__device__ __forceinline__ G3<T> trilinear(const typename GridT<T>::V4* __restrict__ grid, int nu,
                                           size_t plane, int cu, int cv, int cw, T tu, T tv, T tw) {
    typedef typename GridT<T>::V4 V4;
    const V4* p = grid + ((size_t)cw * plane + (size_t)cv * nu + cu);
    V4 c000 = GridT<T>::ld(p), c100 = GridT<T>::ld(p + 1);
    V4 c010 = GridT<T>::ld(p + nu), c110 = GridT<T>::ld(p + nu + 1);
    p += plane;
    V4 c001 = GridT<T>::ld(p), c101 = GridT<T>::ld(p + 1);
    V4 c011 = GridT<T>::ld(p + nu), c111 = GridT<T>::ld(p + nu + 1);
    G3<T> g;
#define TT_TRI(m)                                                        \
    {                                                                    \
        T a00 = tfma(tu, c100.m - c000.m, c000.m);                       \
        T a10 = tfma(tu, c110.m - c010.m, c010.m);                       \
        T a01 = tfma(tu, c101.m - c001.m, c001.m);                       \
        T a11 = tfma(tu, c111.m - c011.m, c011.m);                       \
        T b0 = tfma(tv, a10 - a00, a00);                                 \
        T b1 = tfma(tv, a11 - a01, a01);                                 \
        g.m = tfma(tw, b1 - b0, b0);                                     \
    }
    TT_TRI(x) TT_TRI(y) TT_TRI(z)
#undef TT_TRI
    return g;
}

template <typename T>
struct Ray {
    int iu, iv, iw;
    T fu, fv, fw;     // fractions relative to (iu, iv, iw); may be un-normalised after a step
    T du, dv, dw;
    T s;              // path time c*t accumulated inside the cube
};

template <typename T>
struct Consts {
    int nu, nv, nw;
    size_t plane;
    T ru, rv, hw;     // hw/hu, hw/hv, hw  (plane marching)
    T iu_, iv_, iw_;  // 1/hu, 1/hv, 1/hw  (arc-length stepping)
};

// ---- passive quantities carried along the ray (BASELINE config 4; no implementation in the reference
// checkout -- only call sites, example_kitchensink.py:72-101 -- so parity is unpinned; textbook forms):
//   phase          dphi/ds   = (omega/c) (n - 1),        n = sqrt(1 - ne/nc)
//   Faraday        dalpha/ds = V ne (B . d),              V = e^3 lambda^2 / (8 pi^2 eps0 me^2 c^3)
//   inv. brems.    dln(a)/ds = -kappa / 2                 (kappa: energy absorption coefficient, 1/m)
// with s = c t and d = v/c.  They do not act back on the trajectory, so the RK4 stages of the ray are
// reused as a Simpson quadrature (weights 1, 2, 2, 1).  aux4 = (B_u, B_v, B_w, kappa) on the same grid
// layout as the gradient grid, whose 4th lane is ne/nc.  Accumulated in FP64; scaled by the constants
// in the epilogue.
template <typename T>
struct AuxCtx {
    const typename GridT<T>::V4* aux4;     // may be null: phase only
    double phase, farad, absorb;           // integrals of (n-1), (ne/nc)(B.d), kappa over s
};

template <typename T>
__device__ __forceinline__ void trilinear_w(const typename GridT<T>::V4* __restrict__ grid, int nu, size_t plane,
                                            int cu, int cv, int cw, T tu, T tv, T tw, T& w_only) {
    typedef typename GridT<T>::V4 V4;
    const V4* p = grid + ((size_t)cw * plane + (size_t)cv * nu + cu);
    T c000 = GridT<T>::ld(p).w, c100 = GridT<T>::ld(p + 1).w, c010 = GridT<T>::ld(p + nu).w, c110 = GridT<T>::ld(p + nu + 1).w;
    p += plane;
    T c001 = GridT<T>::ld(p).w, c101 = GridT<T>::ld(p + 1).w, c011 = GridT<T>::ld(p + nu).w, c111 = GridT<T>::ld(p + nu + 1).w;
    T a00 = tfma(tu, c100 - c000, c000), a10 = tfma(tu, c110 - c010, c010);
    T a01 = tfma(tu, c101 - c001, c001), a11 = tfma(tu, c111 - c011, c011);
    T b0 = tfma(tv, a10 - a00, a00), b1 = tfma(tv, a11 - a01, a01);
    w_only = tfma(tw, b1 - b0, b0);
}

template <typename T>
__device__ __forceinline__ void trilinear4(const typename GridT<T>::V4* __restrict__ grid, int nu, size_t plane,
                                           int cu, int cv, int cw, T tu, T tv, T tw, T& x, T& y, T& z, T& w) {
    typedef typename GridT<T>::V4 V4;
    const V4* p = grid + ((size_t)cw * plane + (size_t)cv * nu + cu);
    V4 c000 = GridT<T>::ld(p), c100 = GridT<T>::ld(p + 1), c010 = GridT<T>::ld(p + nu), c110 = GridT<T>::ld(p + nu + 1);
    p += plane;
    V4 c001 = GridT<T>::ld(p), c101 = GridT<T>::ld(p + 1), c011 = GridT<T>::ld(p + nu), c111 = GridT<T>::ld(p + nu + 1);
#define TT_TRI4(m, out)                                                  \
    {                                                                    \
        T a00 = tfma(tu, c100.m - c000.m, c000.m);                       \
        T a10 = tfma(tu, c110.m - c010.m, c010.m);                       \
        T a01 = tfma(tu, c101.m - c001.m, c001.m);                       \
        T a11 = tfma(tu, c111.m - c011.m, c011.m);                       \
        T b0 = tfma(tv, a10 - a00, a00);                                 \
        T b1 = tfma(tv, a11 - a01, a01);                                 \
        out = tfma(tw, b1 - b0, b0);                                     \
    }
    TT_TRI4(x, x) TT_TRI4(y, y) TT_TRI4(z, z) TT_TRI4(w, w)
#undef TT_TRI4
}

// integrands at one stage, multiplied by `scale` (ds/dW = hw/dw when marching in W, 1 in path time)
template <typename T>
__device__ __forceinline__ void aux_rates(const AuxCtx<T>& ctx, const typename GridT<T>::V4* __restrict__ grid,
                                          const Consts<T>& C, int cu, int cv, int cw, T tu, T tv, T tw, T du, T dv,
                                          T dw, T scale, double& fp, double& ff, double& fa) {
    T nn;
    trilinear_w<T>(grid, C.nu, C.plane, cu, cv, cw, tu, tv, tw, nn);       // ne/nc
    const double nr = 1.0 - (double)nn;
    fp = (sqrt(nr > 0.0 ? nr : 0.0) - 1.0) * (double)scale;
    ff = 0.0; fa = 0.0;
    if (ctx.aux4) {
        T bu, bv, bw, kap;
        trilinear4<T>(ctx.aux4, C.nu, C.plane, cu, cv, cw, tu, tv, tw, bu, bv, bw, kap);
        ff = (double)nn * ((double)bu * du + (double)bv * dv + (double)bw * dw) * (double)scale;
        fa = (double)kap * (double)scale;
    }
}

// One RK4 step in W from fraction fwa to fwa + h inside w-cell k (0 <= fwa, fwa + h <= 1).
// Updates fu, fv (un-normalised), d and s of r.  Returns false if a stage saw d_w <= 0.
template <typename T, bool AUX = false>
__device__ __forceinline__ bool zstep(const typename GridT<T>::V4* __restrict__ grid, const Consts<T>& C,
                                      Ray<T>& r, int k, T fwa, T h, AuxCtx<T>* ctx = nullptr) {
    const T half = T(0.5) * h;
    int cu, cv; T tu, tv;
    double fp[4], ff[4], fa[4];
    // stage 1
    cell_of(r.iu, r.fu, C.nu, cu, tu); cell_of(r.iv, r.fv, C.nv, cv, tv);
    G3<T> g = trilinear<T>(grid, C.nu, C.plane, cu, cv, k, tu, tv, fwa);
    T inv = C.hw / r.dw;
    if (AUX) aux_rates<T>(*ctx, grid, C, cu, cv, k, tu, tv, fwa, r.du, r.dv, r.dw, inv, fp[0], ff[0], fa[0]);
    T aU = C.ru * r.du / r.dw, aV = C.rv * r.dv / r.dw;
    T adu = g.x * inv, adv = g.y * inv, adw = g.z * inv, as = inv;
    bool ok = r.dw > T(0);
    // stage 2
    T du = tfma(half, adu, r.du), dv = tfma(half, adv, r.dv), dw = tfma(half, adw, r.dw);
    cell_of(r.iu, tfma(half, aU, r.fu), C.nu, cu, tu); cell_of(r.iv, tfma(half, aV, r.fv), C.nv, cv, tv);
    g = trilinear<T>(grid, C.nu, C.plane, cu, cv, k, tu, tv, fwa + half);
    ok = ok && dw > T(0);
    inv = C.hw / dw;
    if (AUX) aux_rates<T>(*ctx, grid, C, cu, cv, k, tu, tv, fwa + half, du, dv, dw, inv, fp[1], ff[1], fa[1]);
    T bU = C.ru * du / dw, bV = C.rv * dv / dw;
    T bdu = g.x * inv, bdv = g.y * inv, bdw = g.z * inv, bs = inv;
    // stage 3
    du = tfma(half, bdu, r.du); dv = tfma(half, bdv, r.dv); dw = tfma(half, bdw, r.dw);
    cell_of(r.iu, tfma(half, bU, r.fu), C.nu, cu, tu); cell_of(r.iv, tfma(half, bV, r.fv), C.nv, cv, tv);
    g = trilinear<T>(grid, C.nu, C.plane, cu, cv, k, tu, tv, fwa + half);
    ok = ok && dw > T(0);
    inv = C.hw / dw;
    if (AUX) aux_rates<T>(*ctx, grid, C, cu, cv, k, tu, tv, fwa + half, du, dv, dw, inv, fp[2], ff[2], fa[2]);
    T cU = C.ru * du / dw, cV = C.rv * dv / dw;
    T cdu = g.x * inv, cdv = g.y * inv, cdw = g.z * inv, cs = inv;
    // stage 4
    du = tfma(h, cdu, r.du); dv = tfma(h, cdv, r.dv); dw = tfma(h, cdw, r.dw);
    cell_of(r.iu, tfma(h, cU, r.fu), C.nu, cu, tu); cell_of(r.iv, tfma(h, cV, r.fv), C.nv, cv, tv);
    T fwb = fwa + h;
    g = trilinear<T>(grid, C.nu, C.plane, cu, cv, k, tu, tv, fwb > T(1) ? T(1) : fwb);
    ok = ok && dw > T(0);
    inv = C.hw / dw;
    if (AUX) aux_rates<T>(*ctx, grid, C, cu, cv, k, tu, tv, fwb > T(1) ? T(1) : fwb, du, dv, dw, inv, fp[3], ff[3], fa[3]);
    T eU = C.ru * du / dw, eV = C.rv * dv / dw;
    T edu = g.x * inv, edv = g.y * inv, edw = g.z * inv, es = inv;
    const T h6 = h * T(1.0 / 6.0);
    r.fu = tfma(h6, aU + T(2) * (bU + cU) + eU, r.fu);
    r.fv = tfma(h6, aV + T(2) * (bV + cV) + eV, r.fv);
    r.du = tfma(h6, adu + T(2) * (bdu + cdu) + edu, r.du);
    r.dv = tfma(h6, adv + T(2) * (bdv + cdv) + edv, r.dv);
    r.dw = tfma(h6, adw + T(2) * (bdw + cdw) + edw, r.dw);
    r.s = tfma(h6, as + T(2) * (bs + cs) + es, r.s);
    if (AUX) {
        const double w6 = (double)h6;
        ctx->phase += w6 * (fp[0] + 2.0 * (fp[1] + fp[2]) + fp[3]);
        ctx->farad += w6 * (ff[0] + 2.0 * (ff[1] + ff[2]) + ff[3]);
        ctx->absorb += w6 * (fa[0] + 2.0 * (fa[1] + fa[2]) + fa[3]);
    }
    return ok;
}

// One RK4 step of length ds in path time (general direction).  Fractions un-normalised after.
template <typename T, bool AUX = false>
__device__ __forceinline__ void sstep(const typename GridT<T>::V4* __restrict__ grid, const Consts<T>& C,
                                      Ray<T>& r, T ds, AuxCtx<T>* ctx = nullptr) {
    const T half = T(0.5) * ds;
    int cu, cv, cw; T tu, tv, tw;
    double fp[4], ff[4], fa[4];
    int stage = 0;
    auto field = [&](T fu, T fv, T fw, T du, T dv, T dw) {
        cell_of(r.iu, fu, C.nu, cu, tu); cell_of(r.iv, fv, C.nv, cv, tv); cell_of(r.iw, fw, C.nw, cw, tw);
        if (AUX) { aux_rates<T>(*ctx, grid, C, cu, cv, cw, tu, tv, tw, du, dv, dw, T(1), fp[stage], ff[stage], fa[stage]); ++stage; }
        return trilinear<T>(grid, C.nu, C.plane, cu, cv, cw, tu, tv, tw);
    };
    G3<T> g1 = field(r.fu, r.fv, r.fw, r.du, r.dv, r.dw);
    T d1u = r.du, d1v = r.dv, d1w = r.dw;
    T d2u = tfma(half, g1.x, r.du), d2v = tfma(half, g1.y, r.dv), d2w = tfma(half, g1.z, r.dw);
    G3<T> g2 = field(tfma(half * C.iu_, d1u, r.fu), tfma(half * C.iv_, d1v, r.fv), tfma(half * C.iw_, d1w, r.fw), d2u, d2v, d2w);
    T d3u = tfma(half, g2.x, r.du), d3v = tfma(half, g2.y, r.dv), d3w = tfma(half, g2.z, r.dw);
    G3<T> g3 = field(tfma(half * C.iu_, d2u, r.fu), tfma(half * C.iv_, d2v, r.fv), tfma(half * C.iw_, d2w, r.fw), d3u, d3v, d3w);
    T d4u = tfma(ds, g3.x, r.du), d4v = tfma(ds, g3.y, r.dv), d4w = tfma(ds, g3.z, r.dw);
    G3<T> g4 = field(tfma(ds * C.iu_, d3u, r.fu), tfma(ds * C.iv_, d3v, r.fv), tfma(ds * C.iw_, d3w, r.fw), d4u, d4v, d4w);
    const T s6 = ds * T(1.0 / 6.0);
    if (AUX) {
        const double w6 = (double)s6;
        ctx->phase += w6 * (fp[0] + 2.0 * (fp[1] + fp[2]) + fp[3]);
        ctx->farad += w6 * (ff[0] + 2.0 * (ff[1] + ff[2]) + ff[3]);
        ctx->absorb += w6 * (fa[0] + 2.0 * (fa[1] + fa[2]) + fa[3]);
    }
    r.fu = tfma(s6 * C.iu_, d1u + T(2) * (d2u + d3u) + d4u, r.fu);
    r.fv = tfma(s6 * C.iv_, d1v + T(2) * (d2v + d3v) + d4v, r.fv);
    r.fw = tfma(s6 * C.iw_, d1w + T(2) * (d2w + d3w) + d4w, r.fw);
    r.du = tfma(s6, g1.x + T(2) * (g2.x + g3.x) + g4.x, r.du);
    r.dv = tfma(s6, g1.y + T(2) * (g2.y + g3.y) + g4.y, r.dv);
    r.dw = tfma(s6, g1.z + T(2) * (g2.z + g3.z) + g4.z, r.dw);
    r.s += ds;
}

// fraction of the chord old -> new at which coordinate (i + f) leaves [0, n-1]; 2 if it does not
template <typename T>
__device__ __forceinline__ T leave_fraction(int i, T f_old, T f_new, int n) {
    T lo = T(-i), hi = T(n - 1 - i);
    if (f_new < lo) return (lo - f_old) / (f_new - f_old);
    if (f_new > hi) return (hi - f_old) / (f_new - f_old);
    return T(2);
}

template <typename T>
__device__ __forceinline__ void renorm(int& i, T& f) {
    T fl = tfloor(f);
    i += (int)fl;
    f -= fl;
}
// snap a coordinate that should lie on/inside the faces back into [0, n-1]
template <typename T>
__device__ __forceinline__ void clamp_in(int& i, T& f, int n) {
    renorm(i, f);
    if (i < 0) { i = 0; f = T(0); }
    if (i > n - 1 || (i == n - 1 && f > T(0))) { i = n - 1; f = T(0); }
}


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
__device__ __noinline__ void march_generic(const typename GridT<T>::V4* __restrict__ grid, const Consts<T>& C,
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
__device__ __forceinline__ void make_plane(PlaneC<T>& P, const typename GridT<T>::V4& c00,
                                           const typename GridT<T>::V4& c10, const typename GridT<T>::V4& c01,
                                           const typename GridT<T>::V4& c11) {
    P.ax = c00.x; P.bx = c10.x - c00.x; P.cx = c01.x - c00.x; P.dx = (c11.x - c01.x) - P.bx;
    P.ay = c00.y; P.by = c10.y - c00.y; P.cy = c01.y - c00.y; P.dy = (c11.y - c01.y) - P.by;
    P.az = c00.z; P.bz = c10.z - c00.z; P.cz = c01.z - c00.z; P.dz = (c11.z - c01.z) - P.bz;
}
template <typename T>
__device__ __forceinline__ void load_plane(PlaneC<T>& P, const typename GridT<T>::V4* __restrict__ p, int nu) {
    typedef typename GridT<T>::V4 V4;
    V4 c00 = GridT<T>::ld(p), c10 = GridT<T>::ld(p + 1), c01 = GridT<T>::ld(p + nu), c11 = GridT<T>::ld(p + nu + 1);
    make_plane<T>(P, c00, c10, c01, c11);
}
template <typename T>
__device__ __forceinline__ G3<T> eval_plane(const PlaneC<T>& P, T tu, T tv) {
    G3<T> g;
    g.x = tfma(tv, tfma(tu, P.dx, P.cx), tfma(tu, P.bx, P.ax));
    g.y = tfma(tv, tfma(tu, P.dy, P.cy), tfma(tu, P.by, P.ay));
    g.z = tfma(tv, tfma(tu, P.dz, P.cz), tfma(tu, P.bz, P.az));
    return g;
}
template <typename T> __device__ __forceinline__ T trcp(T x);
template <> __device__ __forceinline__ float trcp<float>(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);        // one Newton step: ~1 ulp
}
template <> __device__ __forceinline__ double trcp<double>(double x) { return 1.0 / x; }

template <typename T, bool SPC1>
__device__ __forceinline__ void march_cached(const typename GridT<T>::V4* __restrict__ grid, const Consts<T>& C,
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

struct AuxArgs {
    double omega_over_c;    // phase = omega/c * int (n - 1) ds
    double verdet_nc;       // rotation = V * nc * int (ne/nc) (B . d) ds
};

template <typename T, int VARIANT, bool AUX = false>
__global__ void __launch_bounds__(128, sizeof(T) == 8 ? 2 : TT_TRACE_MIN_BLOCKS) trace_kernel(const typename GridT<T>::V4* __restrict__ grid,
                                                    const double* __restrict__ s0,
                                                    const uint32_t* __restrict__ perm,
                                                    double* __restrict__ rf, double* __restrict__ sf,
                                                    unsigned long long* __restrict__ ray_steps,
                                                    uint8_t* __restrict__ status, TraceArgs A, int only_flagged,
                                                    const typename GridT<T>::V4* __restrict__ aux4 = nullptr,
                                                    double* __restrict__ aux_out = nullptr, AuxArgs AX = AuxArgs()) {
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned steps = 0;
    bool mine = tid < A.np;
    long ray = 0;
    if (mine) {
        // second pass behind trace_event_kernel: only the rays it handed over, visited in ray order
        // (a coalesced scan of the flags; the Morton order only matters for the bulk of the rays)
        ray = (perm && !only_flagged) ? (long)perm[tid] : tid;
        if (only_flagged && status[ray] != TT_RAY_DEFERRED) mine = false;
    }
    if (mine) {
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
        bool inside = true;
#pragma unroll
        for (int k = 0; k < 3; ++k) inside = inside && X[k] >= 0.0 && X[k] <= (double)(A.n[k] - 1);
        if (!inside) {
            double t_in = 0.0, t_out = 1e300;
            bool hit = true;
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
    }
    // ---- ray-step count: warp reduce, one atomic per warp ------------------------------------
    if (ray_steps) {
        unsigned v = steps;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(ray_steps, (unsigned long long)v);
    }
}

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
#ifndef TT_EVENT_MIN_BLOCKS_F64
#define TT_EVENT_MIN_BLOCKS_F64 3
#endif
#ifndef TT_EVENT_PREFETCH
#define TT_EVENT_PREFETCH 0        // planes ahead whose corner rows are prefetched into L1 (prefetch.global.L1).
                                   // Measured on B200 (513^3, 1e8 rays): 0 -> 470.7 ms, 2/4/8 -> 485-487 ms: the
                                   // register prefetch one plane ahead already covers L1 hits; off by default.
#endif
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <typename T>
struct Tri {           // trilinear polynomial of one component in one cell
    T a, b, c, d, a1, b1, c1, d1;
};
// evaluation is "w first": the bilinear coefficients at w-fraction fw (4 FMA), then the bilinear form
// (3 FMA).  RK4 stages 2 and 3 share their w-fraction, so their coefficients are formed once.
template <typename T>
struct Bil {
    T a, b, c, d;
};
template <typename T>
__device__ __forceinline__ Bil<T> tri_at(const Tri<T>& q, T fw) {
    Bil<T> r;
    r.a = tfma(fw, q.a1, q.a); r.b = tfma(fw, q.b1, q.b); r.c = tfma(fw, q.c1, q.c); r.d = tfma(fw, q.d1, q.d);
    return r;
}
template <typename T>
__device__ __forceinline__ T bil_eval(const Bil<T>& q, T tu, T tv) {
    return tfma(tv, tfma(tu, q.d, q.c), tfma(tu, q.b, q.a));
}
// coefficients of plane 0 from its 4 corners; primed = plane 1 minus plane 0
template <typename T>
__device__ __forceinline__ void tri_set(Tri<T>& q, T c00, T c10, T c01, T c11, T e00, T e10, T e01, T e11) {
    q.a = c00; q.b = c10 - c00; q.c = c01 - c00; q.d = (c11 - c01) - q.b;
    T ea = e00, eb = e10 - e00, ec = e01 - e00, ed = (e11 - e01) - eb;
    q.a1 = ea - q.a; q.b1 = eb - q.b; q.c1 = ec - q.c; q.d1 = ed - q.d;
}
// advance one plane: plane 1 becomes plane 0, (n00..n11) are the corners of the new plane 1
template <typename T>
__device__ __forceinline__ void tri_advance(Tri<T>& q, T n00, T n10, T n01, T n11) {
    q.a += q.a1; q.b += q.b1; q.c += q.c1; q.d += q.d1;
    T eb = n10 - n00;
    q.a1 = n00 - q.a; q.b1 = eb - q.b; q.c1 = (n01 - n00) - q.c; q.d1 = ((n11 - n01) - eb) - q.d;
}

template <typename T, bool SPC1>
__global__ void __launch_bounds__(128, sizeof(T) == 8 ? TT_EVENT_MIN_BLOCKS_F64 : TT_EVENT_MIN_BLOCKS)
trace_event_kernel(const typename GridT<T>::V4* __restrict__ grid, const double* __restrict__ s0,
                   const uint32_t* __restrict__ perm, double* __restrict__ rf, double* __restrict__ sf,
                   unsigned long long* __restrict__ ray_steps, uint8_t* __restrict__ status, TraceArgs A) {
    typedef typename GridT<T>::V4 V4;
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned steps = 0;
    if (tid < A.np) {
        const long ray = perm ? (long)perm[tid] : tid;
        const int nu = A.n[0], nv = A.n[1], nw = A.n[2];
        const long long plane = A.plane_elems;
        // ---- prologue ---------------------------------------------------------------------------
        double X[3], D[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            X[k] = (s0[(size_t)A.fa[k] * A.np + ray] - A.o[k]) / A.h[k];
            D[k] = s0[(size_t)(3 + A.fa[k]) * A.np + ray] * (1.0 / kC);
        }
        bool fast = X[0] >= 0.0 && X[0] <= (double)(nu - 1) && X[1] >= 0.0 && X[1] <= (double)(nv - 1) &&
                    X[2] >= 0.0 && X[2] <= (double)(nw - 1) && D[2] > TT_MARCH_MIN_DW;
        // the path-time cap must be out of reach while marching (d_w > 0.75 throughout)
        fast = fast && ((double)(nw - 1) - X[2]) * A.h[2] <= TT_MARCH_MIN_DW * A.s_max;
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
                        if (TT_EVENT_PREFETCH && k + TT_EVENT_PREFETCH <= nw - 1) {   // register-free L1 prefetch
                            prefetch_l1(p + TT_EVENT_PREFETCH * plane);
                            prefetch_l1(p + TT_EVENT_PREFETCH * plane + nu);
                        }
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
                const double t_rest = (A.s_max - (double)s) / kC;
                const double Pf[3] = {Pu, Pv, Pw}, Vf[3] = {Vu, Vv, Vw};
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    sf[(size_t)A.fa[m] * A.np + ray] = Pf[m] + Vf[m] * t_rest;
                    sf[(size_t)(3 + A.fa[m]) * A.np + ray] = Vf[m];
                }
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
__device__ __forceinline__ float lo2(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi2(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
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

template <bool SPC1, bool AUX>
__global__ void __launch_bounds__(128, AUX ? 3 : TT_EVENT_MIN_BLOCKS)
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
        bool fast = X[0] >= 0.0 && X[0] <= (double)(nu - 1) && X[1] >= 0.0 && X[1] <= (double)(nv - 1) &&
                    X[2] >= 0.0 && X[2] <= (double)(nw - 1) && D[2] > TT_MARCH_MIN_DW;
        fast = fast && ((double)(nw - 1) - X[2]) * A.h[2] <= TT_MARCH_MIN_DW * A.s_max;
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
                const f32x2 aUV = mul2(mul2(RUV, duv), bc2(q));
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
                const f32x2 bUV = mul2(mul2(RUV, duv2), bc2(q));
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
                const f32x2 cUV = mul2(mul2(RUV, duv2), bc2(q));
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
                const f32x2 eUV = mul2(mul2(RUV, duv2), bc2(q));
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
                const double t_rest = (A.s_max - (double)s) / kC;
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

// ElectronCube.dndr (particle_tracker.py:243-256): trilinear gradient at arbitrary points,
// zero outside, faces inclusive (scipy _rgi.py:635-642).
template <typename T>
__global__ void dndr_kernel(const typename GridT<T>::V4* __restrict__ grid, TraceArgs A,
                            const double* __restrict__ pos, long npts, double* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
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

static int fill_args(TraceArgs& A, const int n_xyz[3], const double origin_xyz[3],
                     const double spacing_xyz[3], int par) {
    TT_REQUIRE(n_xyz && origin_xyz && spacing_xyz, "null geometry pointer");
    TT_REQUIRE(par >= 0 && par <= 2, "par must be 0, 1 or 2 (got %d)", par);
    Frame f = frame_of(par);
    for (int k = 0; k < 3; ++k) {
        A.fa[k] = f.a[k];
        A.n[k] = n_xyz[f.a[k]];
        A.o[k] = origin_xyz[f.a[k]];
        A.h[k] = spacing_xyz[f.a[k]];
        TT_REQUIRE(A.n[k] >= 2, "every axis needs >= 2 points");
        TT_REQUIRE(A.h[k] > 0, "spacing must be > 0");
    }
    A.plane_elems = (long long)A.n[0] * A.n[1];
    A.hwf = (float)A.h[2]; A.ruf = (float)(A.h[2] / A.h[0]); A.rvf = (float)(A.h[2] / A.h[1]);
    return TT_OK;
}

}  // namespace tt

extern "C" int tt_trace(const tt_trace_params* p, const void* grid4_dev, const double* s0_dev, long np,
                        const uint32_t* perm_dev, double* rf_dev, double* sf_dev,
                        unsigned long long* ray_steps_dev, uint8_t* status_dev, tt_stream_t stream) {
    using namespace tt;
    TT_REQUIRE(p && grid4_dev && s0_dev && rf_dev, "tt_trace: null pointer");
    TT_REQUIRE(np >= 0, "tt_trace: negative ray count");
    TT_REQUIRE(p->dtype == TT_F32 || p->dtype == TT_F64, "tt_trace: dtype must be TT_F32 or TT_F64");
    TT_REQUIRE(p->steps_per_cell >= 1 && p->steps_per_cell <= 1024, "tt_trace: steps_per_cell out of range");
    TT_REQUIRE(p->s_max > 0 && p->extent == p->extent, "tt_trace: s_max must be > 0");
    TT_REQUIRE(np < (1L << 32) || !perm_dev, "tt_trace: perm is 32-bit; trace in bundles of < 2^32 rays");
    TraceArgs A;
    int rc = fill_args(A, p->n_xyz, p->origin_xyz, p->spacing_xyz, p->par);
    if (rc) return rc;
    A.extent = p->extent; A.s_max = p->s_max; A.spc = p->steps_per_cell; A.np = np;
    if (np == 0) return TT_OK;
    const int block = 128;
    const long blocks = (np + block - 1) / block;
    TT_REQUIRE(blocks < (1L << 31), "tt_trace: too many rays for one launch");
    cudaStream_t s = (cudaStream_t)stream;
    // variant: 0 = auto (event marching when a status buffer is given, else the cell-cache kernel),
    //          1 = 8-corner gather per stage, 2 = cell cache, 3 = event marching (needs status_dev)
    //          4 = event marching without the packed FP32x2 arithmetic (cross-check of 3 in FP32)
    int variant = p->variant;
    TT_REQUIRE(variant >= 0 && variant <= 4, "tt_trace: unknown kernel variant %d", variant);
    TT_REQUIRE(variant < 3 || status_dev, "tt_trace: event marching (variant 3/4) needs status_dev");
    if (variant == 0) variant = status_dev ? 3 : 2;
    int only_flagged = 0;
    if (variant == 3 && p->dtype == TT_F32) {
        if (p->steps_per_cell == 1)
            trace_event_kernel_f32x2<true, false><<<(unsigned)blocks, block, 0, s>>>((const float4*)grid4_dev, s0_dev, perm_dev, rf_dev,
                                                                              sf_dev, ray_steps_dev, status_dev, A);
        else
            trace_event_kernel_f32x2<false, false><<<(unsigned)blocks, block, 0, s>>>((const float4*)grid4_dev, s0_dev, perm_dev, rf_dev,
                                                                               sf_dev, ray_steps_dev, status_dev, A);
        int rc2 = launch_check("trace_event_kernel_f32x2");
        if (rc2) return rc2;
        only_flagged = 1;
        variant = 2;
    } else if (variant >= 3) {
        const bool spc1 = p->steps_per_cell == 1;
#define TT_LAUNCH_EV(TYPE, V4T, S1)                                                                                \
    trace_event_kernel<TYPE, S1><<<(unsigned)blocks, block, 0, s>>>((const V4T*)grid4_dev, s0_dev, perm_dev, rf_dev, \
                                                                     sf_dev, ray_steps_dev, status_dev, A)
        if (p->dtype == TT_F32) { if (spc1) TT_LAUNCH_EV(float, float4, true); else TT_LAUNCH_EV(float, float4, false); }
        else { if (spc1) TT_LAUNCH_EV(double, double4, true); else TT_LAUNCH_EV(double, double4, false); }
#undef TT_LAUNCH_EV
        int rc2 = launch_check("trace_event_kernel");
        if (rc2) return rc2;
        only_flagged = 1;          // second pass: the general kernel on the deferred rays only
        variant = 2;
    }
#define TT_LAUNCH(TYPE, V4T, VAR)                                                                                 \
    trace_kernel<TYPE, VAR><<<(unsigned)blocks, block, 0, s>>>((const V4T*)grid4_dev, s0_dev, perm_dev, rf_dev,    \
                                                                sf_dev, ray_steps_dev, status_dev, A, only_flagged)
    if (p->dtype == TT_F32) { if (variant == 1) TT_LAUNCH(float, float4, 1); else TT_LAUNCH(float, float4, 0); }
    else { if (variant == 1) TT_LAUNCH(double, double4, 1); else TT_LAUNCH(double, double4, 0); }
#undef TT_LAUNCH
    return launch_check("trace_kernel");
}

extern "C" int tt_trace_aux(const tt_trace_params* p, const tt_aux_params* a, const void* grid4_dev,
                            const void* aux4_dev, const double* s0_dev, long np, const uint32_t* perm_dev,
                            double* rf_dev, double* sf_dev, double* aux_out_dev, unsigned long long* ray_steps_dev,
                            uint8_t* status_dev, tt_stream_t stream) {
    using namespace tt;
    TT_REQUIRE(p && a && grid4_dev && s0_dev && rf_dev && aux_out_dev, "tt_trace_aux: null pointer");
    TT_REQUIRE(np >= 0, "tt_trace_aux: negative ray count");
    TT_REQUIRE(p->dtype == TT_F32 || p->dtype == TT_F64, "tt_trace_aux: dtype must be TT_F32 or TT_F64");
    TT_REQUIRE(p->steps_per_cell >= 1 && p->steps_per_cell <= 1024, "tt_trace_aux: steps_per_cell out of range");
    TT_REQUIRE(p->s_max > 0, "tt_trace_aux: s_max must be > 0");
    TT_REQUIRE(np < (1L << 32) || !perm_dev, "tt_trace_aux: perm is 32-bit; trace in bundles of < 2^32 rays");
    TT_REQUIRE(a->omega > 0 && a->nc > 0, "tt_trace_aux: omega and nc must be > 0");
    TraceArgs A;
    int rc = fill_args(A, p->n_xyz, p->origin_xyz, p->spacing_xyz, p->par);
    if (rc) return rc;
    A.extent = p->extent; A.s_max = p->s_max; A.spc = p->steps_per_cell; A.np = np;
    if (np == 0) return TT_OK;
    AuxArgs AX;
    AX.omega_over_c = a->omega / kC;
    AX.verdet_nc = a->verdet * a->nc;
    const int block = 128;
    const long blocks = (np + block - 1) / block;
    TT_REQUIRE(blocks < (1L << 31), "tt_trace_aux: too many rays for one launch");
    cudaStream_t s = (cudaStream_t)stream;
    TT_REQUIRE(p->variant >= 0 && p->variant <= 4, "tt_trace_aux: unknown kernel variant %d", p->variant);
    int only_flagged = 0;
    if (p->dtype == TT_F32 && status_dev && (p->variant == 0 || p->variant == 3)) {
        // event marching with the passive quantities on board; the gather kernel then redoes the deferred rays
        if (p->steps_per_cell == 1)
            trace_event_kernel_f32x2<true, true><<<(unsigned)blocks, block, 0, s>>>(
                (const float4*)grid4_dev, s0_dev, perm_dev, rf_dev, sf_dev, ray_steps_dev, status_dev, A,
                (const float4*)aux4_dev, aux_out_dev, AX);
        else
            trace_event_kernel_f32x2<false, true><<<(unsigned)blocks, block, 0, s>>>(
                (const float4*)grid4_dev, s0_dev, perm_dev, rf_dev, sf_dev, ray_steps_dev, status_dev, A,
                (const float4*)aux4_dev, aux_out_dev, AX);
        int rc2 = launch_check("trace_event_kernel_f32x2<aux>");
        if (rc2) return rc2;
        only_flagged = 1;
    }
    if (p->dtype == TT_F32)
        trace_kernel<float, 1, true><<<(unsigned)blocks, block, 0, s>>>((const float4*)grid4_dev, s0_dev, perm_dev, rf_dev, sf_dev,
                                                                         ray_steps_dev, status_dev, A, only_flagged,
                                                                         (const float4*)aux4_dev, aux_out_dev, AX);
    else
        trace_kernel<double, 1, true><<<(unsigned)blocks, block, 0, s>>>((const double4*)grid4_dev, s0_dev, perm_dev, rf_dev, sf_dev,
                                                                          ray_steps_dev, status_dev, A, 0,
                                                                          (const double4*)aux4_dev, aux_out_dev, AX);
    return launch_check("trace_kernel<aux>");
}

extern "C" int tt_dndr(const void* grid4_dev, int grid_dtype, const int n_xyz[3], const double origin_xyz[3],
                       const double spacing_xyz[3], int par, const double* pos_dev, long npts,
                       double* out_dev, tt_stream_t stream) {
    using namespace tt;
    TT_REQUIRE(grid4_dev && pos_dev && out_dev, "tt_dndr: null pointer");
    TT_REQUIRE(grid_dtype == TT_F32 || grid_dtype == TT_F64, "tt_dndr: dtype must be TT_F32 or TT_F64");
    TraceArgs A;
    int rc = fill_args(A, n_xyz, origin_xyz, spacing_xyz, par);
    if (rc) return rc;
    A.extent = 0; A.s_max = 0; A.spc = 1; A.np = npts;
    if (npts <= 0) return TT_OK;
    const int block = 256;
    const long blocks = (npts + block - 1) / block;
    cudaStream_t s = (cudaStream_t)stream;
    if (grid_dtype == TT_F32)
        dndr_kernel<float><<<(unsigned)blocks, block, 0, s>>>((const float4*)grid4_dev, A, pos_dev, npts, out_dev);
    else
        dndr_kernel<double><<<(unsigned)blocks, block, 0, s>>>((const double4*)grid4_dev, A, pos_dev, npts, out_dev);
    return launch_check("dndr_kernel");
}
