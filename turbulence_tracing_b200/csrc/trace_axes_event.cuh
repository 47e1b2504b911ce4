// Event marching on RECTILINEAR grids: the per-ray body, compiled for the device (trace_axes.cu) AND for the
// host (tests/host/axes_event_host.cu runs the very same source on the CPU against the C oracle).
#pragma once
#include "trace_common.cuh"

namespace tt {

struct AxesArgs {
    int n[3];                 // nu, nv, nw
    const double* ax[3];      // node coordinates per frame axis (device)
    int fa[3];                // frame axis -> xyz row
    double extent, s_max;
    int spc;
    long np;
};

// cell c in [0, n-2] with ax[c] <= x < ax[c+1] (clamped at both ends) by bisection
TT_HD int find_cell(const double* __restrict__ ax, int n, double x) {
    int lo = 0, hi = n - 1;                   // invariant: ax[lo] <= x < ax[hi] (after clamping)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= ldg_f64(ax + mid)) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- event marching on a rectilinear grid (the speed path for stretched meshes) ------------------------
// The scheme of trace_event.cu carries over unchanged because it works in INDEX space: inside one cell the
// reference's interpolant (RegularGridInterpolator: weights from the normalised distances (x - x_i) /
// (x_{i+1} - x_i), scipy/interpolate/_rgi.py) is a trilinear polynomial of the cell fractions whatever the
// cell's size, and the equations of motion in the fractions are
//      dU/dW = (h_w/h_u) d_u/d_w,   d(d)/dW = h_w g / d_w,   ds/dW = h_w / d_w
// with the sizes (h_u, h_v, h_w) of the CURRENT cell -- three values refreshed from the node tables whenever
// the ray is relabelled into a neighbouring cell (plane arrival: h_w; u / v face: h_u / h_v).  Every RK4 step
// lies inside one cell (no kink of the field is straddled), no stage needs a cell search, and the eight
// corners are loaded once per cell.  State and arithmetic have the grid's element type (FP32 for a float4 grid:
// positions are (cell, fraction) pairs, so the resolution is 6e-8 of a cell; launch and exit go through FP64);
// the FP64 gather kernel of trace_axes.cu stays the second pass for everything unusual (launched outside the cube beside
// the entry face, steep / backward, side exit, possible time cap, non-finite): those rays are flagged
// TT_RAY_DEFERRED and redone from s0.
// 1/h_u, 1/h_v of the current cell are kept as state (refreshed at u / v crossings only): a plane arrival costs two
// multiplications instead of two divisions (measured on B200, stretched 257^3 mesh, 1e7 rays: float4 grid 34.2 -> 31.0 ms,
// double4 grid 62.1 -> 55.9 ms; same accuracy on the fixtures, profiles/r02_ab_lean_axesrcp.txt)
// Returns the (sub-)plane arrivals of this ray (0 if deferred).
template <typename T>
TT_HD unsigned axes_event_ray(const typename GridT<T>::V4* __restrict__ grid, const double* __restrict__ s0, long ray,
                              double* __restrict__ rf, double* __restrict__ sf, uint8_t* __restrict__ status,
                              const AxesArgs& A, bool& deferred) {
    typedef typename GridT<T>::V4 V4;
    typedef T R;              // arithmetic follows the grid: FP32 state for a float4 grid (positions stay (cell, fraction))
    unsigned steps = 0;
    const int nu = A.n[0], nv = A.n[1], nw = A.n[2];
    const size_t plane = (size_t)nu * nv;
    const double *axu = A.ax[0], *axv = A.ax[1], *axw = A.ax[2];
    // ---- prologue ---------------------------------------------------------------------------------
    double P[3], D[3], lo[3], hi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        P[k] = s0[(size_t)A.fa[k] * A.np + ray];
        D[k] = s0[(size_t)(3 + A.fa[k]) * A.np + ray] * (1.0 / kC);
        lo[k] = ldg_f64(A.ax[k]); hi[k] = ldg_f64(A.ax[k] + A.n[k] - 1);
    }
    double s_pre = 0.0;                       // free flight to the entry face (the field is zero out there)
    if (P[2] < lo[2] && D[2] > TT_MARCH_MIN_DW) {
        s_pre = (lo[2] - P[2]) / D[2];
        P[0] = fma(D[0], s_pre, P[0]);
        P[1] = fma(D[1], s_pre, P[1]);
        P[2] = lo[2];
    }
    bool fast = P[0] >= lo[0] && P[0] <= hi[0] && P[1] >= lo[1] && P[1] <= hi[1] && P[2] >= lo[2] && P[2] <= hi[2] &&
                D[2] > TT_MARCH_MIN_DW;
    // the path-time cap must be out of reach while marching (d_w > 0.75 throughout)
    fast = fast && (hi[2] - P[2]) <= TT_MARCH_MIN_DW * (A.s_max - s_pre);
    int cu = 0, cv = 0, k = nw - 1;
    R tu = 0, tv = 0, fw = 0, hu = 1, hv = 1, hw = 1;
    double xu0 = 0, xv0 = 0;                  // lower node of the current (u, v) cell
    if (fast) {
        cu = find_cell(axu, nu, P[0]); xu0 = ldg_f64(axu + cu); hu = ldg_f64(axu + cu + 1) - xu0; tu = (P[0] - xu0) / hu;
        cv = find_cell(axv, nv, P[1]); xv0 = ldg_f64(axv + cv); hv = ldg_f64(axv + cv + 1) - xv0; tv = (P[1] - xv0) / hv;
        if (P[2] < hi[2]) {
            k = find_cell(axw, nw, P[2]);
            const double w0 = ldg_f64(axw + k);
            hw = ldg_f64(axw + k + 1) - w0; fw = (P[2] - w0) / hw;
        }
    }
    R du = D[0], dv = D[1], dw = D[2], s = 0;
    R ihu = R(1) / hu, ihv = R(1) / hv;
    R ru = hw * ihu, rv = hw * ihv;
    const int spc = A.spc;
    const R hsub = R(1) / (R)spc;
    int j = (int)(fw * (R)spc);               // current sub-plane interval of the w-cell
    j = j > spc - 1 ? spc - 1 : j;

    if (fast && k < nw - 1) {
        const V4* p = grid + ((size_t)k * plane + (size_t)cv * nu + cu);
        Tri<R> qx, qy, qz;
        auto load_cell = [&]() {
            V4 c00 = GridT<T>::ld(p), c10 = GridT<T>::ld(p + 1), c01 = GridT<T>::ld(p + nu), c11 = GridT<T>::ld(p + nu + 1);
            const V4* p1 = p + plane;
            V4 e00 = GridT<T>::ld(p1), e10 = GridT<T>::ld(p1 + 1), e01 = GridT<T>::ld(p1 + nu), e11 = GridT<T>::ld(p1 + nu + 1);
            tri_set<R>(qx, c00.x, c10.x, c01.x, c11.x, e00.x, e10.x, e01.x, e11.x);
            tri_set<R>(qy, c00.y, c10.y, c01.y, c11.y, e00.y, e10.y, e01.y, e11.y);
            tri_set<R>(qz, c00.z, c10.z, c01.z, c11.z, e00.z, e10.z, e01.z, e11.z);
        };
        load_cell();
        while (true) {
            // ---- stage 1 and the length of this step (in w-cell fractions) -------------------------
            R q = trcp<R>(dw), hq = hw * q;
            bool ok = dw > R(TT_MARCH_MIN_DW);
            const R aU = ru * du * q, aV = rv * dv * q;
            const R adu = bil_eval<R>(tri_at<R>(qx, fw), tu, tv) * hq, adv = bil_eval<R>(tri_at<R>(qy, fw), tu, tv) * hq,
                    adw = bil_eval<R>(tri_at<R>(qz, fw), tu, tv) * hq, as = hq;
            const R fw_t = (j + 1 == spc) ? R(1) : (R)(j + 1) * hsub;
            R h = fw_t - fw;
            int cross = 0;                         // +-1: u face, +-2: v face
            {
                const R pu = fma(h, aU, tu), pv = fma(h, aV, tv);
                if (pu > R(1) || pu < R(0) || pv > R(1) || pv < R(0)) {
                    R lu = R(2), lv = R(2);
                    if (aU > R(0)) lu = (R(1) - tu) / (h * aU); else if (aU < R(0)) lu = -tu / (h * aU);
                    if (aV > R(0)) lv = (R(1) - tv) / (h * aV); else if (aV < R(0)) lv = -tv / (h * aV);
                    const R lam = fmin(lu, lv);
                    if (lam < R(1)) {
                        cross = lu <= lv ? (aU > R(0) ? 1 : -1) : (aV > R(0) ? 2 : -2);
                        h *= lam > R(0) ? lam : R(0);
                    }
                }
            }
            const R half = R(0.5) * h;
            // ---- stages 2-4 -------------------------------------------------------------------------
            R su = fma(half, aU, tu), sv = fma(half, aV, tv), sw = fw + half;
            R du2 = fma(half, adu, du), dv2 = fma(half, adv, dv), dw2 = fma(half, adw, dw);
            q = trcp<R>(dw2); hq = hw * q; ok = ok && dw2 > R(0);
            const R bU = ru * du2 * q, bV = rv * dv2 * q;
            const Bil<R> mx = tri_at<R>(qx, sw), my = tri_at<R>(qy, sw), mz = tri_at<R>(qz, sw);   // stages 2 and 3
            const R bdu = bil_eval<R>(mx, su, sv) * hq, bdv = bil_eval<R>(my, su, sv) * hq,
                    bdw = bil_eval<R>(mz, su, sv) * hq, bs = hq;
            su = fma(half, bU, tu); sv = fma(half, bV, tv);
            du2 = fma(half, bdu, du); dv2 = fma(half, bdv, dv); dw2 = fma(half, bdw, dw);
            q = trcp<R>(dw2); hq = hw * q; ok = ok && dw2 > R(0);
            const R cU = ru * du2 * q, cV = rv * dv2 * q;
            const R cdu = bil_eval<R>(mx, su, sv) * hq, cdv = bil_eval<R>(my, su, sv) * hq,
                    cdw = bil_eval<R>(mz, su, sv) * hq, cs = hq;
            su = fma(h, cU, tu); sv = fma(h, cV, tv); sw = fw + h;
            du2 = fma(h, cdu, du); dv2 = fma(h, cdv, dv); dw2 = fma(h, cdw, dw);
            q = trcp<R>(dw2); hq = hw * q; ok = ok && dw2 > R(0);
            const R eU = ru * du2 * q, eV = rv * dv2 * q;
            const R edu = bil_eval<R>(tri_at<R>(qx, sw), su, sv) * hq, edv = bil_eval<R>(tri_at<R>(qy, sw), su, sv) * hq,
                    edw = bil_eval<R>(tri_at<R>(qz, sw), su, sv) * hq, es = hq;
            const R h6 = h * R(1.0 / 6.0);
            tu = fma(h6, aU + R(2) * (bU + cU) + eU, tu);
            tv = fma(h6, aV + R(2) * (bV + cV) + eV, tv);
            du = fma(h6, adu + R(2) * (bdu + cdu) + edu, du);
            dv = fma(h6, adv + R(2) * (bdv + cdv) + edv, dv);
            dw = fma(h6, adw + R(2) * (bdw + cdw) + edw, dw);
            s = fma(h6, as + R(2) * (bs + cs) + es, s);
            if (!(ok && dw > R(TT_MARCH_MIN_DW))) { fast = false; break; }   // steep / turning / NaN
            if (cross == 0) {
                // ---- reached the next (sub-)plane ---------------------------------------------------
                ++steps;
                fw = fw_t;
                if (++j == spc) {
                    j = 0; fw = R(0);
                    if (++k >= nw - 1) break;                                 // far face: done
                    p += plane;
                    hw = ldg_f64(axw + k + 1) - ldg_f64(axw + k);
                    ru = hw * ihu; rv = hw * ihv;
                    load_cell();
                }
            } else {
                // ---- reached a u / v cell face inside the w-cell: relabel and reload ----------------
                fw += h;
                if (cross == 1) { ++cu; p += 1; } else if (cross == -1) { --cu; p -= 1; }
                else if (cross == 2) { ++cv; p += nu; } else { --cv; p -= nu; }
                if (cu < 0 || cu > nu - 2 || cv < 0 || cv > nv - 2) { fast = false; break; }   // side exit
                // the fraction was measured in the old cell's size: re-express it in the new cell's
                if (cross == 1 || cross == -1) {
                    xu0 = ldg_f64(axu + cu);
                    const R hn = ldg_f64(axu + cu + 1) - xu0, sc = hu / hn;
                    tu = cross == 1 ? (tu - R(1)) * sc : fma(tu, sc, R(1));
                    hu = hn;
                    ihu = R(1) / hu; ru = hw * ihu;
                } else {
                    xv0 = ldg_f64(axv + cv);
                    const R hn = ldg_f64(axv + cv + 1) - xv0, sc = hv / hn;
                    tv = cross == 2 ? (tv - R(1)) * sc : fma(tv, sc, R(1));
                    hv = hn;
                    ihv = R(1) / hv; rv = hw * ihv;
                }
                load_cell();
            }
        }
    }
    if (!fast) {
        status[ray] = TT_RAY_DEFERRED;          // the gather kernel redoes this ray from s0
        deferred = true;
        steps = 0;
    } else {
        // ---- epilogue: ray_at_exit (particle_tracker.py:345-380) and the state at time T ---------
        const double Pu = fma((double)tu, ldg_f64(axu + cu + 1) - xu0, xu0), Pv = fma((double)tv, ldg_f64(axv + cv + 1) - xv0, xv0);
        const double Pw = hi[2];
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
