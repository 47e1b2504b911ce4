// K1 (see calc_dndr.cu): the three phases of one 32 x 32 tile -- stencil coefficients of a rectilinear block, the
// stencil itself (reads coalesced along z), the transposed write (coalesced along u) -- as host + device functions of
// (block index, thread index) with the block's scratch passed in: __shared__ memory on the device, a plain struct in
// tests/host/calc_dndr_host.cu, which replays the launch block by block on the CPU.
#pragma once
#include "common.cuh"

namespace tt {

template <typename T> struct Vec4;
template <> struct Vec4<float> { typedef float4 type; };
template <> struct Vec4<double> { typedef double4 type; };

struct DndrArgs {
    int n[3];          // nx, ny, nz
    double inv2h[3];   // 1 / (2 h) per axis (central differences)
    double invh[3];    // 1 / h per axis (one-sided differences on the faces)
    double inv_nc;     // 1 / nc
    int fa[3];         // frame: (u, v, w) -> xyz axis
    double nc, ne_max;
    int third;         // the xyz axis that is neither z nor the u axis
    // FP32-in / FP32-out path: clip level in density units and -1/2 * (1/h | 1/2h) / nc per axis
    float clip_f, inv_nc_f, k1_f[3], k2_f[3];
    // rectilinear grids (tt_calc_dndr_axes): node coordinates per xyz axis on the device, else null
    const double* ax[3];
};

// numpy.gradient on a non-uniformly spaced axis (numpy/lib/_function_base_impl.py:1249-1334, edge_order=1):
// interior  a f[i-1] + b f[i] + c f[i+1]  with  a = -dx2 / (dx1 (dx1 + dx2)),  b = (dx2 - dx1) / (dx1 dx2),
// c = dx1 / (dx2 (dx1 + dx2));  faces  (f[1] - f[0]) / dx  and  (f[n-1] - f[n-2]) / dx, kept as differences
// times 1/dx (stored in c resp. b).
TT_HD void axis_coefficients(const double* __restrict__ x, int i, int n, double& a, double& b,
                                                  double& c) {
    if (i == 0) { a = 0.0; c = 1.0 / (x[1] - x[0]); b = -c; return; }
    if (i == n - 1) { c = 0.0; b = 1.0 / (x[n - 1] - x[n - 2]); a = -b; return; }
    const double dx1 = x[i] - x[i - 1], dx2 = x[i + 1] - x[i];
    a = -dx2 / (dx1 * (dx1 + dx2));
    b = (dx2 - dx1) / (dx1 * dx2);
    c = dx1 / (dx2 * (dx1 + dx2));
}

// ne/nc clipped at ne_max.  Multiplying by 1/nc instead of dividing differs from the reference's
// quotient by at most 1 ulp and keeps the kernel bandwidth-bound (an FP64 division costs ~30 instructions
// and this is evaluated for the voxel and its 6 neighbours).
template <typename TIn>
TT_HD double ne_over_nc(const TIn* __restrict__ ne, size_t idx, double inv_nc,
                                             double ne_max) {
    double v = (double)ne[idx] * inv_nc; // particle_tracker.py:230
    return v > ne_max ? ne_max : v;      // :231 (NaN stays NaN, as with numpy's mask)
}

// numpy.gradient along one axis at index i of n: central inside, one-sided at the two faces
// (edge_order=1, numpy/lib/_function_base_impl.py:1294-1334), uniform spacing h.
template <typename TIn>
TT_HD double axis_gradient(const TIn* __restrict__ ne, size_t idx, size_t stride,
                                                int i, int n, double invh, double inv2h, double centre,
                                                double nc, double ne_max) {
    if (n == 1) return 0.0;
    if (i == 0) return (ne_over_nc(ne, idx + stride, nc, ne_max) - centre) * invh;
    if (i == n - 1) return (centre - ne_over_nc(ne, idx - stride, nc, ne_max)) * invh;
    return (ne_over_nc(ne, idx + stride, nc, ne_max) - ne_over_nc(ne, idx - stride, nc, ne_max)) * inv2h;
}

template <typename TIn>
TT_HD double axis_gradient_nu(const TIn* __restrict__ ne, size_t idx, size_t stride, int i, int n,
                                                   double a, double b, double c, double centre, double inv_nc,
                                                   double ne_max) {
    if (i == 0) return (ne_over_nc(ne, idx + stride, inv_nc, ne_max) - centre) * c;
    if (i == n - 1) return (centre - ne_over_nc(ne, idx - stride, inv_nc, ne_max)) * b;
    return add_rn(add_rn(mul_rn(a, ne_over_nc(ne, idx - stride, inv_nc, ne_max)), mul_rn(b, centre)),
                     mul_rn(c, ne_over_nc(ne, idx + stride, inv_nc, ne_max)));
}

// what one block keeps in shared memory
template <typename TOut>
struct DndrScratch {
    typename Vec4<TOut>::type tile[32][33];
    double coef[3][3][32];        // [0: z, 1: u, 2: third][a, b, c][node in tile]   (rectilinear axes only)
};

// phase 0 (rectilinear axes): the block's 32 z nodes, 32 u nodes and its one node of the third axis
template <typename TIn, typename TOut, int PAR>
TT_HD void dndr_phase0(DndrScratch<TOut>& S, const DndrArgs& a, int bx, int by, int bz, int tx, int ty) {
    typedef typename Vec4<TOut>::type V4;
    // frame (u, v, w) -> xyz axis, compile-time so that the index arrays stay in registers
    constexpr int F0 = PAR == 0 ? 1 : 0, F1 = PAR == 2 ? 1 : 2, F2 = PAR;
    constexpr int ua = F0;                  // xyz axis that is fastest in the output
    constexpr int ta = 1 - F0;              // the axis that is neither z nor u
    const int z0 = bx * 32, u0 = by * 32, t = bz;
    const int nx = a.n[0], ny = a.n[1], nz = a.n[2];
    const size_t sx = (size_t)ny * nz, sy = (size_t)nz;
    const bool rect = a.ax[0] != nullptr;
    (void)F1; (void)F2; (void)ta; (void)nx; (void)sx; (void)sy; (void)z0; (void)u0; (void)t; (void)rect;
    auto& coef = S.coef;
    if (rect) {
        const int tid = ty * 32 + tx;
        if (tid < 65) {
            const int which = tid < 32 ? 0 : (tid < 64 ? 1 : 2);
            const int axis = which == 0 ? 2 : (which == 1 ? ua : ta);
            const int j = tid & 31;
            const int i = which == 0 ? z0 + j : (which == 1 ? u0 + j : t);
            if (i < a.n[axis]) axis_coefficients(a.ax[axis], i, a.n[axis], coef[which][0][j], coef[which][1][j], coef[which][2][j]);
        }
    }

}

// phase 1: the stencil, one z column of 4 voxels per thread
template <typename TIn, typename TOut, int PAR>
TT_HD void dndr_phase1(DndrScratch<TOut>& S, const TIn* __restrict__ ne, const DndrArgs& a, int bx, int by, int bz,
                       int tx, int ty) {
    typedef typename Vec4<TOut>::type V4;
    // frame (u, v, w) -> xyz axis, compile-time so that the index arrays stay in registers
    constexpr int F0 = PAR == 0 ? 1 : 0, F1 = PAR == 2 ? 1 : 2, F2 = PAR;
    constexpr int ua = F0;                  // xyz axis that is fastest in the output
    constexpr int ta = 1 - F0;              // the axis that is neither z nor u
    const int z0 = bx * 32, u0 = by * 32, t = bz;
    const int nx = a.n[0], ny = a.n[1], nz = a.n[2];
    const size_t sx = (size_t)ny * nz, sy = (size_t)nz;
    const bool rect = a.ax[0] != nullptr;
    (void)F1; (void)F2; (void)ta; (void)nx; (void)sx; (void)sy; (void)z0; (void)u0; (void)t; (void)rect;
    auto& coef = S.coef;
    auto& tile = S.tile;
    (void)coef;
    // phase 1: compute, coalesced along z
    const int iz = z0 + tx;
#pragma unroll
    for (int r = 0; r < 32; r += 8) {
        const int iu = u0 + r + ty;
        if (iz < nz && iu < a.n[ua]) {
            int i3[3];
            i3[2] = iz; i3[ua] = iu; i3[ta] = t;
            const size_t idx = (size_t)i3[0] * sx + (size_t)i3[1] * sy + i3[2];
            V4 o;
            if (rect) {
                const double c = ne_over_nc(ne, idx, a.inv_nc, a.ne_max);
                const size_t st[3] = {sx, sy, 1};
                double g[3];
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    const int which = ax == 2 ? 0 : (ax == ua ? 1 : 2);
                    const int j = which == 0 ? (int)tx : (which == 1 ? r + (int)ty : 0);
                    g[ax] = -0.5 * axis_gradient_nu(ne, idx, st[ax], i3[ax], a.n[ax], coef[which][0][j], coef[which][1][j],
                                                    coef[which][2][j], c, a.inv_nc, a.ne_max);
                }
                o.x = (TOut)g[F0]; o.y = (TOut)g[F1]; o.z = (TOut)g[F2]; o.w = (TOut)c;
            } else if constexpr (sizeof(TIn) == 4 && sizeof(TOut) == 4) {
                // FP32 cube in, FP32 grid out: the difference of two neighbouring FP32 densities is exact in
                // FP32 (Sterbenz), so nothing is gained by FP64 here and the conversions would make the
                // kernel XU-bound; differences first, one multiplication by -1/2 / (h nc) afterwards
                const float clipv = a.clip_f;
                auto L = [&](size_t id) { const float v = (float)ne[id]; return v > clipv ? clipv : v; };
                const float c = L(idx);
                float g[3];
                const size_t st[3] = {sx, sy, 1};
                const int nn[3] = {nx, ny, nz};
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    const int i = i3[ax];
                    g[ax] = i == 0 ? (L(idx + st[ax]) - c) * a.k1_f[ax]
                          : i == nn[ax] - 1 ? (c - L(idx - st[ax])) * a.k1_f[ax]
                                            : (L(idx + st[ax]) - L(idx - st[ax])) * a.k2_f[ax];
                }
                o.x = g[F0]; o.y = g[F1]; o.z = g[F2]; o.w = c * a.inv_nc_f;
            } else {
                const double c = ne_over_nc(ne, idx, a.inv_nc, a.ne_max);
                double g[3];
                g[0] = -0.5 * axis_gradient(ne, idx, sx, i3[0], nx, a.invh[0], a.inv2h[0], c, a.inv_nc, a.ne_max);
                g[1] = -0.5 * axis_gradient(ne, idx, sy, i3[1], ny, a.invh[1], a.inv2h[1], c, a.inv_nc, a.ne_max);
                g[2] = -0.5 * axis_gradient(ne, idx, 1, i3[2], nz, a.invh[2], a.inv2h[2], c, a.inv_nc, a.ne_max);
                o.x = (TOut)g[F0]; o.y = (TOut)g[F1]; o.z = (TOut)g[F2]; o.w = (TOut)c;
            }
            tile[r + ty][tx] = o;
        }
    }
}

// phase 2: the transposed write
template <typename TIn, typename TOut, int PAR>
TT_HD void dndr_phase2(const DndrScratch<TOut>& S, typename Vec4<TOut>::type* __restrict__ grid, const DndrArgs& a,
                       int bx, int by, int bz, int tx, int ty) {
    typedef typename Vec4<TOut>::type V4;
    // frame (u, v, w) -> xyz axis, compile-time so that the index arrays stay in registers
    constexpr int F0 = PAR == 0 ? 1 : 0, F1 = PAR == 2 ? 1 : 2, F2 = PAR;
    constexpr int ua = F0;                  // xyz axis that is fastest in the output
    constexpr int ta = 1 - F0;              // the axis that is neither z nor u
    const int z0 = bx * 32, u0 = by * 32, t = bz;
    const int nx = a.n[0], ny = a.n[1], nz = a.n[2];
    const size_t sx = (size_t)ny * nz, sy = (size_t)nz;
    const bool rect = a.ax[0] != nullptr;
    (void)F1; (void)F2; (void)ta; (void)nx; (void)sx; (void)sy; (void)z0; (void)u0; (void)t; (void)rect;
    auto& tile = S.tile;
    // phase 2: write, coalesced along u
    const int ou = u0 + tx;
    const int nu = a.n[F0], nv = a.n[F1];
#pragma unroll
    for (int r = 0; r < 32; r += 8) {
        const int oz = z0 + r + ty;
        if (ou < nu && oz < nz) {
            int i3[3];
            i3[2] = oz; i3[ua] = ou; i3[ta] = t;
            const int iv = i3[F1], iw = i3[F2];
            grid[((size_t)iw * nv + iv) * nu + ou] = tile[tx][r + ty];
        }
    }
}

}  // namespace tt
