// K1: gradient-grid stencil.  Replaces ElectronCube.calc_dndr (particle_tracker.py:220-241):
// three numpy.gradient passes + three RegularGridInterpolator objects become ONE pass that
// writes an interleaved 4-vector grid (g_u, g_v, g_w, ne/nc) in ray-frame order, so that a
// trilinear corner is a single 16-byte (FP32) or 32-byte (FP64) load in the trace kernel.
//
// Memory behaviour: input ne[ix][iy][iz] is z-fastest; output grid[iw][iv][iu] is u-fastest
// and u is never z, so every block transposes a 32(z) x 32(u) tile through shared memory:
// reads are coalesced along z, writes along u.  Algorithmic traffic: read sizeof(ne) per voxel
// (neighbour reads hit L1/L2), write 16 or 32 B per voxel.
#include "calc_dndr_tile.cuh"       // Vec4, DndrArgs, the stencil and the three tile phases (host + device)

namespace tt {

template <typename TIn, typename TOut, int PAR>
__global__ void __launch_bounds__(256) calc_dndr_kernel(const TIn* __restrict__ ne,
                                                        typename Vec4<TOut>::type* __restrict__ grid,
                                                        DndrArgs a) {
    __shared__ DndrScratch<TOut> S;          // calc_dndr_tile.cuh
    const int bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z, tx = threadIdx.x, ty = threadIdx.y;
    if (a.ax[0] != nullptr) {                // rectilinear axes: stencil coefficients of this block's nodes
        dndr_phase0<TIn, TOut, PAR>(S, a, bx, by, bz, tx, ty);
        __syncthreads();
    }
    dndr_phase1<TIn, TOut, PAR>(S, ne, a, bx, by, bz, tx, ty);      // compute, coalesced along z
    __syncthreads();
    dndr_phase2<TIn, TOut, PAR>(S, grid, a, bx, by, bz, tx, ty);    // write, coalesced along u
}

template <typename TIn, typename TOut>
static int launch_dndr(const void* ne, void* grid, const DndrArgs& a, cudaStream_t s) {
    dim3 block(32, 8);
    dim3 gridDim((a.n[2] + 31) / 32, (a.n[a.fa[0]] + 31) / 32, a.n[a.third]);
    typedef typename Vec4<TOut>::type V4;
    const int par = a.fa[2];
    if (par == 0) calc_dndr_kernel<TIn, TOut, 0><<<gridDim, block, 0, s>>>((const TIn*)ne, (V4*)grid, a);
    else if (par == 1) calc_dndr_kernel<TIn, TOut, 1><<<gridDim, block, 0, s>>>((const TIn*)ne, (V4*)grid, a);
    else calc_dndr_kernel<TIn, TOut, 2><<<gridDim, block, 0, s>>>((const TIn*)ne, (V4*)grid, a);
    return launch_check("calc_dndr_kernel");
}

}  // namespace tt

static int calc_dndr_impl(const void* ne_dev, int ne_dtype, const int n_xyz[3], const double spacing_xyz[3],
                          const double* const axes_dev[3], int par, double nc, double ne_max, void* grid4_dev,
                          int grid_dtype, tt_stream_t stream) {
    using namespace tt;
    TT_REQUIRE(ne_dev && grid4_dev && n_xyz && spacing_xyz, "tt_calc_dndr: null pointer");
    TT_REQUIRE(par >= 0 && par <= 2, "tt_calc_dndr: par must be 0, 1 or 2 (got %d)", par);
    TT_REQUIRE((ne_dtype == TT_F32 || ne_dtype == TT_F64) && (grid_dtype == TT_F32 || grid_dtype == TT_F64),
               "tt_calc_dndr: dtype must be TT_F32 or TT_F64");
    DndrArgs a;
    for (int i = 0; i < 3; ++i) {
        TT_REQUIRE(n_xyz[i] >= 2, "tt_calc_dndr: every axis needs >= 2 points (axis %d has %d)", i, n_xyz[i]);
        TT_REQUIRE(spacing_xyz[i] > 0, "tt_calc_dndr: spacing must be > 0");
        a.n[i] = n_xyz[i];
        a.invh[i] = 1.0 / spacing_xyz[i];
        a.inv2h[i] = 1.0 / (2.0 * spacing_xyz[i]);
    }
    TT_REQUIRE(n_xyz[0] <= 65535 * 32 && n_xyz[1] <= 65535 * 32, "tt_calc_dndr: cube too large");
    TT_REQUIRE(nc > 0, "tt_calc_dndr: critical density must be > 0");
    Frame f = frame_of(par);
    for (int i = 0; i < 3; ++i) a.fa[i] = f.a[i];
    a.third = 3 - 2 - a.fa[0];   // axes are {0,1,2}; z = 2 and u = fa[0] are taken
    for (int i = 0; i < 3; ++i) a.ax[i] = axes_dev ? axes_dev[i] : nullptr;
    a.nc = nc;
    a.inv_nc = 1.0 / nc;
    a.ne_max = ne_max;
    a.clip_f = (float)(ne_max * nc);
    a.inv_nc_f = (float)(1.0 / nc);
    for (int i = 0; i < 3; ++i) {
        a.k1_f[i] = (float)(-0.5 * a.invh[i] / nc);
        a.k2_f[i] = (float)(-0.5 * a.inv2h[i] / nc);
    }
    TT_REQUIRE(a.n[a.third] <= 65535, "tt_calc_dndr: axis too long for grid.z");
    cudaStream_t s = (cudaStream_t)stream;
    if (ne_dtype == TT_F32 && grid_dtype == TT_F32) return launch_dndr<float, float>(ne_dev, grid4_dev, a, s);
    if (ne_dtype == TT_F64 && grid_dtype == TT_F32) return launch_dndr<double, float>(ne_dev, grid4_dev, a, s);
    if (ne_dtype == TT_F32 && grid_dtype == TT_F64) return launch_dndr<float, double>(ne_dev, grid4_dev, a, s);
    return launch_dndr<double, double>(ne_dev, grid4_dev, a, s);
}

extern "C" int tt_calc_dndr(const void* ne_dev, int ne_dtype, const int n_xyz[3],
                            const double spacing_xyz[3], int par, double nc, double ne_max,
                            void* grid4_dev, int grid_dtype, tt_stream_t stream) {
    return calc_dndr_impl(ne_dev, ne_dtype, n_xyz, spacing_xyz, nullptr, par, nc, ne_max, grid4_dev, grid_dtype, stream);
}

extern "C" int tt_calc_dndr_axes(const void* ne_dev, int ne_dtype, const int n_xyz[3], const double* x_dev,
                                 const double* y_dev, const double* z_dev, int par, double nc, double ne_max,
                                 void* grid4_dev, int grid_dtype, tt_stream_t stream) {
    TT_REQUIRE(x_dev && y_dev && z_dev, "tt_calc_dndr_axes: null axis pointer");
    const double* const axes[3] = {x_dev, y_dev, z_dev};
    const double unit[3] = {1.0, 1.0, 1.0};      // unused by the rectilinear branch
    return calc_dndr_impl(ne_dev, ne_dtype, n_xyz, unit, axes, par, nc, ne_max, grid4_dev, grid_dtype, stream);
}
