// TEST HARNESS (not product): replays the launch of calc_dndr_kernel on the HOST -- every block, every thread, the
// three phases of csrc/calc_dndr_tile.cuh separated by what is __syncthreads() on the device -- so that
// tests/test_host_kernels.py can compare the interleaved gradient grid with the reference's arrays without a GPU.
#include "calc_dndr_tile.cuh"

template <typename TIn, typename TOut, int PAR>
static void replay(const TIn* ne, typename tt::Vec4<TOut>::type* grid, const tt::DndrArgs& a) {
    using namespace tt;
    const int gx = (a.n[2] + 31) / 32, gy = (a.n[a.fa[0]] + 31) / 32, gz = a.n[a.third];      // launch_dndr's grid
    DndrScratch<TOut>* S = new DndrScratch<TOut>;
    for (int bz = 0; bz < gz; ++bz)
        for (int by = 0; by < gy; ++by)
            for (int bx = 0; bx < gx; ++bx) {
                if (a.ax[0] != nullptr)
                    for (int ty = 0; ty < 8; ++ty)
                        for (int tx = 0; tx < 32; ++tx) dndr_phase0<TIn, TOut, PAR>(*S, a, bx, by, bz, tx, ty);
                for (int ty = 0; ty < 8; ++ty)
                    for (int tx = 0; tx < 32; ++tx) dndr_phase1<TIn, TOut, PAR>(*S, ne, a, bx, by, bz, tx, ty);
                for (int ty = 0; ty < 8; ++ty)
                    for (int tx = 0; tx < 32; ++tx) dndr_phase2<TIn, TOut, PAR>(*S, grid, a, bx, by, bz, tx, ty);
            }
    delete S;
}

template <typename TIn, typename TOut>
static void by_par(const void* ne, void* grid, const tt::DndrArgs& a) {
    typedef typename tt::Vec4<TOut>::type V4;
    const int par = a.fa[2];
    if (par == 0) replay<TIn, TOut, 0>((const TIn*)ne, (V4*)grid, a);
    else if (par == 1) replay<TIn, TOut, 1>((const TIn*)ne, (V4*)grid, a);
    else replay<TIn, TOut, 2>((const TIn*)ne, (V4*)grid, a);
}

// the argument block exactly as calc_dndr_impl (calc_dndr.cu) fills it
extern "C" int host_calc_dndr(const void* ne, int ne_dtype, const int n_xyz[3], const double spacing_xyz[3],
                              const double* x, const double* y, const double* z, int par, double nc, double ne_max,
                              void* grid4, int grid_dtype) {
    using namespace tt;
    DndrArgs a;
    for (int i = 0; i < 3; ++i) {
        a.n[i] = n_xyz[i];
        a.invh[i] = 1.0 / spacing_xyz[i];
        a.inv2h[i] = 1.0 / (2.0 * spacing_xyz[i]);
    }
    Frame f = frame_of(par);
    for (int i = 0; i < 3; ++i) a.fa[i] = f.a[i];
    a.third = 3 - 2 - a.fa[0];
    const double* axes[3] = {x, y, z};
    for (int i = 0; i < 3; ++i) a.ax[i] = x ? axes[i] : nullptr;
    a.nc = nc; a.inv_nc = 1.0 / nc; a.ne_max = ne_max;
    a.clip_f = (float)(ne_max * nc);
    a.inv_nc_f = (float)(1.0 / nc);
    for (int i = 0; i < 3; ++i) {
        a.k1_f[i] = (float)(-0.5 * a.invh[i] / nc);
        a.k2_f[i] = (float)(-0.5 * a.inv2h[i] / nc);
    }
    if (ne_dtype == TT_F32 && grid_dtype == TT_F32) by_par<float, float>(ne, grid4, a);
    else if (ne_dtype == TT_F64 && grid_dtype == TT_F32) by_par<double, float>(ne, grid4, a);
    else if (ne_dtype == TT_F32 && grid_dtype == TT_F64) by_par<float, double>(ne, grid4, a);
    else by_par<double, double>(ne, grid4, a);
    return 0;
}
