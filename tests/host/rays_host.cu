// TEST HARNESS (not product): tt_init_beam's per-ray generator and tt_sort_rays' Morton key (csrc/rays_one.cuh) on the
// HOST for tests/test_host_kernels.py.
#include "rays_one.cuh"

extern "C" int host_init_beam(long np, long first, unsigned long long seed, double beam_size, double divergence,
                              double extent, int par, double* s0) {
    for (long i = 0; i < np; ++i) tt::init_beam_ray(i, np, first, seed, beam_size, divergence, extent, par, s0);
    return 0;
}

extern "C" int host_morton_keys(const double* s0, long np, int par, const double origin_xyz[3], const double spacing_xyz[3],
                                const int n_xyz[3], unsigned int* keys) {
    using namespace tt;
    const Frame f = frame_of(par);
    const int au = f.a[0], av = f.a[1];
    const double wu = spacing_xyz[au] * (n_xyz[au] - 1), wv = spacing_xyz[av] * (n_xyz[av] - 1);      // as tt_sort_rays
    for (long i = 0; i < np; ++i)
        keys[i] = morton_key(s0, i, np, au, av, origin_xyz[au], origin_xyz[av], 65536.0 / wu, 65536.0 / wv);
    return 0;
}
