// Launch-ray generation on device and Morton ordering of rays.
//   tt_init_beam : ElectronCube.init_beam (particle_tracker.py:258-310) with a counter-based RNG
//   tt_sort_rays : Z-order permutation of the rays by transverse launch position, so that the 32
//                  rays of a warp walk through the same / adjacent grid cells (gather locality).
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace tt {

static constexpr double kC = 299792458.0;
static constexpr double kPi = 3.14159265358979323846;

__global__ void __launch_bounds__(256) init_beam_kernel(long np, long first, uint64_t seed, double beam_size,
                                                        double divergence, double extent, int par,
                                                        double* __restrict__ s0) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const uint64_t id = (uint64_t)(first + i);
    // draws of particle_tracker.py:273-278: t, u1, u2, phi, chi (chi ~ N(0,1) via Box-Muller)
    Philox a = philox4x32_10(id, 0, seed), b = philox4x32_10(id, 1, seed), c = philox4x32_10(id, 2, seed);
    const double t = 2.0 * kPi * u01(a.c[0], a.c[1]);
    double u = u01(a.c[2], a.c[3]) + u01(b.c[0], b.c[1]);
    if (u > 1.0) u = 2.0 - u;
    const double phi = kPi * u01(b.c[2], b.c[3]);
    const double chi = divergence * sqrt(-2.0 * log(u01(c.c[0], c.c[1]))) * cos(2.0 * kPi * u01(c.c[2], c.c[3]));
    double st, ct, sp, cp, sc, cc;
    sincos(t, &st, &ct);
    sincos(phi, &sp, &cp);
    sincos(chi, &sc, &cc);
    const double p1 = beam_size * u * ct, p2 = beam_size * u * st;
    const double vpar = kC * cc, v1 = kC * sc * cp, v2 = kC * sc * sp;
    // transverse axes (t1, t2) and the launch plane; 'x' launches at +extent (quirk of :280-289)
    const Frame f = frame_of(par);
    const double ppar = par == 0 ? extent : -extent;
    s0[(size_t)f.a[0] * np + i] = p1;
    s0[(size_t)f.a[1] * np + i] = p2;
    s0[(size_t)f.a[2] * np + i] = ppar;
    s0[(size_t)(3 + f.a[0]) * np + i] = v1;
    s0[(size_t)(3 + f.a[1]) * np + i] = v2;
    s0[(size_t)(3 + f.a[2]) * np + i] = vpar;
}

__device__ __forceinline__ uint32_t spread16(uint32_t v) {
    v &= 0xFFFFu;
    v = (v | (v << 8)) & 0x00FF00FFu;
    v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

__global__ void __launch_bounds__(256) morton_key_kernel(const double* __restrict__ s0, long np, int au, int av,
                                                         double ou, double ov, double su, double sv,
                                                         uint32_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    // 16 bits per transverse axis over the cube's width (su, sv = 65536 / width)
    double qu = (s0[(size_t)au * np + i] - ou) * su, qv = (s0[(size_t)av * np + i] - ov) * sv;
    qu = fmin(fmax(qu, 0.0), 65535.0);
    qv = fmin(fmax(qv, 0.0), 65535.0);
    if (!(qu == qu)) qu = 0.0;
    if (!(qv == qv)) qv = 0.0;
    keys[i] = spread16((uint32_t)qu) | (spread16((uint32_t)qv) << 1);
    idx[i] = (uint32_t)i;
}

static inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

static int sort_temp_bytes(long np, size_t* bytes) {
    size_t temp = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)np);
    if (e != cudaSuccess) return cuda_fail(e, "cub::DeviceRadixSort::SortPairs(size query)");
    *bytes = temp;
    return TT_OK;
}

}  // namespace tt

extern "C" int tt_init_beam(long np, long first_ray, uint64_t seed, double beam_size, double divergence,
                            double extent, int par, double* s0_dev, tt_stream_t stream) {
    using namespace tt;
    TT_REQUIRE(s0_dev, "tt_init_beam: null pointer");
    TT_REQUIRE(np >= 0 && first_ray >= 0, "tt_init_beam: negative count");
    TT_REQUIRE(par >= 0 && par <= 2, "tt_init_beam: par must be 0, 1 or 2");
    if (np == 0) return TT_OK;
    const long blocks = (np + 255) / 256;
    TT_REQUIRE(blocks < (1L << 31), "tt_init_beam: too many rays for one launch");
    init_beam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(np, first_ray, seed, beam_size, divergence,
                                                                         extent, par, s0_dev);
    return launch_check("init_beam_kernel");
}

extern "C" int tt_sort_rays_workspace(long np, size_t* bytes) {
    using namespace tt;
    TT_REQUIRE(bytes, "tt_sort_rays_workspace: null pointer");
    TT_REQUIRE(np >= 0 && np < (1L << 31), "tt_sort_rays: ray count must be < 2^31 per bundle");
    size_t temp = 0;
    int rc = sort_temp_bytes(np > 0 ? np : 1, &temp);
    if (rc) return rc;
    // keys_in, keys_out, idx_in + cub temp
    *bytes = 3 * align256((size_t)(np > 0 ? np : 1) * sizeof(uint32_t)) + align256(temp);
    return TT_OK;
}

extern "C" int tt_sort_rays(const double* s0_dev, long np, int par, const double origin_xyz[3],
                            const double spacing_xyz[3], const int n_xyz[3], uint32_t* perm_dev,
                            void* workspace_dev, size_t workspace_bytes, tt_stream_t stream) {
    using namespace tt;
    TT_REQUIRE(s0_dev && perm_dev && workspace_dev && origin_xyz && spacing_xyz && n_xyz, "tt_sort_rays: null pointer");
    TT_REQUIRE(par >= 0 && par <= 2, "tt_sort_rays: par must be 0, 1 or 2");
    size_t need = 0;
    int rc = tt_sort_rays_workspace(np, &need);
    if (rc) return rc;
    TT_REQUIRE(workspace_bytes >= need, "tt_sort_rays: workspace too small (%zu < %zu)", workspace_bytes, need);
    if (np == 0) return TT_OK;
    const size_t nb = align256((size_t)np * sizeof(uint32_t));
    char* w = (char*)workspace_dev;
    uint32_t* keys_in = (uint32_t*)w;
    uint32_t* keys_out = (uint32_t*)(w + nb);
    uint32_t* idx_in = (uint32_t*)(w + 2 * nb);
    void* temp = w + 3 * nb;
    size_t temp_bytes = workspace_bytes - 3 * nb;
    const Frame f = frame_of(par);
    const int au = f.a[0], av = f.a[1];
    const double wu = spacing_xyz[au] * (n_xyz[au] - 1), wv = spacing_xyz[av] * (n_xyz[av] - 1);
    TT_REQUIRE(wu > 0 && wv > 0, "tt_sort_rays: degenerate cube");
    cudaStream_t s = (cudaStream_t)stream;
    const long blocks = (np + 255) / 256;
    morton_key_kernel<<<(unsigned)blocks, 256, 0, s>>>(s0_dev, np, au, av, origin_xyz[au], origin_xyz[av],
                                                       65536.0 / wu, 65536.0 / wv, keys_in, idx_in);
    rc = launch_check("morton_key_kernel");
    if (rc) return rc;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, idx_in, perm_dev, (int)np, 0,
                                                    32, s);
    if (e != cudaSuccess) return cuda_fail(e, "cub::DeviceRadixSort::SortPairs");
    return TT_OK;
}
