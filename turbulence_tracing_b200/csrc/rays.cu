// Launch-ray generation on device and Morton ordering of rays.
//   tt_init_beam : ElectronCube.init_beam (particle_tracker.py:258-310) with a counter-based RNG
//   tt_sort_rays : Z-order permutation of the rays by transverse launch position, so that the 32
//                  rays of a warp walk through the same / adjacent grid cells (gather locality).
#include "rays_one.cuh"             // init_beam_ray, morton_key (host + device)

#include <cub/device/device_radix_sort.cuh>

namespace tt {

__global__ void __launch_bounds__(256) init_beam_kernel(long np, long first, uint64_t seed, double beam_size,
                                                        double divergence, double extent, int par,
                                                        double* __restrict__ s0) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    init_beam_ray(i, np, first, seed, beam_size, divergence, extent, par, s0);       // rays_one.cuh
}

__global__ void __launch_bounds__(256) morton_key_kernel(const double* __restrict__ s0, long np, int au, int av,
                                                         double ou, double ov, double su, double sv,
                                                         uint32_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    keys[i] = morton_key(s0, i, np, au, av, ou, ov, su, sv);
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
    // only the leading bits that resolve a quarter of a cell: a 513^3 cube sorts 22 bits (3 radix passes instead of 4);
    // rays within one such box keep their storage order (the sort is stable), a ray's result does not depend on its place
    int cells = (n_xyz[au] > n_xyz[av] ? n_xyz[au] : n_xyz[av]) - 1, bits = 2;
    while ((1 << (bits - 2)) < cells && bits < 16) ++bits;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, idx_in, perm_dev, (int)np,
                                                    32 - 2 * bits, 32, s);
    if (e != cudaSuccess) return cuda_fail(e, "cub::DeviceRadixSort::SortPairs");
    return TT_OK;
}
