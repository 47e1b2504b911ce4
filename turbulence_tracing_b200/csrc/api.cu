// Error state, version and the host-buffer convenience entry point of libtt_b200.so.
#include "common.cuh"

#include <string.h>

#include <atomic>
#include <mutex>

namespace tt {

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Stream-ordered 4-byte scratch word ("did the event kernel defer any ray?").  It comes from a memory pool OWNED BY
// THIS LIBRARY (one per device, created on first use, release threshold = keep): cudaMallocAsync from the device's
// default pool right after a synchronisation costs 0.7-15 ms on a B200 VM (the pool hands its memory back to the
// driver at every sync, profiles/r01_diag_mallocasync.txt), and changing the default pool's threshold would change
// the behaviour of every other cudaMallocAsync user in the host process.
void* scratch_alloc(size_t bytes, cudaStream_t s) {
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {};
    static bool tried[64] = {};
    int dev = 0;
    cudaMemPool_t pool = nullptr;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) {
        std::lock_guard<std::mutex> lock(mu);
        if (!tried[dev]) {
            tried[dev] = true;
            cudaMemPoolProps props;
            memset(&props, 0, sizeof(props));
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            cudaMemPool_t made = nullptr;
            if (cudaMemPoolCreate(&made, &props) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(made, cudaMemPoolAttrReleaseThreshold, &keep);
                pools[dev] = made;
            }
            (void)cudaGetLastError();
        }
        pool = pools[dev];
    }
    void* ptr = nullptr;
    const cudaError_t e = pool ? cudaMallocFromPoolAsync(&ptr, bytes, pool, s) : cudaMallocAsync(&ptr, bytes, s);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return ptr;
}

// two zeroed words: [0] "did the first pass defer a ray?", [1] number of deferred rays found by the compaction
unsigned int* scratch_flag(cudaStream_t s) {
    unsigned int* flag = (unsigned int*)scratch_alloc(2 * sizeof(unsigned int), s);
    if (flag) cudaMemsetAsync(flag, 0, 2 * sizeof(unsigned int), s);
    return flag;
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return TT_ERR_CUDA;
}

}  // namespace tt

extern "C" int tt_abi_version(void) { return TT_B200_ABI_VERSION; }

extern "C" const char* tt_last_error(void) { return tt::g_err; }

extern "C" unsigned long long tt_launch_count(void) { return tt::g_launches.load(std::memory_order_relaxed); }

extern "C" int tt_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        tt::cuda_fail(e, "cudaGetDeviceCount");
        return -1;
    }
    return n;
}

namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
};
}  // namespace

// Whole path for one bundle with host arrays (ElectronCube.external_ne + calc_dndr + solve):
// the call a ctypes binding inside the reference would make.
extern "C" int tt_solve_host(const double* ne_host, const int n_xyz[3], const double origin_xyz[3],
                             const double spacing_xyz[3], int par, double nc, double ne_max, double extent,
                             int steps_per_cell, int dtype, const double* s0_host, long np, double* rf_host,
                             double* sf_host, unsigned long long* ray_steps_host) {
    using namespace tt;
    TT_REQUIRE(ne_host && n_xyz && origin_xyz && spacing_xyz && s0_host && rf_host, "tt_solve_host: null pointer");
    TT_REQUIRE(np >= 0 && np < (1L << 31), "tt_solve_host: ray count must be in [0, 2^31)");
    TT_REQUIRE(dtype == TT_F32 || dtype == TT_F64, "tt_solve_host: dtype must be TT_F32 or TT_F64");
    for (int i = 0; i < 3; ++i) TT_REQUIRE(n_xyz[i] >= 2, "tt_solve_host: every axis needs >= 2 points");
    if (tt_device_count() <= 0) {
        set_error("tt_solve_host: no CUDA device (there is no CPU fallback)");
        return TT_ERR_CUDA;
    }
    const size_t nvox = (size_t)n_xyz[0] * n_xyz[1] * n_xyz[2];
    size_t sort_bytes = 0;
    int rc = tt_sort_rays_workspace(np, &sort_bytes);
    if (rc) return rc;
    DevBuf ne, grid, s0, rf, sf, perm, ws, cnt, status;
    TT_CUDA(ne.alloc(nvox * sizeof(double)));
    TT_CUDA(grid.alloc(nvox * (dtype == TT_F32 ? 16 : 32)));
    TT_CUDA(s0.alloc((size_t)np * 6 * sizeof(double)));
    TT_CUDA(rf.alloc((size_t)np * 4 * sizeof(double)));
    if (sf_host) TT_CUDA(sf.alloc((size_t)np * 6 * sizeof(double)));
    TT_CUDA(perm.alloc((size_t)np * sizeof(uint32_t)));
    TT_CUDA(ws.alloc(sort_bytes));
    TT_CUDA(cnt.alloc(sizeof(unsigned long long)));
    TT_CUDA(status.alloc((size_t)np));
    cudaStream_t s = 0;
    // (pageable host buffers are staged by worker threads: tt_h2d_pageable, csrc/h2d.cu)
    { int rc_ = tt_h2d_pageable(ne.p, ne_host, nvox * sizeof(double), (tt_stream_t)s); if (rc_) return rc_; }
    { int rc_ = tt_h2d_pageable(s0.p, s0_host, (size_t)np * 6 * sizeof(double), (tt_stream_t)s); if (rc_) return rc_; }
    TT_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), s));
    rc = tt_calc_dndr(ne.p, TT_F64, n_xyz, spacing_xyz, par, nc, ne_max, grid.p, dtype, s);
    if (rc) return rc;
    rc = tt_sort_rays((const double*)s0.p, np, par, origin_xyz, spacing_xyz, n_xyz, (uint32_t*)perm.p, ws.p,
                      sort_bytes, s);
    if (rc) return rc;
    tt_trace_params p;
    memset(&p, 0, sizeof(p));
    for (int i = 0; i < 3; ++i) {
        p.n_xyz[i] = n_xyz[i];
        p.origin_xyz[i] = origin_xyz[i];
        p.spacing_xyz[i] = spacing_xyz[i];
    }
    p.par = par;
    p.extent = extent;
    p.s_max = 2.8284271247461903 * extent;   // sqrt(8) * extent (particle_tracker.py:317)
    p.steps_per_cell = steps_per_cell;
    p.dtype = dtype;
    // FP32 at one step per cell: the production path over the face-coefficient grid when it fits (as ElectronCube.solve)
    DevBuf faces;
    const size_t face_bytes = (dtype == TT_F32 && steps_per_cell == 1) ? tt_face_grid_bytes(n_xyz, par) : 0;
    if (face_bytes) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && face_bytes + (1ull << 30) < free_b && faces.alloc(face_bytes) != cudaSuccess)
            (void)cudaGetLastError();
    }
    if (faces.p) {
        rc = tt_build_face_grid(grid.p, n_xyz, spacing_xyz, par, faces.p, s);
        if (rc) return rc;
        rc = tt_trace_faces(&p, grid.p, faces.p, (const double*)s0.p, np, (const uint32_t*)perm.p, (double*)rf.p, (double*)sf.p,
                            (unsigned long long*)cnt.p, (uint8_t*)status.p, s);
    } else {
        rc = tt_trace(&p, grid.p, (const double*)s0.p, np, (const uint32_t*)perm.p, (double*)rf.p, (double*)sf.p,
                      (unsigned long long*)cnt.p, (uint8_t*)status.p, s);
    }
    if (rc) return rc;
    TT_CUDA(cudaMemcpyAsync(rf_host, rf.p, (size_t)np * 4 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (sf_host) TT_CUDA(cudaMemcpyAsync(sf_host, sf.p, (size_t)np * 6 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (ray_steps_host)
        TT_CUDA(cudaMemcpyAsync(ray_steps_host, cnt.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    TT_CUDA(cudaStreamSynchronize(s));
    return TT_OK;
}
