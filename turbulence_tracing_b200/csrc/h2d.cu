// tt_h2d_pageable: host -> device copy of PAGEABLE memory (the plain numpy arrays a drop-in caller of ElectronCube.solve
// passes: s0 is 4.8 GB for 1e8 rays) at more than the ~10 GB/s of cudaMemcpyAsync's own single-threaded staging.
//
// The reference keeps its rays in host numpy arrays (particle_tracker.py:258-310 init_beam -> self.s0) and its MPI
// example builds them per rank (example_MPI.py:117-131); the GPU path has to move them.  With pinned buffers the
// upload hides behind the trace (bench e2e: 333 ms for 1e8 rays); from pageable memory it was the bound (508 ms).
// Here worker threads copy the source piece by piece into a process-wide ring of pinned buffers and queue one DMA per
// piece on the caller's stream: the staging memcpy runs on several cores, the DMA engine streams behind it.
//
// Semantics = cudaMemcpyAsync from pageable memory: on return every source byte has been read (the caller may reuse
// the array), the last DMAs may still be in flight on `stream`.
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include <sched.h>
#include "common.cuh"

namespace tt {

static constexpr size_t kPiece = 4u << 20;         // bytes per staging buffer
static constexpr int kSlots = 32;                  // ring: 128 MB of pinned memory per device, allocated at first use

struct StagingRing {
    char* buf[kSlots] = {};
    cudaEvent_t done[kSlots] = {};                 // DMA out of the slot finished
    bool used[kSlots] = {};
    bool ok = false;
};
static StagingRing g_ring[64];
static std::mutex g_ring_mutex;                    // one upload at a time per process: the ring is shared

static int ring_for(int dev, StagingRing** out) {
    StagingRing& r = g_ring[dev];
    if (!r.ok) {
        for (int i = 0; i < kSlots; ++i) {
            cudaError_t e = cudaHostAlloc((void**)&r.buf[i], kPiece, cudaHostAllocDefault);
            if (e != cudaSuccess) return cuda_fail(e, "tt_h2d_pageable: cudaHostAlloc of the staging ring");
            e = cudaEventCreateWithFlags(&r.done[i], cudaEventDisableTiming);
            if (e != cudaSuccess) return cuda_fail(e, "tt_h2d_pageable: cudaEventCreate");
        }
        r.ok = true;
    }
    *out = &r;
    return TT_OK;
}

static int worker_count() {
    if (const char* s = std::getenv("TT_H2D_THREADS")) {
        const int n = std::atoi(s);
        if (n >= 1) return n > 16 ? 16 : n;
    }
    cpu_set_t set;
    int n = 4;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);      // the CPUs this rank is bound to
    n = n / 2;                                                                   // leave room for the caller's other threads
    return n < 1 ? 1 : (n > 8 ? 8 : n);
}

}  // namespace tt

extern "C" int tt_h2d_pageable(void* dst_dev, const void* src_host, size_t bytes, tt_stream_t stream) {
    using namespace tt;
    TT_REQUIRE(bytes == 0 || (dst_dev && src_host), "tt_h2d_pageable: null pointer");
    if (bytes == 0) return TT_OK;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "tt_h2d_pageable: cudaGetDevice");
    TT_REQUIRE(dev >= 0 && dev < 64, "tt_h2d_pageable: device index %d out of range", dev);
    cudaStream_t s = (cudaStream_t)stream;
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, src_host) == cudaSuccess &&
                        (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
    (void)cudaGetLastError();                      // (an unregistered pointer is not an error here)
    if (bytes < 2 * kPiece || pinned) {            // small, or already page-locked: one plain (asynchronous) copy
        e = cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, s);
        return e == cudaSuccess ? TT_OK : cuda_fail(e, "tt_h2d_pageable: cudaMemcpyAsync");
    }
    std::lock_guard<std::mutex> lock(g_ring_mutex);
    StagingRing* ring = nullptr;
    int rc = ring_for(dev, &ring);
    if (rc) return rc;
    const size_t pieces = (bytes + kPiece - 1) / kPiece;
    int nthreads = worker_count();
    if ((size_t)nthreads > pieces) nthreads = (int)pieces;
    std::atomic<int> failed{0};
    // Piece i is staged in slot i % kSlots, by worker i % T (every worker walks its pieces in ascending order).  A slot is
    // refilled only after the DMA of the piece that used it before has finished: seq[slot] counts the pieces queued on the
    // slot in this call (published after the event record), so the worker of piece i first waits until piece i - kSlots
    // has been queued (a few microseconds at most: that piece is kSlots pieces older), then for its event.
    std::atomic<unsigned> seq[kSlots];
    for (int i = 0; i < kSlots; ++i) seq[i].store(0);
    std::mutex issue;                              // serialises cudaMemcpyAsync + event record on the stream
    auto work = [&](int t) {
        if (cudaSetDevice(dev) != cudaSuccess) { failed = 1; return; }
        for (size_t i = (size_t)t; i < pieces && !failed.load(); i += (size_t)nthreads) {
            const int slot = (int)(i % kSlots);
            const unsigned round = (unsigned)(i / kSlots);
            while (seq[slot].load(std::memory_order_acquire) != round) {
                if (failed.load()) return;
                std::this_thread::yield();
            }
            if ((round > 0 || ring->used[slot]) && cudaEventSynchronize(ring->done[slot]) != cudaSuccess) { failed = 1; return; }
            const size_t off = i * kPiece, n = bytes - off < kPiece ? bytes - off : kPiece;
            std::memcpy(ring->buf[slot], (const char*)src_host + off, n);
            {
                std::lock_guard<std::mutex> g(issue);
                if (cudaMemcpyAsync((char*)dst_dev + off, ring->buf[slot], n, cudaMemcpyHostToDevice, s) != cudaSuccess ||
                    cudaEventRecord(ring->done[slot], s) != cudaSuccess) { failed = 1; return; }
            }
            ring->used[slot] = true;
            seq[slot].store(round + 1, std::memory_order_release);
        }
    };
    std::vector<std::thread> pool;
    try {
        for (int t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
    } catch (...) {                                // no thread to be had: the pieces of the missing workers would never come
        failed = 2;
    }
    work(0);
    for (auto& th : pool) th.join();
    if (failed.load() == 2) {
        set_error("tt_h2d_pageable: could not start the staging threads (TT_H2D_THREADS=1 stages on the calling thread)");
        return TT_ERR_CUDA;
    }
    if (failed.load()) return cuda_fail(cudaGetLastError(), "tt_h2d_pageable: staging copy failed");
    return TT_OK;
}
