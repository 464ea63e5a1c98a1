// Host -> device staging of a caller's PAGEABLE array (what a numpy caller passes; the reference takes any
// `&mut ArrayRef`, src/quantile/mod.rs:417-428).
//
// cudaMemcpyAsync from pageable memory is staged by the driver through one internal pinned buffer by one thread
// (~11 GB/s measured on the B200 host, a fifth of the PCIe link).  Here a few host threads each run their own pipeline
// over an interleaved set of chunks: memcpy of a chunk into one of the thread's two pinned bounce buffers, asynchronous
// copy of that buffer to the device on the thread's own stream, while the thread already fills its other buffer.  The
// ctx stream then waits for the threads' last copies.  Pinned / registered sources keep the single asynchronous copy.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "nds_internal.h"

namespace nds {

namespace {
constexpr size_t STAGE_CHUNK = 4u << 20;        // bytes per bounce buffer
constexpr size_t STAGE_MIN_BYTES = 8u << 20;    // below this the plain copy wins (thread start-up)
constexpr int STAGE_MAX_THREADS = 8;
constexpr int STAGE_BUFS = 2;                   // bounce buffers per thread
}  // namespace

struct StagePool {
    int nthreads = 0;
    unsigned char *bounce[STAGE_MAX_THREADS][STAGE_BUFS] = {};
    cudaStream_t stream[STAGE_MAX_THREADS] = {};
    cudaEvent_t buf_free[STAGE_MAX_THREADS][STAGE_BUFS] = {};   // the copy out of a bounce buffer has finished
    cudaEvent_t done[STAGE_MAX_THREADS] = {};                   // a thread's last copy of the current call
    cudaEvent_t start = nullptr;                                // the ctx stream has reached the call
};

static int stage_thread_count() {
    if (const char *e = std::getenv("NDS_STAGE_THREADS")) return std::max(0, std::min(STAGE_MAX_THREADS, std::atoi(e)));
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::max(2u, std::min<unsigned>(STAGE_MAX_THREADS, hw / 2));
}

static StagePool *stage_pool(nds_ctx *ctx) {
    if (ctx->stage) return static_cast<StagePool *>(ctx->stage);
    StagePool *p = new StagePool();
    p->nthreads = stage_thread_count();
    bool ok = p->nthreads > 0 && cudaEventCreateWithFlags(&p->start, cudaEventDisableTiming) == cudaSuccess;
    for (int t = 0; ok && t < p->nthreads; ++t) {
        ok = cudaStreamCreateWithFlags(&p->stream[t], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&p->done[t], cudaEventDisableTiming) == cudaSuccess;
        for (int b = 0; ok && b < STAGE_BUFS; ++b)
            ok = cudaHostAlloc((void **)&p->bounce[t][b], STAGE_CHUNK, cudaHostAllocDefault) == cudaSuccess &&
                 cudaEventCreateWithFlags(&p->buf_free[t][b], cudaEventDisableTiming) == cudaSuccess;
    }
    ctx->stage = p;
    if (!ok) {
        cudaGetLastError();
        p->nthreads = 0;   // keep the pool as a marker: the plain copy is used from now on
    }
    return p;
}

void stage_destroy(nds_ctx *ctx) {
    StagePool *p = static_cast<StagePool *>(ctx->stage);
    if (!p) return;
    for (int t = 0; t < STAGE_MAX_THREADS; ++t) {
        if (p->stream[t]) cudaStreamSynchronize(p->stream[t]);
        for (int b = 0; b < STAGE_BUFS; ++b) {
            if (p->bounce[t][b]) cudaFreeHost(p->bounce[t][b]);
            if (p->buf_free[t][b]) cudaEventDestroy(p->buf_free[t][b]);
        }
        if (p->done[t]) cudaEventDestroy(p->done[t]);
        if (p->stream[t]) cudaStreamDestroy(p->stream[t]);
    }
    if (p->start) cudaEventDestroy(p->start);
    delete p;
    ctx->stage = nullptr;
}

// `dst` (device) <- `src` (host), in stream order of ctx->stream on return (the copies themselves may still be running)
cudaError_t stage_h2d(nds_ctx *ctx, void *dst, const void *src, size_t bytes) {
    bool pageable = false;
    if (bytes >= STAGE_MIN_BYTES) {
        cudaPointerAttributes attr;
        pageable = cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeUnregistered;
        cudaGetLastError();
    }
    StagePool *p = pageable ? stage_pool(ctx) : nullptr;
    if (!p || p->nthreads == 0) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);

    // `dst` is arena scratch that earlier work on the ctx stream may still be reading: the threads' streams start behind it
    cudaError_t e = cudaEventRecord(p->start, ctx->stream);
    if (e != cudaSuccess) return e;
    const size_t nchunks = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
    const int nt = (int)std::min<size_t>((size_t)p->nthreads, nchunks);
    std::vector<cudaError_t> err((size_t)nt, cudaSuccess);
    auto work = [&](int t) {
        cudaError_t te = cudaSetDevice(ctx->device);
        if (te == cudaSuccess) te = cudaStreamWaitEvent(p->stream[t], p->start, 0);
        size_t turn = 0;
        for (size_t c = (size_t)t; c < nchunks && te == cudaSuccess; c += (size_t)nt, ++turn) {
            const int b = (int)(turn % STAGE_BUFS);
            const size_t off = c * STAGE_CHUNK, len = std::min(STAGE_CHUNK, bytes - off);
            te = cudaEventSynchronize(p->buf_free[t][b]);   // (also the copy an earlier call left behind; a fresh event is ready)
            if (te != cudaSuccess) break;
            std::memcpy(p->bounce[t][b], (const char *)src + off, len);
            te = cudaMemcpyAsync((char *)dst + off, p->bounce[t][b], len, cudaMemcpyHostToDevice, p->stream[t]);
            if (te == cudaSuccess) te = cudaEventRecord(p->buf_free[t][b], p->stream[t]);
        }
        if (te == cudaSuccess) te = cudaEventRecord(p->done[t], p->stream[t]);
        err[(size_t)t] = te;
    };
    std::vector<std::thread> threads;
    threads.reserve((size_t)nt);
    for (int t = 1; t < nt; ++t) {
        try {
            threads.emplace_back(work, t);
        } catch (...) {   // no more threads to be had: this one takes the chunks itself (nothing may unwind across the C ABI)
            work(t);
        }
    }
    work(0);
    for (auto &th : threads) th.join();
    for (int t = 0; t < nt; ++t) {
        if (err[(size_t)t] != cudaSuccess) e = err[(size_t)t];
        else if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, p->done[t], 0);
    }
    if (e != cudaSuccess) {   // leave nothing in flight that still reads the bounce buffers of a failed call
        for (int t = 0; t < nt; ++t) cudaStreamSynchronize(p->stream[t]);
    }
    return e;
}

}  // namespace nds
