// Peer-memory exchange for the sharded 1-D quantile path: the small all-reduce / all-gather steps between the
// selection kernels, done by OUR kernels over NVLink instead of by a library collective.
//
// Every rank owns one "inbox" buffer in its HBM that all peers map (CUDA IPC between the one-process-per-GPU ranks, or
// plain peer access inside one process).  An exchange step `seq` is two small kernels on the ctx stream:
//
//   push    the rank writes its payload into slot seq % XCHG_SLOTS of EVERY peer's inbox (16-byte stores straight over
//           NVLink / NVSwitch), fences at system scope, and the last CTA to finish releases one flag per peer
//           (st.release.sys, value seq + 1: flags are never reset, so there is no second hand-shake);
//   reduce  (or gather) spins on the nranks flags of its OWN inbox (ld.acquire.sys, local memory), then sums (copies) the
//           payloads that the peers have left there.
//
// What this buys over ncclAllReduce on these 16 B ... 500 KB messages: ~5 us instead of ~20-30 us per step, and -- because
// the kernels are ours -- a step can be SKIPPED or SIZED on the device (`gate`, `dev_words`: words every rank derives from
// already-reduced data, hence identical everywhere), so the refinement loop of the bracket select is enqueued without
// a host round trip: idle rounds cost two empty launches, not a collective.  Slot reuse is safe with >= 2 slots because
// a rank consumes step k before it pushes step k + 1 (stream order), and nobody can push step k + XCHG_SLOTS before
// having consumed step k + XCHG_SLOTS - 1, which needs everybody's push of that step.
//
// One process per GPU and one GPU per rank: the waiting kernels of different ranks always run on different devices
// (B200_PROFILING.md forbids mutually waiting launches on ONE GPU).  A rank that waits longer than the time-out gives up,
// raises STATUS_BIT_XCHG and the call returns NDS_NCCL_ERROR instead of hanging the box.
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "nds_internal.h"

namespace nds {

namespace {

struct XchgRecord {  // what the ranks tell each other at set-up (all-gathered once, through NCCL)
    cudaIpcMemHandle_t handle;
    unsigned long long pid;
    unsigned long long ptr;
    int device;
    int ok;
    char pad[128 - sizeof(cudaIpcMemHandle_t) - 24];
};
static_assert(sizeof(XchgRecord) == 128, "record size");

struct XchgHost {
    XchgDev dev;
    void *local = nullptr;             // this rank's inbox allocation
    void *mapped[XCHG_MAX_RANKS] = {};  // peers' inboxes as mapped here (IPC), to be closed at teardown
    bool ipc[XCHG_MAX_RANKS] = {};
    unsigned char *stage = nullptr;    // small separate allocation for the set-up / tear-down collectives
    int nranks = 0;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
constexpr int XCHG_SEQ_BITS = 40;  // flag word = (sequence number + 1) | payload length in 16-byte units << 40
constexpr unsigned long long XCHG_SEQ_MASK = (1ull << XCHG_SEQ_BITS) - 1ull;
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ size_t step_bytes(size_t bytes, const unsigned long long *dev_words) {
    if (dev_words) {
        const size_t b = ((size_t)*dev_words * 8 + 15) & ~(size_t)15;
        if (b < bytes) bytes = b;
    }
    return bytes;
}

// Phase 1 of a step: this CTA's share of the payload -> slot (seq % XCHG_SLOTS, me) of EVERY rank's inbox; the last CTA
// of the grid to finish releases one flag per rank.  Plain stores, one barrier, then ONE system-scope fence by thread 0:
// the barrier orders the CTA's stores before the fence (PTX causality is cumulative over bar.sync), the fence orders
// them before the counter update, and the last CTA fences again before it releases the flags.
__device__ __forceinline__ void xchg_push(const XchgDev &x, unsigned long long seq, const uint4 *__restrict__ src, size_t n16) {
    const unsigned slot = (unsigned)(seq % XCHG_SLOTS);
    const size_t off = (size_t)slot * x.slot_stride + (size_t)x.me * x.src_stride;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = src[i];
#pragma unroll 1
        for (int j = 0; j < x.nranks; ++j) {
            int r = x.me + j;  // the local copy first, then the peers in a rank-dependent order
            if (r >= x.nranks) r -= x.nranks;
            reinterpret_cast<uint4 *>(x.inbox[r] + off)[i] = v;
        }
    }
    // One system-scope fence per CTA (thread 0, after the barrier that orders the CTA's stores before it), a counter
    // that tells the last CTA it is the last, and then ONE more fence in each of the first nranks threads, all at the
    // same time, before each of them releases the flag of "its" rank with a relaxed store.  (A st.release per flag from
    // one thread costs one NVLink round trip per rank, one after the other: 8 ranks, ~25 us.)  The flag word carries the
    // payload's length in its upper bits.
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        int last = 1;
        if (gridDim.x > 1) {
            __threadfence_system();
            const unsigned long long old = atomicAdd(&x.done[slot], 1ull);
            last = old == (unsigned long long)gridDim.x - 1ull;
            if (last) x.done[slot] = 0;
        }
        s_last = last;
    }
    __syncthreads();
    if (s_last && (int)threadIdx.x < x.nranks) {
        __threadfence_system();
        st_relaxed_sys(x.flags[threadIdx.x] + (size_t)slot * XCHG_MAX_RANKS + x.me, (seq + 1ull) | ((unsigned long long)n16 << XCHG_SEQ_BITS));
    }
}

// Phase 2: wait until every rank's payload of step `seq` sits in this rank's inbox (thread r waits for rank r).
__device__ __forceinline__ void xchg_wait(const XchgDev &x, unsigned long long seq, int *status) {
    if ((int)threadIdx.x < x.nranks) {
        const unsigned long long *f = x.flags[x.me] + (size_t)(seq % XCHG_SLOTS) * XCHG_MAX_RANKS + threadIdx.x;
        if ((ld_acquire_sys(f) & XCHG_SEQ_MASK) != seq + 1ull && !(*reinterpret_cast<volatile int *>(status) & STATUS_BIT_XCHG)) {
            const unsigned long long t0 = global_ns();
            while ((ld_acquire_sys(f) & XCHG_SEQ_MASK) != seq + 1ull) {
                if (global_ns() - t0 > x.timeout_ns) {
                    atomicOr(status, STATUS_BIT_XCHG);
                    break;
                }
            }
        }
    }
    __syncthreads();
}

// One exchange step as ONE kernel: push, wait, then sum what the peers have left in this rank's inbox.  Every CTA reduces
// exactly the 16-byte chunks it pushed itself, so the sum may overwrite the payload in place.  All CTAs wait inside the
// kernel, hence the grid is kept far below the SM count (they are all resident).
__global__ void __launch_bounds__(256)
xchg_allreduce_kernel(XchgDev x, unsigned long long seq, unsigned long long *__restrict__ buf, size_t bytes,
                      const unsigned long long *dev_words, const unsigned long long *gate, int *status) {
    if (gate && *gate == 0) return;
    bytes = step_bytes(bytes, dev_words);
    const size_t n16 = bytes / 16;
    xchg_push(x, seq, reinterpret_cast<const uint4 *>(buf), n16);
    xchg_wait(x, seq, status);
    const unsigned char *base = x.inbox[x.me] + (size_t)(seq % XCHG_SLOTS) * x.slot_stride;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        unsigned long long s0 = 0, s1 = 0;
        for (int r = 0; r < x.nranks; ++r) {
            const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(base + (size_t)r * x.src_stride) + i);
            s0 += (unsigned long long)v.x | ((unsigned long long)v.y << 32);
            s1 += (unsigned long long)v.z | ((unsigned long long)v.w << 32);
        }
        buf[2 * i] = s0;
        buf[2 * i + 1] = s1;
    }
}

__global__ void __launch_bounds__(256)
xchg_allgather_kernel(XchgDev x, unsigned long long seq, const uint4 *__restrict__ send, uint4 *__restrict__ recv,
                      size_t bytes_per_rank, const unsigned long long *dev_words, const unsigned long long *gate, int *status) {
    if (gate && *gate == 0) return;
    const size_t n16_max = bytes_per_rank / 16;
    const size_t n16 = step_bytes(bytes_per_rank, dev_words) / 16;  // what THIS rank sends; the peers' lengths may differ
    xchg_push(x, seq, send, n16);
    xchg_wait(x, seq, status);
    const unsigned char *base = x.inbox[x.me] + (size_t)(seq % XCHG_SLOTS) * x.slot_stride;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const unsigned long long *fl = x.flags[x.me] + (size_t)(seq % XCHG_SLOTS) * XCHG_MAX_RANKS;
    for (int r = 0; r < x.nranks; ++r) {  // every rank's payload has its own length (it came with the flag)
        size_t ncopy = (size_t)(__ldcg(fl + r) >> XCHG_SEQ_BITS);
        if (ncopy > n16_max) ncopy = n16_max;
        const uint4 *s = reinterpret_cast<const uint4 *>(base + (size_t)r * x.src_stride);
        uint4 *d = recv + (size_t)r * n16_max;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncopy; i += stride) d[i] = __ldcg(s + i);
    }
}

unsigned grid_for(const nds_ctx *ctx, size_t bytes) {
    size_t g = (bytes + 16 * 256 * 2 - 1) / (16 * 256 * 2);  // ~8 KB per CTA
    if (g < 1) g = 1;
    const size_t cap = (size_t)std::min(ctx->sm_count / 2, 32);  // every CTA waits inside the kernel: all must be resident
    if (g > cap) g = cap;
    return (unsigned)g;
}

}  // namespace

bool xchg_ready(const nds_ctx *ctx) { return ctx->xchg != nullptr; }

cudaError_t xchg_allreduce_u64(nds_ctx *ctx, unsigned long long *buf, size_t words, const unsigned long long *dev_words,
                               const unsigned long long *gate) {
    XchgHost *h = (XchgHost *)ctx->xchg;
    if (!h) return cudaErrorNotSupported;
    const size_t bytes = (words * 8 + 15) & ~(size_t)15;
    if (bytes > h->dev.src_stride) return cudaErrorInvalidValue;
    const unsigned long long seq = ctx->xseq++;
    xchg_allreduce_kernel<<<grid_for(ctx, bytes), 256, 0, ctx->stream>>>(h->dev, seq, buf, bytes, dev_words, gate, ctx->d_status);
    ctx->launches += 1;
    ctx->xchg_steps += 1;
    return cudaGetLastError();
}

cudaError_t xchg_allgather(nds_ctx *ctx, const void *send, void *recv, size_t bytes_per_rank, const unsigned long long *dev_words,
                           const unsigned long long *gate) {
    XchgHost *h = (XchgHost *)ctx->xchg;
    if (!h) return cudaErrorNotSupported;
    if (bytes_per_rank % 16 != 0 || bytes_per_rank > h->dev.src_stride) return cudaErrorInvalidValue;
    const unsigned long long seq = ctx->xseq++;
    xchg_allgather_kernel<<<grid_for(ctx, bytes_per_rank), 256, 0, ctx->stream>>>(h->dev, seq, (const uint4 *)send, (uint4 *)recv,
                                                                               bytes_per_rank, dev_words, gate, ctx->d_status);
    ctx->launches += 1;
    ctx->xchg_steps += 1;
    return cudaGetLastError();
}

void xchg_teardown(nds_ctx *ctx, bool collective) {
    XchgHost *h = (XchgHost *)ctx->xchg;
    if (!h) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < h->nranks; ++r)
        if (h->ipc[r] && h->mapped[r]) cudaIpcCloseMemHandle(h->mapped[r]);
    if (collective && ctx->nccl_comm && h->stage) {  // nobody frees an inbox that a peer may still have mapped
        std::string err;
        cudaMemsetAsync(h->stage, 0, 8, ctx->stream);
        if (nccl_allreduce_sum_u64(ctx, h->stage, 1, ctx->stream, &err) == 0) cudaStreamSynchronize(ctx->stream);
    }
    if (h->stage) cudaFree(h->stage);
    if (h->local) cudaFree(h->local);
    cudaGetLastError();
    delete h;
    ctx->xchg = nullptr;
}

// Collective over the ranks of ctx->nccl_comm.  Returns true when every rank has mapped every inbox; on any failure
// all ranks agree to stay on the NCCL collectives.
bool xchg_setup(nds_ctx *ctx) {
    if (ctx->xchg) xchg_teardown(ctx, true);
    const int nranks = ctx->nranks, me = ctx->rank;
    if (nranks < 2 || nranks > XCHG_MAX_RANKS || !ctx->nccl_comm) return false;
    if (const char *e = getenv("NDS_XCHG"))
        if (!strcmp(e, "nccl")) return false;  // (every rank reads the same environment)
    cudaSetDevice(ctx->device);
    XchgHost *h = new XchgHost();
    h->nranks = nranks;
    const size_t src_stride = XCHG_SRC_BYTES;
    const size_t slot_stride = src_stride * XCHG_MAX_RANKS;
    const size_t inbox_bytes = slot_stride * XCHG_SLOTS;
    const size_t flag_bytes = (size_t)2 * XCHG_SLOTS * XCHG_MAX_RANKS * 8;  // flags, then the payload lengths
    const size_t total = inbox_bytes + flag_bytes + 256 /* done counters */ + 256;
    XchgRecord mine;
    std::memset(&mine, 0, sizeof mine);
    mine.pid = (unsigned long long)getpid();
    mine.device = ctx->device;
    mine.ok = 1;
    // staging for this set-up's own (NCCL) collectives: separate from the inbox, so that a rank whose inbox allocation
    // failed still takes part in them and everybody agrees to stay on NCCL
    unsigned char *stage = nullptr;
    if (cudaMalloc((void **)&stage, 4096 + (size_t)(nranks + 1) * sizeof(XchgRecord)) != cudaSuccess) {
        cudaGetLastError();
        delete h;
        return false;  // (out of device memory at communicator set-up: nothing will work; the peers time out in NCCL)
    }
    if (cudaMalloc(&h->local, total) != cudaSuccess) {
        cudaGetLastError();
        h->local = nullptr;
        mine.ok = 0;
    }
    unsigned char *base = (unsigned char *)h->local;
    if (base) {
        if (cudaMemset(base, 0, total) != cudaSuccess) mine.ok = 0;
        if (cudaIpcGetMemHandle(&mine.handle, base) != cudaSuccess) {
            cudaGetLastError();
            mine.ok = 0;
        }
        mine.ptr = (unsigned long long)(uintptr_t)base;
    }
    std::vector<XchgRecord> all(nranks);
    std::string err;
    bool good = cudaMemcpy(stage + 2048, &mine, sizeof mine, cudaMemcpyHostToDevice) == cudaSuccess &&
                nccl_allgather_bytes(ctx, stage + 2048, stage + 2048 + sizeof mine, sizeof mine, ctx->stream, &err) == 0 &&
                cudaStreamSynchronize(ctx->stream) == cudaSuccess &&
                cudaMemcpy(all.data(), stage + 2048 + sizeof mine, (size_t)nranks * sizeof mine, cudaMemcpyDeviceToHost) == cudaSuccess;
    unsigned long long bad = good ? 0 : 1;
    if (good) {
        for (int r = 0; r < nranks && !bad; ++r) {
            if (!all[r].ok) { bad = 1; break; }
            if (r == me) {
                h->mapped[r] = base;
                continue;
            }
            if (all[r].pid == mine.pid) {  // ranks of one process: plain peer access
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, ctx->device, all[r].device) != cudaSuccess || !can) { bad = 1; break; }
                cudaError_t pe = cudaDeviceEnablePeerAccess(all[r].device, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { bad = 1; break; }
                cudaGetLastError();
                h->mapped[r] = (void *)(uintptr_t)all[r].ptr;
            } else {
                void *p = nullptr;
                if (cudaIpcOpenMemHandle(&p, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError();
                    bad = 1;
                    break;
                }
                h->mapped[r] = p;
                h->ipc[r] = true;
            }
        }
    }
    // agree: one failure anywhere keeps every rank on NCCL
    unsigned long long agreed = 1;
    if (cudaMemcpy(stage, &bad, 8, cudaMemcpyHostToDevice) == cudaSuccess &&
        nccl_allreduce_sum_u64(ctx, stage, 1, ctx->stream, &err) == 0 && cudaStreamSynchronize(ctx->stream) == cudaSuccess &&
        cudaMemcpy(&agreed, stage, 8, cudaMemcpyDeviceToHost) == cudaSuccess) {
    } else {
        agreed = 1;
    }
    h->stage = stage;
    ctx->xchg = h;
    if (agreed != 0) {
        xchg_teardown(ctx, true);
        return false;
    }
    XchgDev &d = h->dev;
    std::memset(&d, 0, sizeof d);
    d.nranks = nranks;
    d.me = me;
    d.src_stride = src_stride;
    d.slot_stride = slot_stride;
    double ms = 5000.0;
    if (const char *e = getenv("NDS_XCHG_TIMEOUT_MS")) ms = atof(e) > 0 ? atof(e) : ms;
    d.timeout_ns = (unsigned long long)(ms * 1e6);
    for (int r = 0; r < nranks; ++r) {
        d.inbox[r] = (unsigned char *)h->mapped[r];
        d.flags[r] = (unsigned long long *)((unsigned char *)h->mapped[r] + inbox_bytes);
    }
    d.done = (unsigned long long *)(base + inbox_bytes + flag_bytes);
    ctx->xseq = 0;
    return true;
}

}  // namespace nds
