// Host-side internals shared by the translation units of libndstats_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "nds_common.cuh"

namespace nds { struct PendingCall; }

// ---- peer-memory exchange between the ranks of a sharded call (nds_xchg.cu) ------------------------------------
constexpr int XCHG_MAX_RANKS = 8;
constexpr int XCHG_SLOTS = 4;                          // exchange steps whose payloads may be in flight
constexpr size_t XCHG_SRC_BYTES = 1040 * 1024;         // largest payload of one rank in one step (the keys of the small segments: 1 MB + header)
namespace nds {
struct XchgDev {                                       // passed by value to the exchange kernels
    unsigned char *inbox[XCHG_MAX_RANKS];              // rank r's inbox as mapped into this process ([me]: local memory)
    unsigned long long *flags[XCHG_MAX_RANKS];         // rank r's flags: [slot][source rank] = sequence number + 1
    unsigned long long *done;                          // local: CTAs of a push kernel that have finished, per slot
    size_t src_stride, slot_stride;
    unsigned long long timeout_ns;
    int nranks, me;
};
}

struct nds_ctx {
    int device = 0;
    int sm_count = 148;
    int max_smem_optin = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    // grow-only device arena (scratch for staged inputs, key buffers, histograms, rank state)
    void *arena = nullptr;
    size_t arena_bytes = 0;
    int *d_status = nullptr;       // deferred status bits
    int *h_status = nullptr;       // pinned mirror
    uint64_t launches = 0;
    std::string last_error;
    const char *last_path = "none";
    std::string forced_path = "auto";
    // enqueued bracket-select calls whose rare fallback (a rank escaped its bracket) is decided at the next
    // nds_ctx_synchronize (see STATUS_BIT_FALLBACK)
    std::vector<nds::PendingCall> *pending = nullptr;
    // NCCL (dlopen'd lazily, see nds_nccl.cpp)
    void *nccl_comm = nullptr;
    int nranks = 1;
    int rank = 0;
    // peer-memory exchange (nds_xchg.cu): set up by nds_comm_init_rank when every rank can map every peer's inbox
    void *xchg = nullptr;
    unsigned long long xseq = 0;     // sequence number of the next exchange step (identical on every rank)
    uint64_t xchg_steps = 0;         // exchange steps enqueued so far (peer-memory or NCCL)
    uint64_t host_syncs = 0;         // host round trips inside sharded calls so far
    void *stage = nullptr;           // pinned bounce buffers + streams for pageable host inputs (nds_stage.cu)
};

namespace nds {

// Everything one quantile call needs once pointers are on the device.
struct DeviceCall {
    nds_dtype dtype;
    const void *data;        // device pointer to the logical first element
    const uint8_t *valid;    // optional validity mask with the same geometry (nullptr = none)
    LaneGeom geom;
    QuantArgs q;
    void *out;               // device, C-order result
    uint8_t *out_valid;      // optional, C-order result shape
    double brk_wsum = 0.0;   // bracket select: sum over quantiles of 2 c sqrt(p (1 - p)), see nds_bracket.cu
    bool lanes_padded = false;  // `data` is the layout stage's scratch: every lane is followed by readable padding up
                                // to a multiple of 16 bytes (the lane length itself need not be one)
};

// An enqueued long-lane call that may still need the radix fallback (decided at nds_ctx_synchronize)
struct PendingCall {
    DeviceCall call;                         // call.q.qs / call.q.ranks are re-uploaded from `words`
    std::vector<unsigned long long> words;   // the nq quantiles (as raw doubles) or ranks
    bool is_ranks;
};

// Arena: returns a device pointer to at least `bytes` (256-byte aligned); contents are scratch.
cudaError_t arena_reserve(nds_ctx *ctx, size_t bytes);

// Host -> device copy of a caller's array in stream order of ctx->stream: pageable sources of a few MB and more go
// through several host threads with pinned bounce buffers (nds_stage.cu), everything else is one cudaMemcpyAsync.
cudaError_t stage_h2d(nds_ctx *ctx, void *dst, const void *src, size_t bytes);
void stage_destroy(nds_ctx *ctx);

// --- selection paths (each returns cudaSuccess or the first launch error) -------------------------
// CTA-per-lane MSB radix select; any dtype / stride / lane length / mask.  The correctness workhorse.
cudaError_t launch_generic_cta(nds_ctx *ctx, const DeviceCall &c);
// One (or few) very long lanes: grid-wide multi-pass radix select, histograms in the arena.
// `scratch` must hold long_scratch_bytes(nq).  When ctx->nccl_comm is set and `sharded` is true the
// per-pass histograms and element counts are summed over ranks with ncclAllReduce.
size_t long_scratch_bytes(int nq);
const void *long_total_count_ptr(const void *scratch, int nq);  // device address of the global element count
cudaError_t launch_long_radix(nds_ctx *ctx, const DeviceCall &c, void *scratch, bool sharded);
// One long contiguous lane in ONE streaming pass: sampled value brackets, private-counter classification,
// candidate compaction (nds_bracket.cu).  `*failed` != 0 (synchronous calls) means a rank escaped its bracket
// or scratch ran out: the caller re-runs the call with launch_long_radix.  Asynchronous calls raise
// STATUS_BIT_FALLBACK in the ctx status word instead.
bool bracket_supported(const nds_ctx *ctx, const DeviceCall &c, bool sharded, bool forced);
double brk_csigma();
size_t bracket_scratch_bytes(const DeviceCall &c, bool sharded);
cudaError_t launch_bracket(nds_ctx *ctx, const DeviceCall &c, void *scratch, bool sharded, bool synchronous,
                           int *failed, unsigned long long *n_total_out);
// Contiguous short lanes (f32/f64/ints up to a few thousand elements): warp-per-lane bit-sliced select
// fed by TMA bulk copies.  Returns cudaErrorNotSupported when the call does not qualify.
bool short_bitslice_supported(const nds_ctx *ctx, const DeviceCall &c);
cudaError_t launch_short_bitslice(nds_ctx *ctx, const DeviceCall &c);

// Strided lanes whose neighbours are contiguous (column lanes): CTA per strip of 8 lanes, bit-sliced.
bool cols_bitslice_supported(const nds_ctx *ctx, const DeviceCall &c);
cudaError_t launch_cols_bitslice(nds_ctx *ctx, const DeviceCall &c);

// Contiguous lanes of up to 786432 elements that the warp-per-lane kernel does not take (longer than 8192, or too few
// lanes): CTA -- or cluster of CTAs -- per lane, bit-sliced (nds_mid.cu).
bool mid_bitslice_supported(const nds_ctx *ctx, const DeviceCall &c);
int mid_bitslice_cluster(const nds_ctx *ctx, const DeviceCall &c);   // CTAs that share one lane (1, 2, 4 or 8)
cudaError_t launch_mid_bitslice(nds_ctx *ctx, const DeviceCall &c);

// Layout stage (nds_layout.cu): the lanes of `c` gathered into dst[lane * pitch + r] (elements of 4 or 8 bytes).
cudaError_t launch_gather_lanes(nds_ctx *ctx, const DeviceCall &c, size_t esize, void *dst, unsigned long long pitch);

// Peer-memory exchange (nds_xchg.cu).  `dev_words` (optional, device): the number of 8-byte words to exchange, `gate`
// (optional, device): skip the step when the word is 0 -- both must hold the same value on every rank.
bool xchg_setup(nds_ctx *ctx);                      // collective; false = stay on NCCL
void xchg_teardown(nds_ctx *ctx, bool collective);
bool xchg_ready(const nds_ctx *ctx);
cudaError_t xchg_allreduce_u64(nds_ctx *ctx, unsigned long long *buf, size_t words, const unsigned long long *dev_words,
                               const unsigned long long *gate);
cudaError_t xchg_allgather(nds_ctx *ctx, const void *send, void *recv, size_t bytes_per_rank,
                           const unsigned long long *dev_words, const unsigned long long *gate);

// NCCL shim (nds_nccl.cpp)
int nccl_allreduce_sum_u64(nds_ctx *ctx, void *buf, size_t count, cudaStream_t stream, std::string *err);
int nccl_allgather_bytes(nds_ctx *ctx, const void *send, void *recv, size_t bytes_per_rank, cudaStream_t stream,
                         std::string *err);

}  // namespace nds
