/*
 * ndstats_b200 — C ABI of the B200-native bulk quantile path.
 *
 * This header is the drop-in boundary for ONE path of rust-ndarray/ndarray-stats v0.7.0: the bulk
 * quantile path.  The reference has no FFI of its own; the boundary it exposes is a set of sealed
 * Rust extension traits.  Every entry point below states which reference interface it replaces
 * (paths relative to the reference checkout).  The Rust shim that binds these symbols and restores
 * the trait surface is shown in INTEGRATION.md (sources in ndarray-stats_b200/rust/).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross this boundary;
 *   - `data` / `out` may be host or device pointers (detected with cudaPointerGetAttributes); host
 *     buffers are staged through the context's device arena;
 *   - `shape` and `strides_elems` describe an ndarray-style view: strides are in ELEMENTS, may be
 *     negative, `data` points at the logical first element (ndarray's `as_ptr()`);
 *   - `out` is always written C-contiguous with shape[axis] replaced by nq
 *     (src/quantile/mod.rs:451-452,469); single-q wrappers drop that axis (mod.rs:507);
 *   - the input is never modified (the reference permutes lanes in place and leaves the order
 *     unspecified, mod.rs:190-195, so "unchanged" is conformant);
 *   - there is no CPU fallback: without a CUDA device every compute entry point returns
 *     NDS_CUDA_ERROR.
 */
#ifndef NDSTATS_B200_H
#define NDSTATS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDS_MAX_DIMS 12
#define NDS_ABI_VERSION 1

typedef struct nds_ctx nds_ctx;

typedef enum nds_status {
    NDS_OK = 0,
    NDS_EMPTY_INPUT = 1,      /* QuantileError::EmptyInput          src/errors.rs:121-122 */
    NDS_INVALID_QUANTILE = 2, /* QuantileError::InvalidQuantile(q)  src/errors.rs:123-124; *bad_q_index = first offender */
    NDS_NAN_IN_INPUT = 3,     /* f32/f64 input held a NaN on a non-skipnan call: N32/N64 cannot represent
                                 it (src/quantile/mod.rs:213-216), so the shim panics like noisy_float */
    NDS_UNSUPPORTED = 4,      /* dtype / rank count / layout outside what the device path covers */
    NDS_CUDA_ERROR = 5,
    NDS_NCCL_ERROR = 6,
    NDS_BAD_ARGUMENT = 7,     /* null pointer, ndim out of range, axis out of bounds (the reference panics) */
    NDS_CAST_OVERFLOW = 8,    /* integer Linear: T::from_f64(..).unwrap() would panic, interpolate.rs:142 */
    NDS_INDEX_OUT_OF_BOUNDS = 9, /* Sort1dExt: index >= n (the reference panics, src/sort.rs:27,38) */
    NDS_UNDEFINED_ORDER = 10    /* MinMaxError::UndefinedOrder (a NaN met min/max/argmin/argmax)  src/errors.rs:24-25 */
} nds_status;

typedef enum nds_dtype {
    NDS_I8 = 0, NDS_I16 = 1, NDS_I32 = 2, NDS_I64 = 3,
    NDS_U8 = 4, NDS_U16 = 5, NDS_U32 = 6, NDS_U64 = 7,
    NDS_F32 = 8, /* N32 on plain calls, raw f32 on skipnan calls */
    NDS_F64 = 9  /* N64 / f64 */
} nds_dtype;

/* interpolate::{Lower, Higher, Nearest, Midpoint, Linear} — src/quantile/interpolate.rs:51-62 */
typedef enum nds_interp {
    NDS_LOWER = 0, NDS_HIGHER = 1, NDS_NEAREST = 2, NDS_MIDPOINT = 3, NDS_LINEAR = 4
} nds_interp;

/* ---- library / context ------------------------------------------------------------------ */

int nds_abi_version(void);
const char *nds_status_string(nds_status s); /* Display strings of src/errors.rs:126-135 for 1 and 2 */
size_t nds_dtype_size(nds_dtype t);

/* One context per host thread and GPU: stream, device scratch arena, deferred status word.
 * (The reference is stateless and single-threaded; a ctx is not thread-safe.) */
nds_status nds_ctx_create(int device, nds_ctx **out);
void nds_ctx_destroy(nds_ctx *ctx);
/* Borrow an externally owned cudaStream_t (e.g. torch's current stream); NULL restores the ctx's own.
 * To run on the legacy default stream pass cudaStreamLegacy ((cudaStream_t)0x1), not NULL. */
nds_status nds_ctx_set_stream(nds_ctx *ctx, void *cuda_stream);
void *nds_ctx_get_stream(nds_ctx *ctx);
/* Wait for everything enqueued on the ctx stream; returns (and clears) the deferred data-dependent
 * status of *_enqueue calls: NDS_NAN_IN_INPUT, NDS_CAST_OVERFLOW or a CUDA error. */
nds_status nds_ctx_synchronize(nds_ctx *ctx);
const char *nds_ctx_last_error(nds_ctx *ctx);
/* Number of kernels this ctx has launched so far (bench.py's gpu_launches is a difference of these). */
uint64_t nds_ctx_kernel_launches(nds_ctx *ctx);
/* Name of the selection path the last call dispatched to: "short_bitslice", "mid_bitslice", "cols_bitslice", "bracket",
 * "bracket_sharded", "bracket->long_radix" (a rank escaped its sampled bracket and the radix passes re-ran the call),
 * "long_radix", "long_radix_sharded", "generic_cta", "minmax". */
const char *nds_ctx_last_path(nds_ctx *ctx);
/* Force a selection path for testing: "short_bitslice", "mid_bitslice", "cols_bitslice", "bracket", "long_radix", "generic_cta"
 * (NULL or "auto" = default dispatch).  Returns NDS_BAD_ARGUMENT for unknown names. */
nds_status nds_ctx_force_path(nds_ctx *ctx, const char *name);

/* Host-only argument check with the reference's precedence (src/quantile/mod.rs:440-455): first q outside
 * [0,1] -> NDS_INVALID_QUANTILE (+ *bad_q_index); shape[axis] == 0 -> NDS_EMPTY_INPUT; bad ndim / axis ->
 * NDS_BAD_ARGUMENT; otherwise NDS_OK.  *result_elems (optional) receives the size of the result, 0 meaning
 * "Ok(empty array)".  Needs no GPU and no context; every compute entry point applies the same check. */
nds_status nds_validate_quantile_args(int ndim, const size_t *shape, int axis, const double *qs, size_t nq,
                                      nds_interp interp, size_t *bad_q_index, size_t *result_elems);

/* ---- the quantile path --------------------------------------------------------------------- */

/*
 * QuantileExt::quantiles_axis_mut            src/quantile/mod.rs:417-493   (skipnan = 0, any nq)
 * QuantileExt::quantile_axis_mut             src/quantile/mod.rs:495-508   (skipnan = 0, nq = 1)
 * QuantileExt::quantile_axis_skipnan_mut     src/quantile/mod.rs:510-544   (skipnan = 1; the reference
 *                                            has only the single-q form, any nq is accepted here)
 * Quantile1dExt::{quantile_mut,quantiles_mut} src/quantile/mod.rs:612-633  (ndim = 1, axis = 0)
 *
 * Error precedence as the reference: first q outside [0,1] (NaN included) in qs order ->
 * NDS_INVALID_QUANTILE with *bad_q_index set (mod.rs:440-444, 522-524); then shape[axis] == 0 ->
 * NDS_EMPTY_INPUT (mod.rs:446-449, 526-528); then a zero-size result -> NDS_OK, nothing written
 * (mod.rs:451-455).  skipnan lanes without any non-NaN value yield the canonical quiet NaN
 * (f32 0x7FC00000, f64 0x7FF8000000000000; src/maybe_nan/mod.rs:126-131); for integer dtypes skipnan
 * is a no-op.  Synchronous: `out` is complete on return.
 */
nds_status nds_quantiles_axis(nds_ctx *ctx, nds_dtype dtype, const void *data, int ndim, const size_t *shape,
                              const ptrdiff_t *strides_elems, int axis, const double *qs, size_t nq,
                              nds_interp interp, int skipnan, void *out, size_t *bad_q_index);

/*
 * Same computation for device-resident `data` and `out`, enqueued on the ctx stream without a host
 * synchronisation.  Argument errors are returned immediately; data-dependent conditions
 * (NDS_NAN_IN_INPUT, NDS_CAST_OVERFLOW) are reported by the next nds_ctx_synchronize().
 *
 * Calls on short / column lanes (`nds_ctx_last_path`: "short_bitslice", "cols_bitslice", "generic_cta", "long_radix")
 * are complete in stream order: a consumer on the same stream may read `out` right behind the call.  A LONG lane that
 * takes the one-pass bracket select ("bracket") is complete in stream order unless a requested rank escaped its
 * sampled bracket (~1e-4 per call for 99 quantiles at the default 5 sigma; always for lanes whose segments stay large
 * after three refinement rounds): then `out` is PROVISIONAL -- it holds unspecified values -- until
 * nds_ctx_synchronize() has re-run the call on the radix passes, which re-reads `data`.  So for long lanes: keep `data`
 * unmodified and do not consume `out` before nds_ctx_synchronize() returns; a caller that needs strict stream order on
 * long lanes forces the radix passes (nds_ctx_force_path(ctx, "long_radix")) or uses the synchronous entry point.
 */
nds_status nds_quantiles_axis_enqueue(nds_ctx *ctx, nds_dtype dtype, const void *data, int ndim,
                                      const size_t *shape, const ptrdiff_t *strides_elems, int axis,
                                      const double *qs, size_t nq, nds_interp interp, int skipnan, void *out,
                                      size_t *bad_q_index);

/*
 * skipnan over Option<T> elements (impl MaybeNan for Option<i32> ..., src/maybe_nan/mod.rs:152-213):
 * `valid` is a byte mask with the same shape and element strides as `data` (0 = None).  `out_valid`
 * (C-order, result shape) receives 0 for lanes that held no Some value (the reference returns None).
 */
nds_status nds_quantiles_axis_masked(nds_ctx *ctx, nds_dtype dtype, const void *data, const uint8_t *valid,
                                     int ndim, const size_t *shape, const ptrdiff_t *strides_elems, int axis,
                                     const double *qs, size_t nq, nds_interp interp, void *out,
                                     uint8_t *out_valid, size_t *bad_q_index);

/*
 * Sort1dExt::get_many_from_sorted_mut   src/sort.rs:121-131,184-202  (k indexes, any order, duplicates allowed)
 * Sort1dExt::get_from_sorted_mut        src/sort.rs:99-120           (k = 1)
 * `out_values[j]` is the element that would sit at `indexes[j]` in the sorted array; the shim sorts and
 * de-duplicates the pairs to rebuild the reference's IndexMap (ascending keys).  Any index >= n ->
 * NDS_INDEX_OUT_OF_BOUNDS (the reference panics).  n == 0 with k == 0 -> NDS_OK.
 */
nds_status nds_get_many_from_sorted(nds_ctx *ctx, nds_dtype dtype, const void *data, size_t n,
                                    ptrdiff_t stride_elems, const size_t *indexes, size_t k, void *out_values);

/*
 * Sort1dExt::partition_mut               src/sort.rs:133-168 (post-conditions: tests/sort.rs:5-33)
 * Partitions the 1-D array IN PLACE around the value at `pivot_index` and stores the pivot's new position in
 * `*partition_index`: every element before it is < the pivot, every element after it is >= the pivot.  Index AND
 * arrangement are those of the reference's sequential Hoare walk, bit for bit (both follow from the input alone:
 * nds_partition.cu gives the closed form).  `data` may be a host array (staged, copied back) or device memory.
 * pivot_index >= n -> NDS_INDEX_OUT_OF_BOUNDS (the reference panics, src/sort.rs:77-78).  n == 1 -> index 0, nothing
 * moves (the reference's walk runs off a one-element array).  Floats stand for N32 / N64: no NaN.  Synchronous.
 */
nds_status nds_partition(nds_ctx *ctx, nds_dtype dtype, void *data, size_t n, ptrdiff_t stride_elems,
                         size_t pivot_index, size_t *partition_index);

/*
 * QuantileExt::{argmin, argmax, min, max}                          src/quantile/mod.rs:283-300, 320-333, 347-363, 384-397
 * QuantileExt::{argmin_skipnan, argmax_skipnan, min_skipnan, max_skipnan}   mod.rs:302-318, 335-345, 365-382, 399-409
 * over the WHOLE array (any shape / strides).  `out_value` (host, one element, optional) receives the extremum,
 * `out_index` (host, ndim entries, optional) its coordinates (`D::Pattern`).  Among equal extrema the first in logical
 * (C) order is reported, as the reference's strict comparisons do; -0.0 and +0.0 are equal.  One exception, as in the
 * reference: max_skipnan (want_max, skipnan, out_index == NULL) folds with `acc.max(elem)` and Ord::max returns its second
 * operand on a tie, so the LAST of equal maxima is returned (max_skipnan([+0.0, -0.0]) is -0.0).
 *   NDS_EMPTY_INPUT      the array has no element (MinMaxError::EmptyInput / EmptyInput), or -- skipnan -- none that is
 *                        not NaN (arg*_skipnan: EmptyInput; min/max_skipnan: `out_value` holds the quiet NaN they return)
 *   NDS_UNDEFINED_ORDER  skipnan == 0 and the array holds a NaN (MinMaxError::UndefinedOrder)
 * For integer dtypes skipnan is a no-op.  Synchronous.
 */
nds_status nds_minmax(nds_ctx *ctx, nds_dtype dtype, const void *data, int ndim, const size_t *shape,
                      const ptrdiff_t *strides_elems, int want_max, int skipnan, void *out_value, size_t *out_index);

/* HistogramExt::histogram (src/histogram/histograms.rs:137-147; Grid::index_of, src/histogram/grid.rs:203-218;
 * Edges::indices_of, src/histogram/bins.rs:217-229): the n-dimensional histogram of the rows of an
 * (n_obs x n_dims) matrix -- element strides stride_obs between observations, stride_dim between the coordinates of one
 * -- over a grid given by per-dimension SORTED edges (`edges`: the n_dims edge arrays one after the other, n_edges[d]
 * of them for dimension d; bins are left-closed and right-open, dimension d has n_edges[d] - 1 of them).  An observation
 * outside any projection is skipped, as the reference does (add_observation's BinNotFound is dropped, histograms.rs:141).
 * `counts` (host or device) receives the C-order array of prod(n_edges[d] - 1) 64-bit counts.  Synchronous. */
nds_status nds_histogram(nds_ctx *ctx, nds_dtype dtype, const void *data, size_t n_obs, size_t n_dims,
                         ptrdiff_t stride_obs, ptrdiff_t stride_dim, const void *edges, const size_t *n_edges,
                         unsigned long long *counts);

/* ---- synthetic benchmark inputs ---------------------------------------- */

/* Counter-based splitmix64 streams: element i is a function of mix(seed + first_index + i) only, so bench.py's
 * device-resident inputs and the host arrays of its CPU reference arm hold the same bytes (tests/test_synth.py
 * restates the streams in numpy).  Not part of the reference's surface: bench / test tooling that ships with the
 * library so that the inputs are generated at HBM speed. */
typedef enum nds_synth_kind {
    NDS_SYNTH_F64_UNIFORM = 0,       /* (z >> 11) * 2^-53 */
    NDS_SYNTH_F32_UNIFORM = 1,       /* (z >> 40) * 2^-24 */
    NDS_SYNTH_F32_UNIFORM_NAN10 = 2, /* the same, NaN (0x7FC00000) where mix(0x9E37 + seed + i) % 10 == 0 */
    NDS_SYNTH_U32_16VALUES = 3,      /* (z & 15) * 0x11111111 */
    NDS_SYNTH_U32_16VALUES_ADV = 4,  /* 0x80000000 + (z & 15) */
    NDS_SYNTH_I64_16VALUES = 5       /* ((z & 15) - 8) * 2^40 */
} nds_synth_kind;
/* Fills n elements at dst_device (device memory) on the ctx stream; asynchronous. */
nds_status nds_synth_fill(nds_ctx *ctx, nds_synth_kind kind, void *dst_device, size_t n, uint64_t seed,
                          uint64_t first_index);

/* ---- one 1-D array sharded over several GPUs (one process per GPU) ------------------------ */

#define NDS_UNIQUE_ID_BYTES 128
/* Rank 0 creates the id and ships it to the others out of band (torch.distributed store, MPI, file). */
nds_status nds_comm_get_unique_id(void *id_out /* NDS_UNIQUE_ID_BYTES */);
nds_status nds_comm_init_rank(nds_ctx *ctx, int nranks, int rank, const void *id);
nds_status nds_comm_destroy(nds_ctx *ctx);
/* How the ranks of this context exchange data in sharded calls, and how much of it has happened so far:
 * out[0] = 1 when the exchange steps run as our own kernels over peer memory (every rank has mapped every peer's inbox
 * at nds_comm_init_rank: NVLink / NVSwitch P2P), 0 when they are NCCL collectives; out[1] = exchange steps enqueued so
 * far; out[2] = host round trips inside sharded calls so far; out[3] = ranks. */
nds_status nds_comm_stats(nds_ctx *ctx, uint64_t out[4]);
/*
 * Quantile1dExt::quantiles_mut (src/quantile/mod.rs:623-633) over the concatenation of every rank's
 * shard.  Collective: all ranks call with the same qs / interp / skipnan (1, 2, 4 or 8 ranks take the one-pass
 * bracket select, other counts the radix passes).  Every rank streams its own shard once; what is exchanged on the
 * ctx stream are counts (element count, sample bucket counts, class counts, digit histograms) and a few thousand keys
 * (the common splitter sample, the last survivors) -- written straight into the peers' memory over NVLink by the
 * library's own kernels when the GPUs can map each other (nds_comm_stats), NCCL collectives otherwise; the elements of
 * the array never leave their GPU.  Every rank receives the full result in `out` (nq values).
 */
nds_status nds_quantiles_1d_sharded(nds_ctx *ctx, nds_dtype dtype, const void *local_shard, size_t n_local,
                                    const double *qs, size_t nq, nds_interp interp, int skipnan, void *out,
                                    size_t *bad_q_index);

#ifdef __cplusplus
}
#endif
#endif /* NDSTATS_B200_H */
