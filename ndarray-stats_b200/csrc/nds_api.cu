// C ABI of libndstats_b200.so (include/ndstats_b200.h): argument validation with the reference's
// error precedence, host<->device staging, dispatch to the selection paths.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "nds_internal.h"

using namespace nds;

namespace {

const size_t kDtypeSize[10] = {1, 2, 4, 8, 1, 2, 4, 8, 4, 8};

nds_status fail(nds_ctx *ctx, nds_status st, const std::string &msg) {
    if (ctx) ctx->last_error = msg;
    return st;
}
nds_status cuda_fail(nds_ctx *ctx, cudaError_t e, const char *where) {
    char buf[512];
    std::snprintf(buf, sizeof buf, "%s: %s", where, cudaGetErrorString(e));
    if (ctx) ctx->last_error = buf;
    return NDS_CUDA_ERROR;
}
#define CUDA_TRY(expr)                                              \
    do {                                                            \
        cudaError_t e__ = (expr);                                   \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #expr);  \
    } while (0)

size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// clear some bits of the deferred status word in stream order (the other bits belong to calls still in flight)
__global__ void status_clear_bits_kernel(int *status, int bits) { atomicAnd(status, ~bits); }

bool is_device_pointer(const void *p) {
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

// span of elements touched by a strided view, relative to the logical origin
void view_span(int ndim, const size_t *shape, const ptrdiff_t *strides, long long &min_off, long long &max_off) {
    min_off = 0;
    max_off = 0;
    for (int d = 0; d < ndim; ++d) {
        if (shape[d] == 0) continue;
        long long ext = (long long)(shape[d] - 1) * (long long)strides[d];
        if (ext < 0) min_off += ext;
        else max_off += ext;
    }
}

struct Parsed {
    LaneGeom geom;
    size_t out_elems;
    bool empty_result;
};

// Validation in the reference's order (src/quantile/mod.rs:440-455 / 522-528).
nds_status parse_call(nds_ctx *ctx, int dtype, int ndim, const size_t *shape, const ptrdiff_t *strides, int axis,
                      const double *qs, size_t nq, int interp, size_t *bad_q, Parsed &p) {
    if (!ctx) return NDS_BAD_ARGUMENT;
    if (dtype < NDS_I8 || dtype > NDS_F64) return fail(ctx, NDS_UNSUPPORTED, "unknown dtype");
    if (ndim < 1 || ndim > NDS_MAX_DIMS) return fail(ctx, NDS_BAD_ARGUMENT, "ndim out of range");
    if (!shape || !strides) return fail(ctx, NDS_BAD_ARGUMENT, "null shape/strides");
    if (axis < 0 || axis >= ndim) return fail(ctx, NDS_BAD_ARGUMENT, "axis out of bounds (the reference panics)");
    if (interp < NDS_LOWER || interp > NDS_LINEAR) return fail(ctx, NDS_BAD_ARGUMENT, "unknown interpolation");
    if (nq && !qs) return fail(ctx, NDS_BAD_ARGUMENT, "null qs");
    if (nq > (size_t)1 << 24) return fail(ctx, NDS_UNSUPPORTED, "more than 2^24 quantiles");
    for (size_t t = 0; t < nq; ++t) {
        if (!((qs[t] >= 0.) && (qs[t] <= 1.))) {
            if (bad_q) *bad_q = t;
            char buf[96];
            std::snprintf(buf, sizeof buf, "%g is not between 0. and 1. (inclusive).", qs[t]);
            return fail(ctx, NDS_INVALID_QUANTILE, buf);
        }
    }
    if (shape[axis] == 0) return fail(ctx, NDS_EMPTY_INPUT, "Empty input.");
    LaneGeom &g = p.geom;
    std::memset(&g, 0, sizeof g);
    g.axis_len = shape[axis];
    g.axis_stride = strides[axis];
    g.nlanes = 1;
    g.n_outer = 0;
    long long ostride = 1;
    // C-order result strides with shape[axis] = nq
    std::vector<long long> out_stride(ndim);
    for (int d = ndim - 1; d >= 0; --d) {
        out_stride[d] = ostride;
        ostride *= (long long)(d == axis ? nq : shape[d]);
    }
    g.out_axis_stride = out_stride[axis];
    for (int d = 0; d < ndim; ++d) {
        if (d == axis) continue;
        g.oshape[g.n_outer] = shape[d];
        g.in_ostride[g.n_outer] = strides[d];
        g.out_ostride[g.n_outer] = out_stride[d];
        g.nlanes *= shape[d];
        g.n_outer++;
    }
    p.out_elems = (size_t)g.nlanes * nq;
    p.empty_result = p.out_elems == 0;
    return NDS_OK;
}

nds_status status_from_bits(nds_ctx *ctx, int bits) {
    // STATUS_BIT_FALLBACK is consumed by nds_ctx_synchronize before this is called
    if (bits & STATUS_BIT_XCHG) return fail(ctx, NDS_NCCL_ERROR, "peer-memory exchange timed out waiting for another rank");
    if (bits & STATUS_BIT_NAN) return fail(ctx, NDS_NAN_IN_INPUT, "NaN in a non-skipnan float input (N32/N64 cannot hold NaN)");
    if (bits & STATUS_BIT_CAST_OVERFLOW)
        return fail(ctx, NDS_CAST_OVERFLOW, "integer Linear interpolation: T::from_f64 out of range (the reference panics)");
    return NDS_OK;
}

// Geometry of the layout stage's scratch copy of the lanes of `g`: lane l (C order over the outer dims) starts at
// l * pitch, elements are contiguous; the result's geometry is unchanged.
LaneGeom gathered_geom(const LaneGeom &g, size_t esize, unsigned long long &pitch) {
    LaneGeom o = g;
    const unsigned long long vec = esize ? 16 / esize : 1;
    pitch = vec ? (g.axis_len + vec - 1) / vec * vec : g.axis_len;
    o.axis_stride = 1;
    long long acc = (long long)pitch;
    for (int d = g.n_outer - 1; d >= 0; --d) {
        o.in_ostride[d] = acc;
        acc *= (long long)g.oshape[d];
    }
    return o;
}

// Grid-wide selection of ONE lane at a time (bracket select, else the radix passes: every requested rank per pass) against
// the CTA-per-lane kernel (one rank at a time, one CTA per lane).  `fast_lanes`: a bit-sliced lane kernel takes the call.
// Neither of the two general paths is right for every shape, so a rough cost model picks: the CTA kernel scans a lane
// once per distinct rank and key byte at ~25 GB/s per CTA (shared-memory atomics) and at most ~3 TB/s over all CTAs;
// the grid-wide paths pay ~150 us of launches per lane and then stream it at ~3 TB/s, 1.4 times (bracket select) or once
// per key byte (radix passes).
bool use_long_path(const nds_ctx *ctx, const LaneGeom &g, size_t nq, int interp, size_t esize, bool fast_lanes, bool mid_ok,
                   int mid_cluster, bool mid_gather) {
    if (ctx->forced_path == "long_radix") return true;
    if (ctx->forced_path != "auto") return false;
    if (g.axis_len >= (1ull << 32)) return true;   // the lane kernels index with 32 bits
    if (fast_lanes) return false;
    if (g.axis_len < 8192) return false;
    const double n = (double)g.axis_len, lanes = (double)g.nlanes, es = (double)esize;
    const double ranks = (double)nq * ((interp == NDS_LOWER || interp == NDS_HIGHER) ? 1.0 : 2.0);
    if (mid_ok) {
        // the CTA-per-lane bit-sliced kernel reads a lane once at ~40 GB/s per SM and then walks ~2.5 us per rank; only
        // a few lanes with many ranks are quicker one after the other over the whole grid
        const double cl = (double)mid_cluster;   // CTAs that share a lane
        const double rounds = std::ceil(lanes / std::floor((double)ctx->sm_count / cl));
        double t_mid = rounds * (5e-6 + n * es / (40e9 * cl) + (double)nq * (cl > 1 ? 5e-6 : 2.5e-6));
        if (mid_gather) t_mid += 5e-6 + 2.0 * lanes * n * es / 2e12;   // the layout stage: one more read and one write
        const bool bracket_like = g.axis_stride == 1 && esize >= 4 && g.axis_len >= (1ull << 16) && nq <= 16;
        const double t_grid = lanes * (150e-6 + n * es * (bracket_like ? 1.4 : es) / 3e12);
        return t_grid < t_mid;
    }
    const double conc = std::min(lanes, (double)ctx->sm_count * 8.0);
    const double bw_cta = std::min(25e9, 3e12 / conc);
    const double t_cta = std::ceil(lanes / conc) * ranks * es * n * es / bw_cta;
    const bool bracket_like = g.axis_stride == 1 && esize >= 4 && g.axis_len >= (1ull << 16) && nq <= 16;
    const double t_long = lanes * (150e-6 + n * es * (bracket_like ? 1.4 : es) / 3e12);
    return t_long < t_cta;
}

// Common body once everything is parsed. `host_io`: data/out (and masks) are host pointers to stage.
nds_status run_call(nds_ctx *ctx, nds_dtype dtype, const void *data, const uint8_t *valid, int ndim,
                    const size_t *shape, const ptrdiff_t *strides, const Parsed &p, const double *qs,
                    const unsigned long long *ranks, size_t nq, int interp, int skipnan, void *out,
                    uint8_t *out_valid, bool synchronous, bool sharded = false) {
    const size_t esize = kDtypeSize[dtype];
    CUDA_TRY(cudaSetDevice(ctx->device));
    const bool data_on_device = data ? is_device_pointer(data) : true;  // an empty shard has no buffer
    const bool out_on_device = is_device_pointer(out);
    if (!synchronous && (!data_on_device || !out_on_device))
        return fail(ctx, NDS_BAD_ARGUMENT, "enqueue entry points need device-resident data and out");
    if (valid && is_device_pointer(valid) != data_on_device)
        return fail(ctx, NDS_BAD_ARGUMENT, "data and valid must live in the same memory space");
    if (out_valid && is_device_pointer(out_valid) != out_on_device)
        return fail(ctx, NDS_BAD_ARGUMENT, "out and out_valid must live in the same memory space");

    long long min_off = 0, max_off = 0;
    view_span(ndim, shape, strides, min_off, max_off);
    const size_t span_elems = (size_t)(max_off - min_off + 1);

    bool fast_lanes = false, mid_ok = false, mid_gather = false;
    int mid_cluster = 1;
    unsigned long long gather_pitch = 0;
    if (!sharded && ctx->forced_path == "auto") {
        DeviceCall pre;
        pre.dtype = dtype;
        pre.data = data_on_device ? data : (const void *)nullptr;  // (a staged host array is 256-byte aligned)
        pre.valid = valid;
        pre.out_valid = out_valid;
        pre.geom = p.geom;
        pre.q.nq = (int)nq;
        fast_lanes = short_bitslice_supported(ctx, pre) || cols_bitslice_supported(ctx, pre);
        mid_ok = !fast_lanes && mid_bitslice_supported(ctx, pre);
        if (!fast_lanes && !mid_ok) {   // through the layout stage?
            pre.geom = gathered_geom(p.geom, esize, gather_pitch);
            pre.lanes_padded = true;
            mid_gather = mid_ok = (esize == 4 || esize == 8) && mid_bitslice_supported(ctx, pre);
        }
        if (mid_ok) mid_cluster = mid_bitslice_cluster(ctx, pre);
    } else if (!sharded && ctx->forced_path == "gather") {
        DeviceCall pre;
        pre.dtype = dtype;
        pre.data = nullptr;
        pre.valid = valid;
        pre.out_valid = out_valid;
        pre.geom = gathered_geom(p.geom, esize, gather_pitch);
        pre.lanes_padded = true;
        pre.q.nq = (int)nq;
        mid_gather = (esize == 4 || esize == 8) && mid_bitslice_supported(ctx, pre);
        if (!mid_gather) return fail(ctx, NDS_UNSUPPORTED, "forced path gather does not support this call");
    }
    const bool long_path = sharded || use_long_path(ctx, p.geom, nq, interp, esize, fast_lanes, mid_ok, mid_cluster, mid_gather);
    if (long_path) mid_gather = false;
    // ---- carve the arena ---------------------------------------------------------------------------
    size_t off = 0;
    auto carve = [&](size_t bytes) {
        size_t o = off;
        off += align_up(bytes);
        return o;
    };
    const size_t o_qs = carve(nq * sizeof(double));
    const size_t o_in = data_on_device ? 0 : carve(span_elems * esize);
    const size_t o_valid = (valid && !data_on_device) ? carve(span_elems) : 0;
    const size_t o_out = out_on_device ? 0 : carve(p.out_elems * esize);
    const size_t o_outv = (out_valid && !out_on_device) ? carve(p.out_elems) : 0;
    const size_t o_long = long_path ? carve(long_scratch_bytes((int)nq)) : 0;
    const size_t o_gather = mid_gather ? carve((size_t)p.geom.nlanes * gather_pitch * esize) : 0;
    // bracket select (one streaming pass) is tried first on long lanes it supports
    DeviceCall probe;
    probe.dtype = dtype;
    probe.valid = valid;
    probe.out_valid = out_valid;
    probe.geom = p.geom;
    probe.q.nq = (int)nq;
    {
        double w = 0.0;
        const double c = brk_csigma();
        const double nm1 = p.geom.axis_len > 1 ? (double)(p.geom.axis_len - 1) : 1.0;
        for (size_t t = 0; t < nq; ++t) {
            double q = ranks ? (double)ranks[t] / nm1 : qs[t];
            q = q < 0. ? 0. : (q > 1. ? 1. : q);
            w += 2.0 * c * std::sqrt(q * (1.0 - q)) + 1e-3;
        }
        probe.brk_wsum = w;
    }
    const bool want_bracket = ctx->forced_path == "auto" || ctx->forced_path == "bracket";
    const bool use_bracket = (long_path || ctx->forced_path == "bracket") && want_bracket &&
                             bracket_supported(ctx, probe, sharded, ctx->forced_path == "bracket");
    if (ctx->forced_path == "bracket" && !use_bracket)
        return fail(ctx, NDS_UNSUPPORTED, "forced path bracket does not support this call");
    const size_t o_brk = use_bracket ? carve(bracket_scratch_bytes(probe, sharded)) : 0;
    const size_t o_long_fb = (use_bracket && !long_path) ? carve(long_scratch_bytes((int)nq)) : o_long;
    const size_t o_ntotal = sharded ? carve(sizeof(unsigned long long)) : 0;
    CUDA_TRY(arena_reserve(ctx, off));
    char *arena = (char *)ctx->arena;

    DeviceCall c;
    c.dtype = dtype;
    c.geom = p.geom;
    c.brk_wsum = probe.brk_wsum;
    c.q.qs = ranks ? nullptr : (const double *)(arena + o_qs);
    c.q.ranks = ranks ? (const unsigned long long *)(arena + o_qs) : nullptr;
    c.q.nq = (int)nq;
    c.q.interp = interp;
    c.q.skipnan = skipnan;
    c.q.status = ctx->d_status;
    c.q.inline_q = 0;
    if (!ranks && nq <= (size_t)NDS_Q_INLINE) {
        c.q.inline_q = 1;
        for (size_t t = 0; t < nq; ++t) c.q.qv[t] = qs[t];
    } else {
        CUDA_TRY(cudaMemcpyAsync(arena + o_qs, ranks ? (const void *)ranks : (const void *)qs, nq * sizeof(double),
                                 cudaMemcpyHostToDevice, ctx->stream));
    }
    if (data_on_device) {
        c.data = data;
        c.valid = valid;
    } else {
        const char *src = (const char *)data + min_off * (long long)esize;
        CUDA_TRY(stage_h2d(ctx, arena + o_in, src, span_elems * esize));
        c.data = arena + o_in - min_off * (long long)esize;
        c.valid = nullptr;
        if (valid) {
            CUDA_TRY(stage_h2d(ctx, arena + o_valid, valid + min_off, span_elems));
            c.valid = (const uint8_t *)(arena + o_valid) - min_off;
        }
    }
    c.out = out_on_device ? out : (void *)(arena + o_out);
    c.out_valid = out_valid ? (out_on_device ? out_valid : (uint8_t *)(arena + o_outv)) : nullptr;

    // ---- dispatch ----------------------------------------------------------------------------------
    cudaError_t e;
    unsigned long long bracket_n_total = 0;
    bool bracket_done = false;
    const bool want_short = ctx->forced_path == "auto" || ctx->forced_path == "short_bitslice";
    const bool want_cols = ctx->forced_path == "auto" || ctx->forced_path == "cols_bitslice";
    if (ctx->forced_path == "short_bitslice" && !short_bitslice_supported(ctx, c))
        return fail(ctx, NDS_UNSUPPORTED, "forced path short_bitslice does not support this call");
    if (ctx->forced_path == "cols_bitslice" && !cols_bitslice_supported(ctx, c))
        return fail(ctx, NDS_UNSUPPORTED, "forced path cols_bitslice does not support this call");
    const bool want_mid = ctx->forced_path == "auto" || ctx->forced_path == "mid_bitslice";
    if (ctx->forced_path == "mid_bitslice" && !mid_bitslice_supported(ctx, c))
        return fail(ctx, NDS_UNSUPPORTED, "forced path mid_bitslice does not support this call");
    if (use_bracket) {
        int failed = 0;
        e = launch_bracket(ctx, c, arena + o_brk, sharded, synchronous, &failed, &bracket_n_total);
        bracket_done = e == cudaSuccess && synchronous && !failed;
        if (e == cudaSuccess && synchronous && failed) {
            // a rank escaped its sampled bracket (or scratch ran out): the exact multi-pass radix select takes over
            // (only the give-up bit: NaN / overflow bits of earlier, not yet synchronised enqueued calls must survive)
            status_clear_bits_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_status, STATUS_BIT_FALLBACK);
            ctx->launches += 1;
            e = launch_long_radix(ctx, c, arena + o_long_fb, sharded);
            ctx->last_path = "bracket->long_radix";
        } else if (e == cudaSuccess && !synchronous) {
            if (!ctx->pending) ctx->pending = new std::vector<PendingCall>();
            PendingCall pc;
            pc.call = c;
            pc.is_ranks = ranks != nullptr;
            pc.words.resize(nq);
            std::memcpy(pc.words.data(), ranks ? (const void *)ranks : (const void *)qs, nq * 8);
            ctx->pending->push_back(std::move(pc));
        }
    } else if (long_path) {
        e = launch_long_radix(ctx, c, arena + o_long, sharded);
    } else if (want_short && short_bitslice_supported(ctx, c)) {
        e = launch_short_bitslice(ctx, c);
    } else if (want_cols && cols_bitslice_supported(ctx, c)) {
        e = launch_cols_bitslice(ctx, c);
    } else if (want_mid && mid_bitslice_supported(ctx, c)) {
        e = launch_mid_bitslice(ctx, c);
    } else if (mid_gather) {
        // layout stage: the lanes gathered into aligned, padded, contiguous scratch, then the CTA-per-lane kernel
        e = launch_gather_lanes(ctx, c, esize, arena + o_gather, gather_pitch);
        if (e == cudaSuccess) {
            DeviceCall c2 = c;
            c2.data = arena + o_gather;
            c2.geom = gathered_geom(c.geom, esize, gather_pitch);
            c2.lanes_padded = true;
            e = launch_mid_bitslice(ctx, c2);
            ctx->last_path = "gather->mid_bitslice";
        }
    } else {
        e = launch_generic_cta(ctx, c);
    }
    if (e != cudaSuccess) return cuda_fail(ctx, e, ctx->last_path);

    if (!out_on_device) {
        CUDA_TRY(cudaMemcpyAsync(out, c.out, p.out_elems * esize, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_valid) CUDA_TRY(cudaMemcpyAsync(out_valid, c.out_valid, p.out_elems, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (sharded) {
        // global element count (summed over ranks in the first pass): 0 => QuantileError::EmptyInput
        unsigned long long n_total = bracket_n_total;
        if (!bracket_done) {
            CUDA_TRY(cudaMemcpyAsync(arena + o_ntotal, long_total_count_ptr(arena + o_long_fb, (int)nq),
                                     sizeof n_total, cudaMemcpyDeviceToDevice, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(&n_total, arena + o_ntotal, sizeof n_total, cudaMemcpyDeviceToHost, ctx->stream));
        }
        nds_status st = nds_ctx_synchronize(ctx);
        if (n_total == 0) return fail(ctx, NDS_EMPTY_INPUT, "Empty input.");
        return st;
    }
    if (synchronous) return nds_ctx_synchronize(ctx);
    return NDS_OK;
}

}  // namespace

namespace nds {
cudaError_t arena_reserve(nds_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->arena_bytes) return cudaSuccess;
    if (ctx->arena) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return e;
        cudaFree(ctx->arena);
        ctx->arena = nullptr;
        ctx->arena_bytes = 0;
    }
    size_t want = bytes + bytes / 8 + (1 << 20);
    cudaError_t e = cudaMalloc(&ctx->arena, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&ctx->arena, want);
        if (e != cudaSuccess) return e;
    }
    ctx->arena_bytes = want;
    return cudaSuccess;
}
}  // namespace nds

extern "C" {

int nds_abi_version(void) { return NDS_ABI_VERSION; }

const char *nds_status_string(nds_status s) {
    switch (s) {
    case NDS_OK: return "ok";
    case NDS_EMPTY_INPUT: return "Empty input.";                               // src/errors.rs:128
    case NDS_INVALID_QUANTILE: return "q is not between 0. and 1. (inclusive).";  // src/errors.rs:129-131
    case NDS_NAN_IN_INPUT: return "NaN in input of a non-skipnan call";
    case NDS_UNSUPPORTED: return "unsupported dtype / layout";
    case NDS_CUDA_ERROR: return "CUDA error";
    case NDS_NCCL_ERROR: return "NCCL error";
    case NDS_BAD_ARGUMENT: return "bad argument";
    case NDS_CAST_OVERFLOW: return "integer interpolation overflow";
    case NDS_INDEX_OUT_OF_BOUNDS: return "index out of bounds";
    case NDS_UNDEFINED_ORDER: return "Undefined ordering between a tested pair of values.";  // src/errors.rs:31-33
    }
    return "unknown status";
}

nds_status nds_validate_quantile_args(int ndim, const size_t *shape, int axis, const double *qs, size_t nq,
                                      nds_interp interp, size_t *bad_q_index, size_t *result_elems) {
    if (ndim < 1 || ndim > NDS_MAX_DIMS || !shape) return NDS_BAD_ARGUMENT;
    if (axis < 0 || axis >= ndim) return NDS_BAD_ARGUMENT;
    if (interp < NDS_LOWER || interp > NDS_LINEAR) return NDS_BAD_ARGUMENT;
    if (nq && !qs) return NDS_BAD_ARGUMENT;
    for (size_t t = 0; t < nq; ++t) {
        if (!((qs[t] >= 0.) && (qs[t] <= 1.))) {
            if (bad_q_index) *bad_q_index = t;
            return NDS_INVALID_QUANTILE;
        }
    }
    if (shape[axis] == 0) return NDS_EMPTY_INPUT;
    size_t elems = nq;
    for (int d = 0; d < ndim; ++d)
        if (d != axis) elems *= shape[d];
    if (result_elems) *result_elems = elems;
    return NDS_OK;
}

size_t nds_dtype_size(nds_dtype t) { return (t >= NDS_I8 && t <= NDS_F64) ? kDtypeSize[t] : 0; }

nds_status nds_ctx_create(int device, nds_ctx **out) {
    if (!out) return NDS_BAD_ARGUMENT;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return NDS_CUDA_ERROR;  // no CPU fallback: the product path needs a GPU
    }
    if (device < 0 || device >= count) return NDS_BAD_ARGUMENT;
    nds_ctx *ctx = new (std::nothrow) nds_ctx();
    if (!ctx) return NDS_CUDA_ERROR;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) goto bad;
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) goto bad;
    ctx->stream = ctx->own_stream;
    if (cudaMalloc(&ctx->d_status, sizeof(int)) != cudaSuccess) goto bad;
    if (cudaMemset(ctx->d_status, 0, sizeof(int)) != cudaSuccess) goto bad;
    if (cudaMallocHost(&ctx->h_status, sizeof(int)) != cudaSuccess) goto bad;
    *ctx->h_status = 0;
    *out = ctx;
    return NDS_OK;
bad:
    cudaGetLastError();
    nds_ctx_destroy(ctx);
    return NDS_CUDA_ERROR;
}

void nds_ctx_destroy(nds_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->nccl_comm) nds_comm_destroy(ctx);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    stage_destroy(ctx);
    if (ctx->arena) cudaFree(ctx->arena);
    delete ctx->pending;
    if (ctx->d_status) cudaFree(ctx->d_status);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

nds_status nds_ctx_set_stream(nds_ctx *ctx, void *cuda_stream) {
    if (!ctx) return NDS_BAD_ARGUMENT;
    cudaStream_t next = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    if (next == ctx->stream) return NDS_OK;
    cudaStreamSynchronize(ctx->stream);  // the arena is stream-ordered: drain the old stream before switching
    ctx->stream = next;
    return NDS_OK;
}
void *nds_ctx_get_stream(nds_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

nds_status nds_ctx_synchronize(nds_ctx *ctx) {
    if (!ctx) return NDS_BAD_ARGUMENT;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(ctx->d_status, 0, sizeof(int), ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    int bits = *ctx->h_status;
    *ctx->h_status = 0;
    if (ctx->pending && !ctx->pending->empty()) {
        std::vector<PendingCall> todo;
        todo.swap(*ctx->pending);
        if (bits & STATUS_BIT_FALLBACK) {
            // some enqueued bracket-select call gave up; which one is not recorded, so every pending long-lane
            // call is re-run on the exact radix passes (rare: a 5-sigma event per bracket, or scratch exhaustion)
            for (PendingCall &pc : todo) {
                const size_t nq = pc.words.size();
                const size_t qbytes = (nq * 8 + 255) / 256 * 256;
                CUDA_TRY(arena_reserve(ctx, qbytes + long_scratch_bytes((int)nq)));
                char *arena = (char *)ctx->arena;
                CUDA_TRY(cudaMemcpyAsync(arena, pc.words.data(), nq * 8, cudaMemcpyHostToDevice, ctx->stream));
                pc.call.q.qs = pc.is_ranks ? nullptr : (const double *)arena;
                pc.call.q.ranks = pc.is_ranks ? (const unsigned long long *)arena : nullptr;
                cudaError_t e = launch_long_radix(ctx, pc.call, arena + qbytes, false);
                if (e != cudaSuccess) return cuda_fail(ctx, e, "long_radix fallback");
                ctx->last_path = "bracket->long_radix";
                CUDA_TRY(cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                CUDA_TRY(cudaMemsetAsync(ctx->d_status, 0, sizeof(int), ctx->stream));
                CUDA_TRY(cudaStreamSynchronize(ctx->stream));
                bits |= *ctx->h_status;
                *ctx->h_status = 0;
            }
        }
    }
    bits &= ~STATUS_BIT_FALLBACK;
    return status_from_bits(ctx, bits);
}

const char *nds_ctx_last_error(nds_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }
uint64_t nds_ctx_kernel_launches(nds_ctx *ctx) { return ctx ? ctx->launches : 0; }
const char *nds_ctx_last_path(nds_ctx *ctx) { return ctx ? ctx->last_path : "none"; }

nds_status nds_ctx_force_path(nds_ctx *ctx, const char *name) {
    if (!ctx) return NDS_BAD_ARGUMENT;
    std::string n = name ? name : "auto";
    if (n != "auto" && n != "generic_cta" && n != "long_radix" && n != "short_bitslice" && n != "cols_bitslice" &&
        n != "mid_bitslice" && n != "gather" && n != "bracket")
        return fail(ctx, NDS_BAD_ARGUMENT, "unknown path name");
    ctx->forced_path = n;
    return NDS_OK;
}

static nds_status quantiles_common(nds_ctx *ctx, nds_dtype dtype, const void *data, const uint8_t *valid, int ndim,
                                   const size_t *shape, const ptrdiff_t *strides, int axis, const double *qs,
                                   size_t nq, nds_interp interp, int skipnan, void *out, uint8_t *out_valid,
                                   size_t *bad_q, bool synchronous) {
    Parsed p;
    nds_status st = parse_call(ctx, dtype, ndim, shape, strides, axis, qs, nq, interp, bad_q, p);
    if (st != NDS_OK) return st;
    if (p.empty_result) return NDS_OK;
    if (!data || !out) return fail(ctx, NDS_BAD_ARGUMENT, "null data/out");
    return run_call(ctx, dtype, data, valid, ndim, shape, strides, p, qs, nullptr, nq, interp, skipnan, out,
                    out_valid, synchronous);
}

nds_status nds_quantiles_axis(nds_ctx *ctx, nds_dtype dtype, const void *data, int ndim, const size_t *shape,
                              const ptrdiff_t *strides_elems, int axis, const double *qs, size_t nq,
                              nds_interp interp, int skipnan, void *out, size_t *bad_q_index) {
    return quantiles_common(ctx, dtype, data, nullptr, ndim, shape, strides_elems, axis, qs, nq, interp, skipnan, out,
                            nullptr, bad_q_index, true);
}

nds_status nds_quantiles_axis_enqueue(nds_ctx *ctx, nds_dtype dtype, const void *data, int ndim,
                                      const size_t *shape, const ptrdiff_t *strides_elems, int axis,
                                      const double *qs, size_t nq, nds_interp interp, int skipnan, void *out,
                                      size_t *bad_q_index) {
    return quantiles_common(ctx, dtype, data, nullptr, ndim, shape, strides_elems, axis, qs, nq, interp, skipnan, out,
                            nullptr, bad_q_index, false);
}

nds_status nds_quantiles_axis_masked(nds_ctx *ctx, nds_dtype dtype, const void *data, const uint8_t *valid,
                                     int ndim, const size_t *shape, const ptrdiff_t *strides_elems, int axis,
                                     const double *qs, size_t nq, nds_interp interp, void *out,
                                     uint8_t *out_valid, size_t *bad_q_index) {
    if (ctx && (!valid || !out_valid)) {
        // still honour the reference's precedence for argument errors first
        Parsed p;
        nds_status st = parse_call(ctx, dtype, ndim, shape, strides_elems, axis, qs, nq, interp, bad_q_index, p);
        if (st != NDS_OK) return st;
        if (p.empty_result) return NDS_OK;
        return fail(ctx, NDS_BAD_ARGUMENT, "null valid/out_valid");
    }
    return quantiles_common(ctx, dtype, data, valid, ndim, shape, strides_elems, axis, qs, nq, interp, 1, out,
                            out_valid, bad_q_index, true);
}

nds_status nds_get_many_from_sorted(nds_ctx *ctx, nds_dtype dtype, const void *data, size_t n,
                                    ptrdiff_t stride_elems, const size_t *indexes, size_t k, void *out_values) {
    if (!ctx) return NDS_BAD_ARGUMENT;
    if (dtype < NDS_I8 || dtype > NDS_F64) return fail(ctx, NDS_UNSUPPORTED, "unknown dtype");
    if (k && !indexes) return fail(ctx, NDS_BAD_ARGUMENT, "null indexes");
    for (size_t j = 0; j < k; ++j)
        if (indexes[j] >= n) return fail(ctx, NDS_INDEX_OUT_OF_BOUNDS, "index >= n (the reference panics)");
    if (k == 0) return NDS_OK;
    if (!data || !out_values) return fail(ctx, NDS_BAD_ARGUMENT, "null data/out");
    Parsed p;
    std::memset(&p, 0, sizeof p);
    p.geom.n_outer = 0;
    p.geom.axis_len = n;
    p.geom.axis_stride = stride_elems;
    p.geom.out_axis_stride = 1;
    p.geom.nlanes = 1;
    p.out_elems = k;
    std::vector<unsigned long long> ranks(indexes, indexes + k);
    size_t shape1 = n;
    ptrdiff_t stride1 = stride_elems;
    return run_call(ctx, dtype, data, nullptr, 1, &shape1, &stride1, p, nullptr, ranks.data(), k, NDS_LOWER, 0,
                    out_values, nullptr, true);
}

nds_status nds_quantiles_1d_sharded(nds_ctx *ctx, nds_dtype dtype, const void *local_shard, size_t n_local,
                                    const double *qs, size_t nq, nds_interp interp, int skipnan, void *out,
                                    size_t *bad_q_index) {
    // validation shared with the single-GPU entry, except that an empty LOCAL shard is fine
    size_t shape1 = n_local ? n_local : 1;
    ptrdiff_t stride1 = 1;
    Parsed p;
    nds_status st = parse_call(ctx, dtype, 1, &shape1, &stride1, 0, qs, nq, interp, bad_q_index, p);
    if (st != NDS_OK) return st;
    if (!ctx->nccl_comm && ctx->nranks != 1) return fail(ctx, NDS_NCCL_ERROR, "no communicator");
    if (nq == 0) return NDS_OK;  // (an all-empty input with nq == 0 is EmptyInput in the reference; unreachable here)
    if ((n_local && !local_shard) || !out) return fail(ctx, NDS_BAD_ARGUMENT, "null shard/out");
    p.geom.axis_len = n_local;
    shape1 = n_local;
    return run_call(ctx, dtype, local_shard, nullptr, 1, &shape1, &stride1, p, qs, nullptr, nq, interp, skipnan, out,
                    nullptr, true, true);
}

}  // extern "C"
