// QuantileExt::{min, max, argmin, argmax} and the *_skipnan variants (src/quantile/mod.rs:283-415 of the reference):
// one streaming pass over the whole array, SURVEY.md §8(f) row N2.
//
// The reference walks the array in logical (C) order and replaces its current extremum only on a STRICT
// improvement (`partial_cmp(..) == Less` / `Greater`, mod.rs:293, 351; `m <= elem` keeps m, mod.rs:309, 368), so among
// equal extrema the FIRST in logical order wins, and -0.0 == +0.0 are equal.  Here: every thread keeps the best
// (key, logical index) of its elements, with the key of -0.0 folded onto +0.0; ties go to the smaller index; the
// element returned is the one stored at that index, bit for bit.  Any NaN makes the non-skipnan calls fail with
// UndefinedOrder (every element is compared, mod.rs:292, 321); the skipnan calls ignore NaNs and report "no value"
// when nothing else is left (EmptyInput for arg*, NaN for min/max_skipnan, mod.rs:314-318, 336-345).
#include <algorithm>

#include "nds_internal.h"

namespace nds {
namespace {

constexpr int MM_THREADS = 256;

struct MmGeom {
    int ndim;
    unsigned long long shape[NDS_MAX_DIMS];
    long long strides[NDS_MAX_DIMS];
    unsigned long long total;
    int contiguous;  // C-order, unit innermost stride: logical index == element offset
};

struct MmPartial {
    unsigned long long key, index, nan_count;
    int found, pad;
};

__device__ __forceinline__ long long mm_offset(const MmGeom &g, unsigned long long i) {
    if (g.contiguous) return (long long)i;
    long long off = 0;
    for (int d = g.ndim - 1; d >= 0; --d) {
        const unsigned long long q = i / g.shape[d];
        off += (long long)(i - q * g.shape[d]) * g.strides[d];
        i = q;
    }
    return off;
}

// key with -0.0 folded onto +0.0 (they compare equal in the reference)
template <typename T> __device__ __forceinline__ typename Traits<T>::Key mm_key(T x) {
    if (Traits<T>::is_float && x == T(0)) x = T(0);
    return Traits<T>::to_key(x);
}

// `tie_last`: among equal extrema the LAST in logical order wins (max_skipnan folds with `acc.max(elem)`, and Ord::max
// returns its second operand on a tie, src/quantile/mod.rs:399-409); everywhere else the first one does.
template <bool MAX> __device__ __forceinline__ bool mm_better(unsigned long long k, unsigned long long i, int f,
                                                              unsigned long long bk, unsigned long long bi, int bf, int tie_last) {
    if (!f) return false;
    if (!bf) return true;
    if (k != bk) return MAX ? k > bk : k < bk;
    return tie_last ? i > bi : i < bi;
}

template <typename T, bool MAX>
__global__ void __launch_bounds__(MM_THREADS) minmax_partial_kernel(const T *__restrict__ data, MmGeom g, MmPartial *partial, int tie_last) {
    using Tr = Traits<T>;
    constexpr int VEC = 16 / (int)sizeof(T);
    unsigned long long bk = 0, bi = 0, nans = 0;
    int bf = 0;
    auto visit = [&](T x, unsigned long long i) {
        if (Tr::is_nan(x)) {
            nans += 1;
            return;
        }
        const unsigned long long k = (unsigned long long)mm_key<T>(x);
        // a thread sees increasing i, so a tie never replaces -- unless the last of equal extrema is wanted
        if (!bf || (MAX ? k > bk : k < bk) || (tie_last && k == bk)) {
            bk = k;
            bi = i;
            bf = 1;
        }
    };
    const unsigned long long tid = (unsigned long long)blockIdx.x * MM_THREADS + threadIdx.x;
    const unsigned long long nthreads = (unsigned long long)gridDim.x * MM_THREADS;
    if (g.contiguous && (reinterpret_cast<uintptr_t>(data) % 16) == 0) {
        const unsigned long long nvec = g.total / VEC;
        const uint4 *v = reinterpret_cast<const uint4 *>(data);
#pragma unroll 4
        for (unsigned long long j = tid; j < nvec; j += nthreads) {
            const uint4 q = __ldg(v + j);
            T x[VEC];
            memcpy(x, &q, 16);
#pragma unroll
            for (int u = 0; u < VEC; ++u) visit(x[u], j * VEC + u);
        }
        for (unsigned long long i = nvec * VEC + tid; i < g.total; i += nthreads) visit(data[i], i);
    } else {
        for (unsigned long long i = tid; i < g.total; i += nthreads) visit(data[mm_offset(g, i)], i);
    }
    // warp, then CTA reduction of (found, key, index) and the NaN count
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long ok = __shfl_down_sync(0xffffffffu, bk, d), oi = __shfl_down_sync(0xffffffffu, bi, d);
        const int of = __shfl_down_sync(0xffffffffu, bf, d);
        nans += __shfl_down_sync(0xffffffffu, nans, d);
        if (mm_better<MAX>(ok, oi, of, bk, bi, bf, tie_last)) { bk = ok; bi = oi; bf = of; }
    }
    __shared__ MmPartial sp[MM_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sp[warp].key = bk; sp[warp].index = bi; sp[warp].found = bf; sp[warp].nan_count = nans; }
    __syncthreads();
    if (threadIdx.x == 0) {
        MmPartial r = sp[0];
        for (int w = 1; w < MM_THREADS / 32; ++w) {
            r.nan_count += sp[w].nan_count;
            if (mm_better<MAX>(sp[w].key, sp[w].index, sp[w].found, r.key, r.index, r.found, tie_last)) {
                r.key = sp[w].key; r.index = sp[w].index; r.found = sp[w].found;
            }
        }
        partial[blockIdx.x] = r;
    }
}

// result record (device): found, nan_count, linear logical index, the element itself
struct MmResult {
    unsigned long long index, nan_count, value_bits;
    int found, pad;
};

template <typename T, bool MAX>
__global__ void __launch_bounds__(MM_THREADS) minmax_final_kernel(const T *__restrict__ data, MmGeom g, const MmPartial *partial,
                                                                  int nparts, MmResult *res, int tie_last) {
    __shared__ MmPartial sp[MM_THREADS];
    MmPartial r;
    r.key = r.index = r.nan_count = 0;
    r.found = 0;
    for (int p = threadIdx.x; p < nparts; p += MM_THREADS) {
        const MmPartial q = partial[p];
        r.nan_count += q.nan_count;
        if (mm_better<MAX>(q.key, q.index, q.found, r.key, r.index, r.found, tie_last)) { r.key = q.key; r.index = q.index; r.found = q.found; }
    }
    sp[threadIdx.x] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 1; t < MM_THREADS; ++t) {
            r.nan_count += sp[t].nan_count;
            if (mm_better<MAX>(sp[t].key, sp[t].index, sp[t].found, r.key, r.index, r.found, tie_last)) {
                r.key = sp[t].key; r.index = sp[t].index; r.found = sp[t].found;
            }
        }
        MmResult out;
        out.found = r.found;
        out.nan_count = r.nan_count;
        out.index = r.index;
        out.value_bits = 0;
        if (r.found) {
            const T x = data[mm_offset(g, r.index)];
            memcpy(&out.value_bits, &x, sizeof(T));
        }
        *res = out;
    }
}

template <typename T>
cudaError_t launch_minmax_typed(nds_ctx *ctx, const void *data, const MmGeom &g, bool want_max, MmPartial *partial, int nparts,
                                MmResult *res, int tie_last) {
    if (want_max) {
        minmax_partial_kernel<T, true><<<nparts, MM_THREADS, 0, ctx->stream>>>((const T *)data, g, partial, tie_last);
        minmax_final_kernel<T, true><<<1, MM_THREADS, 0, ctx->stream>>>((const T *)data, g, partial, nparts, res, tie_last);
    } else {
        minmax_partial_kernel<T, false><<<nparts, MM_THREADS, 0, ctx->stream>>>((const T *)data, g, partial, tie_last);
        minmax_final_kernel<T, false><<<1, MM_THREADS, 0, ctx->stream>>>((const T *)data, g, partial, nparts, res, tie_last);
    }
    ctx->launches += 2;
    return cudaGetLastError();
}

}  // namespace
}  // namespace nds

using namespace nds;

extern "C" nds_status nds_minmax(nds_ctx *ctx, nds_dtype dtype, const void *data, int ndim, const size_t *shape,
                                 const ptrdiff_t *strides_elems, int want_max, int skipnan, void *out_value,
                                 size_t *out_index) {
    if (!ctx) return NDS_BAD_ARGUMENT;
    auto fail = [&](nds_status st, const char *msg) {
        ctx->last_error = msg;
        return st;
    };
    if (dtype < NDS_I8 || dtype > NDS_F64) return fail(NDS_UNSUPPORTED, "unknown dtype");
    if (ndim < 0 || ndim > NDS_MAX_DIMS || (ndim && (!shape || !strides_elems))) return fail(NDS_BAD_ARGUMENT, "bad shape");
    const size_t esize = nds_dtype_size(dtype);
    MmGeom g;
    g.ndim = ndim;
    g.total = 1;
    long long min_off = 0, max_off = 0;
    long long expect = 1;
    g.contiguous = 1;
    for (int d = ndim - 1; d >= 0; --d) {
        g.shape[d] = shape[d];
        g.strides[d] = strides_elems[d];
        if (shape[d] > 1 && strides_elems[d] != expect) g.contiguous = 0;
        expect *= (long long)shape[d];
        g.total *= shape[d];
    }
    for (int d = 0; d < ndim; ++d) {
        if (shape[d] == 0) continue;
        const long long ext = (long long)(shape[d] - 1) * (long long)strides_elems[d];
        if (ext < 0) min_off += ext;
        else max_off += ext;
    }
    const bool is_float = dtype == NDS_F32 || dtype == NDS_F64;
    auto write_nan = [&]() {
        if (!out_value) return;
        if (dtype == NDS_F32) { const uint32_t b = 0x7FC00000u; memcpy(out_value, &b, 4); }
        if (dtype == NDS_F64) { const uint64_t b = 0x7FF8000000000000ull; memcpy(out_value, &b, 8); }
    };
    if (g.total == 0) {  // mod.rs:287, 322 (EmptyInput); min/max_skipnan of nothing is NaN, mod.rs:336-345
        if (skipnan && is_float) write_nan();
        return fail(NDS_EMPTY_INPUT, "Empty input.");
    }
    if (!data) return fail(NDS_BAD_ARGUMENT, "null data");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return NDS_CUDA_ERROR;

    cudaPointerAttributes attr;
    bool on_device = cudaPointerGetAttributes(&attr, data) == cudaSuccess &&
                     (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    const int nparts = (int)std::min<unsigned long long>((g.total + MM_THREADS * 8 - 1) / (MM_THREADS * 8),
                                                         (unsigned long long)ctx->sm_count * 8);
    const size_t span = (size_t)(max_off - min_off + 1);
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t o_part = 0, o_res = up(sizeof(MmPartial) * (size_t)nparts), o_in = o_res + up(sizeof(MmResult));
    if (arena_reserve(ctx, o_in + (on_device ? 0 : up(span * esize))) != cudaSuccess) return fail(NDS_CUDA_ERROR, "arena");
    char *arena = (char *)ctx->arena;
    const void *dptr = data;
    if (!on_device) {
        const char *src = (const char *)data + min_off * (long long)esize;
        if (stage_h2d(ctx, arena + o_in, src, span * esize) != cudaSuccess)
            return fail(NDS_CUDA_ERROR, "H2D copy");
        dptr = arena + o_in - min_off * (long long)esize;
    }
    // max_skipnan (the value alone is asked for): the last of equal maxima, see mm_better
    const int tie_last = (want_max && skipnan && !out_index) ? 1 : 0;
    MmPartial *partial = (MmPartial *)(arena + o_part);
    MmResult *res = (MmResult *)(arena + o_res);
    cudaError_t e;
    switch (dtype) {
    case NDS_I8: e = launch_minmax_typed<int8_t>(ctx, dptr, g, want_max, partial, nparts, res, tie_last); break;
    case NDS_I16: e = launch_minmax_typed<int16_t>(ctx, dptr, g, want_max, partial, nparts, res, tie_last); break;
    case NDS_I32: e = launch_minmax_typed<int32_t>(ctx, dptr, g, want_max, partial, nparts, res, tie_last); break;
    case NDS_I64: e = launch_minmax_typed<int64_t>(ctx, dptr, g, want_max, partial, nparts, res, tie_last); break;
    case NDS_U8: e = launch_minmax_typed<uint8_t>(ctx, dptr, g, want_max, partial, nparts, res, tie_last); break;
    case NDS_U16: e = launch_minmax_typed<uint16_t>(ctx, dptr, g, want_max, partial, nparts, res, tie_last); break;
    case NDS_U32: e = launch_minmax_typed<uint32_t>(ctx, dptr, g, want_max, partial, nparts, res, tie_last); break;
    case NDS_U64: e = launch_minmax_typed<uint64_t>(ctx, dptr, g, want_max, partial, nparts, res, tie_last); break;
    case NDS_F32: e = launch_minmax_typed<float>(ctx, dptr, g, want_max, partial, nparts, res, tie_last); break;
    default: e = launch_minmax_typed<double>(ctx, dptr, g, want_max, partial, nparts, res, tie_last); break;
    }
    ctx->last_path = "minmax";
    if (e != cudaSuccess) return fail(NDS_CUDA_ERROR, cudaGetErrorString(e));
    MmResult h;
    if (cudaMemcpyAsync(&h, res, sizeof h, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return fail(NDS_CUDA_ERROR, "result copy");
    if (!skipnan && h.nan_count) return fail(NDS_UNDEFINED_ORDER, "Undefined ordering between a tested pair of values.");
    if (!h.found) {  // nothing but NaNs
        write_nan();
        return fail(NDS_EMPTY_INPUT, "Empty input.");
    }
    if (out_value) memcpy(out_value, &h.value_bits, esize);
    if (out_index) {  // D::Pattern: the coordinates of the logical index
        unsigned long long i = h.index;
        for (int d = ndim - 1; d >= 0; --d) {
            out_index[d] = (size_t)(i % shape[d]);
            i /= shape[d];
        }
    }
    return NDS_OK;
}
