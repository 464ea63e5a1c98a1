// Sort1dExt::partition_mut (src/sort.rs:133-168 of the reference) on the device, SURVEY.md §8(f) row N1.
//
// The reference is a sequential Hoare partition: the pivot goes to position 0, `i` walks up from 1 to the next element
// that is >= the pivot, `j` walks down from n - 1 to the next element that is < the pivot, the two are swapped, and so on
// until the walks meet; the pivot is then swapped into position i - 1, which is returned.  Both the returned index and
// the arrangement it leaves are functions of the input alone, and they have a closed form that needs no sequential walk:
//
//   m      = number of elements < pivot among positions 1 .. n-1 (after the pivot went to 0)      -> the returned index
//   L_k    = the k-th position <= m (ascending) whose element is >= pivot      ("misplaced on the left")
//   R_k    = the k-th position >  m (descending) whose element is <  pivot     ("misplaced on the right"; as many as L)
//   Hoare swaps exactly the pairs (L_k, R_k), k = 0 .. t-1, then position 0 with position m.
//
// (`i` can only stop on an L position and `j` on an R position while i < j, and they meet them in exactly that order.)
// So: one counting pass (m), one pass that writes the two position lists through a prefix sum of the "< pivot" flags,
// one pass that swaps the pairs.  tests/test_partition.py checks index AND arrangement against a faithful sequential
// restatement of the reference's walk, bit for bit.  Comparisons are the element type's own `<` (IEEE for floats, which stand for
// N32 / N64 here: -0.0 == +0.0, no NaN).
#include <algorithm>
#include <cstring>

#include "nds_internal.h"

namespace nds {
namespace {

constexpr int PT_THREADS = 256;
constexpr int PT_ITEMS = 8;                       // consecutive positions per thread and tile
constexpr int PT_TILE = PT_THREADS * PT_ITEMS;

struct PtHeader {
    unsigned long long m;        // elements < pivot = the partition index
    unsigned long long t;        // misplaced pairs
    unsigned long long pivot_bits;
};

template <typename T> __device__ __forceinline__ T pt_load(const T *a, long long stride, unsigned long long i) {
    return a[(long long)i * stride];
}

// the pivot goes to position 0 (src/sort.rs:137-138)
template <typename T> __global__ void part_begin_kernel(T *a, long long stride, unsigned long long pivot_index, PtHeader *hdr) {
    const T v = a[(long long)pivot_index * stride];
    a[(long long)pivot_index * stride] = a[0];
    a[0] = v;
    unsigned long long bits = 0;
    memcpy(&bits, &v, sizeof(T));
    hdr->pivot_bits = bits;
    hdr->m = 0;
    hdr->t = 0;
}

template <typename T> __device__ __forceinline__ T pt_pivot(const PtHeader *hdr) {
    const unsigned long long bits = hdr->pivot_bits;
    T v;
    memcpy(&v, &bits, sizeof(T));
    return v;
}

// CTA c owns positions [1 + c * chunk, 1 + (c + 1) * chunk) of [1, n)
template <typename T>
__global__ void __launch_bounds__(PT_THREADS) part_count_kernel(const T *__restrict__ a, long long stride, unsigned long long n,
                                                                unsigned long long chunk, const PtHeader *hdr,
                                                                unsigned long long *cta_less) {
    const T v = pt_pivot<T>(hdr);
    const unsigned long long p0 = 1ull + (unsigned long long)blockIdx.x * chunk;
    const unsigned long long p1 = min(p0 + chunk, n);
    unsigned cnt = 0;
    for (unsigned long long p = p0 + threadIdx.x; p < p1; p += PT_THREADS) cnt += pt_load(a, stride, p) < v ? 1u : 0u;
    __shared__ unsigned wsum[PT_THREADS / 32];
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int w = 0; w < PT_THREADS / 32; ++w) s += wsum[w];
        cta_less[blockIdx.x] = s;
    }
}

// exclusive prefix of the per-CTA counts (in place) and m
__global__ void __launch_bounds__(1024) part_scan_kernel(unsigned long long *cta_less, int nctas, PtHeader *hdr) {
    __shared__ unsigned long long part[1024];
    const int per = (nctas + 1023) / 1024;
    const int b = threadIdx.x * per, e = min(b + per, nctas);
    unsigned long long s = 0;
    for (int i = b; i < e; ++i) s += cta_less[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; ++i) {
            const unsigned long long x = part[i];
            part[i] = run;
            run += x;
        }
        hdr->m = run;
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (int i = b; i < e; ++i) {
        const unsigned long long x = cta_less[i];
        cta_less[i] = run;
        run += x;
    }
}

// second pass: the positions of the misplaced elements, in Hoare's meeting order
template <typename T, typename Pos>
__global__ void __launch_bounds__(PT_THREADS) part_lists_kernel(const T *__restrict__ a, long long stride, unsigned long long n,
                                                                unsigned long long chunk, PtHeader *hdr,
                                                                const unsigned long long *cta_less_before, Pos *lpos, Pos *rpos) {
    const T v = pt_pivot<T>(hdr);
    const unsigned long long m = hdr->m;
    const unsigned long long p0 = 1ull + (unsigned long long)blockIdx.x * chunk;
    const unsigned long long p1 = min(p0 + chunk, n);
    unsigned long long less_before = cta_less_before[blockIdx.x];   // "< pivot" among positions [1, tile start)
    __shared__ unsigned wsum[PT_THREADS / 32];
    __shared__ unsigned tile_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned my_l = 0;
    for (unsigned long long tile = p0; tile < p1; tile += PT_TILE) {
        const unsigned long long q0 = tile + (unsigned long long)threadIdx.x * PT_ITEMS;
        unsigned flags = 0;   // bit i: position q0 + i exists and its element is < pivot
        unsigned have = 0;
#pragma unroll
        for (int i = 0; i < PT_ITEMS; ++i) {
            const unsigned long long p = q0 + i;
            if (p < p1) {
                have |= 1u << i;
                if (pt_load(a, stride, p) < v) flags |= 1u << i;
            }
        }
        const unsigned cnt = __popc(flags);
        unsigned incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned x = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += x;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned run = 0;
            for (int w = 0; w < PT_THREADS / 32; ++w) {
                const unsigned x = wsum[w];
                wsum[w] = run;
                run += x;
            }
            tile_total = run;
        }
        __syncthreads();
        unsigned long long before = less_before + wsum[warp] + (incl - cnt);   // "< pivot" among [1, q0)
#pragma unroll
        for (int i = 0; i < PT_ITEMS; ++i) {
            if (!((have >> i) & 1u)) break;
            const unsigned long long p = q0 + i;
            const bool less = (flags >> i) & 1u;
            if (!less && p <= m) {            // the ((p - 1) - before)-th ">= pivot" element from the left
                lpos[(p - 1ull) - before] = (Pos)p;
                my_l += 1;
            } else if (less && p > m) {       // "< pivot" elements to its right: m - before - 1
                rpos[m - before - 1ull] = (Pos)p;
            }
            before += less ? 1u : 0u;
        }
        less_before += tile_total;
        __syncthreads();
    }
    my_l = __reduce_add_sync(0xffffffffu, my_l);
    if (lane == 0 && my_l) atomicAdd(&hdr->t, (unsigned long long)my_l);
}

template <typename T, typename Pos>
__global__ void __launch_bounds__(PT_THREADS) part_swap_kernel(T *a, long long stride, const PtHeader *hdr, const Pos *lpos,
                                                               const Pos *rpos) {
    const unsigned long long t = hdr->t;
    for (unsigned long long k = (unsigned long long)blockIdx.x * PT_THREADS + threadIdx.x; k < t;
         k += (unsigned long long)gridDim.x * PT_THREADS) {
        const long long i = (long long)lpos[k] * stride, j = (long long)rpos[k] * stride;
        const T x = a[i];
        a[i] = a[j];
        a[j] = x;
    }
}

// the pivot goes to its place (src/sort.rs:164-165)
template <typename T> __global__ void part_end_kernel(T *a, long long stride, const PtHeader *hdr) {
    const long long im = (long long)hdr->m * stride;
    const T x = a[0];
    a[0] = a[im];
    a[im] = x;
}

template <typename T, typename Pos>
cudaError_t partition_typed2(nds_ctx *ctx, T *a, long long stride, unsigned long long n, unsigned long long pivot_index,
                             PtHeader *hdr, unsigned long long *cta_less, int nctas, unsigned long long chunk, void *lists) {
    cudaStream_t s = ctx->stream;
    Pos *lpos = reinterpret_cast<Pos *>(lists);
    Pos *rpos = lpos + (n / 2 + 1);
    part_begin_kernel<T><<<1, 1, 0, s>>>(a, stride, pivot_index, hdr);
    part_count_kernel<T><<<nctas, PT_THREADS, 0, s>>>(a, stride, n, chunk, hdr, cta_less);
    part_scan_kernel<<<1, 1024, 0, s>>>(cta_less, nctas, hdr);
    part_lists_kernel<T, Pos><<<nctas, PT_THREADS, 0, s>>>(a, stride, n, chunk, hdr, cta_less, lpos, rpos);
    part_swap_kernel<T, Pos><<<ctx->sm_count * 8, PT_THREADS, 0, s>>>(a, stride, hdr, lpos, rpos);
    part_end_kernel<T><<<1, 1, 0, s>>>(a, stride, hdr);
    ctx->launches += 6;
    return cudaGetLastError();
}

}  // namespace
}  // namespace nds

using namespace nds;

extern "C" nds_status nds_partition(nds_ctx *ctx, nds_dtype dtype, void *data, size_t n, ptrdiff_t stride_elems,
                                    size_t pivot_index, size_t *partition_index) {
    if (!ctx) return NDS_BAD_ARGUMENT;
    auto fail = [&](nds_status st, const char *msg) {
        ctx->last_error = msg;
        return st;
    };
    if (dtype < NDS_I8 || dtype > NDS_F64) return fail(NDS_UNSUPPORTED, "unknown dtype");
    if (pivot_index >= n) return fail(NDS_INDEX_OUT_OF_BOUNDS, "pivot_index >= n (the reference panics)");
    if (!data || !partition_index) return fail(NDS_BAD_ARGUMENT, "null data / partition_index");
    if (n == 1) {  // nothing to move (the reference's walk runs off the array here: j = 0 is decremented, src/sort.rs:152-157)
        *partition_index = 0;
        return NDS_OK;
    }
    if (cudaSetDevice(ctx->device) != cudaSuccess) return NDS_CUDA_ERROR;
    static const size_t esizes[] = {1, 2, 4, 8, 1, 2, 4, 8, 4, 8};
    const size_t esize = esizes[dtype];
    cudaPointerAttributes attr;
    const bool on_device = cudaPointerGetAttributes(&attr, data) == cudaSuccess &&
                           (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
    cudaGetLastError();

    const unsigned long long nn = n;
    const unsigned long long tiles = (nn - 1 + PT_TILE - 1) / PT_TILE;
    const int nctas = (int)std::min<unsigned long long>(tiles, (unsigned long long)ctx->sm_count * 8);
    const unsigned long long chunk = (tiles + nctas - 1) / nctas * PT_TILE;
    const bool wide = nn > 0xffffffffull;
    const size_t pos_bytes = wide ? 8 : 4;
    const long long ext = (long long)(nn - 1) * (long long)stride_elems;
    const long long min_off = ext < 0 ? ext : 0;
    const size_t span = (size_t)(ext < 0 ? -ext : ext) + 1;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t o_hdr = 0, o_cta = up(sizeof(PtHeader)), o_lists = o_cta + up((size_t)nctas * 8),
                 o_in = o_lists + up(2 * (size_t)(nn / 2 + 1) * pos_bytes);
    if (arena_reserve(ctx, o_in + (on_device ? 0 : up(span * esize))) != cudaSuccess) return fail(NDS_CUDA_ERROR, "arena");
    char *arena = (char *)ctx->arena;
    char *dptr = (char *)data;
    if (!on_device) {
        const char *src = (const char *)data + min_off * (long long)esize;
        if (stage_h2d(ctx, arena + o_in, src, span * esize) != cudaSuccess)
            return fail(NDS_CUDA_ERROR, "H2D copy");
        dptr = arena + o_in - min_off * (long long)esize;
    }
    PtHeader *hdr = (PtHeader *)(arena + o_hdr);
    unsigned long long *cta_less = (unsigned long long *)(arena + o_cta);
    void *lists = arena + o_lists;
    cudaError_t e;
#define NDS_PT_CASE(T)                                                                                                   \
    e = wide ? partition_typed2<T, unsigned long long>(ctx, (T *)dptr, stride_elems, nn, pivot_index, hdr, cta_less, nctas, \
                                                        chunk, lists)                                                      \
             : partition_typed2<T, unsigned>(ctx, (T *)dptr, stride_elems, nn, pivot_index, hdr, cta_less, nctas, chunk,   \
                                             lists)
    switch (dtype) {
    case NDS_I8: NDS_PT_CASE(int8_t); break;
    case NDS_I16: NDS_PT_CASE(int16_t); break;
    case NDS_I32: NDS_PT_CASE(int32_t); break;
    case NDS_I64: NDS_PT_CASE(int64_t); break;
    case NDS_U8: NDS_PT_CASE(uint8_t); break;
    case NDS_U16: NDS_PT_CASE(uint16_t); break;
    case NDS_U32: NDS_PT_CASE(uint32_t); break;
    case NDS_U64: NDS_PT_CASE(uint64_t); break;
    case NDS_F32: NDS_PT_CASE(float); break;
    default: NDS_PT_CASE(double); break;
    }
#undef NDS_PT_CASE
    ctx->last_path = "partition";
    if (e != cudaSuccess) return fail(NDS_CUDA_ERROR, cudaGetErrorString(e));
    PtHeader h;
    if (!on_device) {
        char *dst = (char *)data + min_off * (long long)esize;
        if (cudaMemcpyAsync(dst, arena + o_in, span * esize, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
            return fail(NDS_CUDA_ERROR, "D2H copy");
    }
    if (cudaMemcpyAsync(&h, hdr, sizeof h, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return fail(NDS_CUDA_ERROR, "result copy");
    *partition_index = (size_t)h.m;
    return NDS_OK;
}
