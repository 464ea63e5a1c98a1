// HistogramExt::histogram (src/histogram/histograms.rs:137-147): every row of an (observations x dimensions) matrix is
// looked up in a grid of per-dimension bins (Bins::index_of -> Edges::indices_of, src/histogram/bins.rs:217-229:
// left-closed, right-open intervals between consecutive sorted edges; a point outside any projection is skipped) and
// the cell's count is bumped.  SURVEY.md section 8(f), row N3: the consumer of the bin-building strategies.
//
// One thread per observation: a binary search per dimension over that dimension's edges (staged in shared memory when
// they fit), then one atomic add -- into a per-CTA shared-memory copy of the counts when the grid is small (flushed once
// per CTA), straight into global memory otherwise.
#include <algorithm>
#include <cstring>

#include "nds_internal.h"

namespace nds {

namespace {

constexpr int HIST_THREADS = 256;
constexpr int HIST_SMEM_EDGES = 4096;     // edges (all dimensions together) staged in shared memory
constexpr int HIST_SMEM_CELLS = 8192;     // grid cells counted in shared memory

struct HistGeom {
    unsigned long long n_obs;
    int n_dims;
    long long stride_obs, stride_dim;                 // element strides of the observation matrix
    unsigned long long edge_off[NDS_MAX_DIMS];        // first edge of dimension d in the concatenated edge array
    unsigned long long n_edges[NDS_MAX_DIMS];
    unsigned long long cell_stride[NDS_MAX_DIMS];     // C-order strides of the count array
    unsigned long long n_cells, n_edges_total;
};

template <typename T>
__global__ void __launch_bounds__(HIST_THREADS)
histogram_kernel(const T *__restrict__ data, const T *__restrict__ edges, HistGeom g, unsigned long long *__restrict__ counts,
                 int edges_in_smem, int cells_in_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sedges = reinterpret_cast<T *>(smem_raw);
    unsigned *scells = reinterpret_cast<unsigned *>(smem_raw + (edges_in_smem ? ((size_t)g.n_edges_total * sizeof(T) + 15) / 16 * 16 : 0));
    if (edges_in_smem)
        for (unsigned long long i = threadIdx.x; i < g.n_edges_total; i += HIST_THREADS) sedges[i] = edges[i];
    if (cells_in_smem)
        for (unsigned long long i = threadIdx.x; i < g.n_cells; i += HIST_THREADS) scells[i] = 0u;
    __syncthreads();
    const T *E = edges_in_smem ? sedges : edges;
    const unsigned long long stride = (unsigned long long)gridDim.x * HIST_THREADS;
    for (unsigned long long i = (unsigned long long)blockIdx.x * HIST_THREADS + threadIdx.x; i < g.n_obs; i += stride) {
        unsigned long long cell = 0;
        bool inside = true;
        for (int d = 0; d < g.n_dims && inside; ++d) {
            const T v = data[(long long)i * g.stride_obs + (long long)d * g.stride_dim];
            const T *e = E + g.edge_off[d];
            const unsigned long long ne = g.n_edges[d];
            // number of edges <= v (upper bound); NaN compares false everywhere: 0 edges, outside
            unsigned long long lo = 0, hi = ne;
            while (lo < hi) {
                const unsigned long long mid = (lo + hi) >> 1;
                if (e[mid] <= v) lo = mid + 1;
                else hi = mid;
            }
            // bins.rs:220-228: v == last edge or beyond -> None; below the first edge -> None
            if (lo == 0 || lo >= ne || !(v >= e[0])) inside = false;
            else cell += (lo - 1) * g.cell_stride[d];
        }
        if (!inside) continue;
        if (cells_in_smem) atomicAdd(&scells[cell], 1u);
        else atomicAdd(&counts[cell], 1ull);
    }
    if (cells_in_smem) {
        __syncthreads();
        for (unsigned long long i = threadIdx.x; i < g.n_cells; i += HIST_THREADS) {
            const unsigned c = scells[i];
            if (c) atomicAdd(&counts[i], (unsigned long long)c);
        }
    }
}

template <typename T>
cudaError_t launch_hist(nds_ctx *ctx, const void *data, const void *edges, const HistGeom &g, unsigned long long *counts) {
    const int edges_in_smem = g.n_edges_total <= (unsigned long long)HIST_SMEM_EDGES ? 1 : 0;
    const int cells_in_smem = g.n_cells <= (unsigned long long)HIST_SMEM_CELLS ? 1 : 0;
    const size_t smem = (edges_in_smem ? ((size_t)g.n_edges_total * sizeof(T) + 15) / 16 * 16 : 0) + (cells_in_smem ? (size_t)g.n_cells * 4 : 0);
    const unsigned long long want = (g.n_obs + HIST_THREADS - 1) / HIST_THREADS;
    const unsigned grid = (unsigned)std::min<unsigned long long>(std::max<unsigned long long>(want, 1), (unsigned long long)ctx->sm_count * 8);
    auto kernel = histogram_kernel<T>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kernel<<<grid, HIST_THREADS, smem, ctx->stream>>>((const T *)data, (const T *)edges, g, counts, edges_in_smem, cells_in_smem);
    ctx->launches += 1;
    return cudaGetLastError();
}

}  // namespace
}  // namespace nds

using namespace nds;

extern "C" nds_status nds_histogram(nds_ctx *ctx, nds_dtype dtype, const void *data, size_t n_obs, size_t n_dims,
                                    ptrdiff_t stride_obs, ptrdiff_t stride_dim, const void *edges, const size_t *n_edges,
                                    unsigned long long *counts) {
    if (!ctx) return NDS_BAD_ARGUMENT;
    auto fail = [&](nds_status st, const char *msg) {
        ctx->last_error = msg;
        return st;
    };
    if (dtype < NDS_I8 || dtype > NDS_F64) return fail(NDS_UNSUPPORTED, "unknown dtype");
    if (n_dims < 1 || n_dims > (size_t)NDS_MAX_DIMS || !n_edges || !counts) return fail(NDS_BAD_ARGUMENT, "bad grid");
    const size_t esize = nds_dtype_size(dtype);
    HistGeom g;
    std::memset(&g, 0, sizeof g);
    g.n_obs = n_obs;
    g.n_dims = (int)n_dims;
    g.stride_obs = stride_obs;
    g.stride_dim = stride_dim;
    unsigned long long off = 0, cells = 1;
    for (size_t d = 0; d < n_dims; ++d) {
        g.edge_off[d] = off;
        g.n_edges[d] = n_edges[d];
        off += n_edges[d];
        cells *= n_edges[d] >= 2 ? n_edges[d] - 1 : 0;     // Bins::len (bins.rs:291-299)
    }
    unsigned long long cs = 1;
    for (int d = (int)n_dims - 1; d >= 0; --d) {
        g.cell_stride[d] = cs;
        cs *= n_edges[d] >= 2 ? n_edges[d] - 1 : 0;
    }
    g.n_cells = cells;
    g.n_edges_total = off;
    if (cells == 0) return NDS_OK;                           // an empty grid has no counts to fill
    if (off && !edges) return fail(NDS_BAD_ARGUMENT, "null edges");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return NDS_CUDA_ERROR;
    auto on_dev = [](const void *p) {
        cudaPointerAttributes attr;
        const bool dev = cudaPointerGetAttributes(&attr, p) == cudaSuccess &&
                         (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
        cudaGetLastError();
        return dev;
    };
    const bool data_dev = n_obs ? on_dev(data) : true, counts_dev = on_dev(counts);
    // span of the observation matrix
    long long min_off = 0, max_off = 0;
    if (n_obs) {
        const long long e0 = (long long)(n_obs - 1) * stride_obs, e1 = (long long)(n_dims - 1) * stride_dim;
        min_off = std::min(0ll, e0) + std::min(0ll, e1);
        max_off = std::max(0ll, e0) + std::max(0ll, e1);
    }
    const size_t span = n_obs ? (size_t)(max_off - min_off + 1) : 0;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t o_edges = 0, o_counts = up(off * esize), o_in = o_counts + up(cells * 8);
    if (arena_reserve(ctx, o_in + (data_dev ? 0 : up(span * esize))) != cudaSuccess) return fail(NDS_CUDA_ERROR, "arena");
    char *arena = (char *)ctx->arena;
    if (cudaMemcpyAsync(arena + o_edges, edges, off * esize, cudaMemcpyDefault, ctx->stream) != cudaSuccess)
        return fail(NDS_CUDA_ERROR, "edges copy");
    unsigned long long *dcounts = counts_dev ? counts : (unsigned long long *)(arena + o_counts);
    if (cudaMemsetAsync(dcounts, 0, cells * 8, ctx->stream) != cudaSuccess) return fail(NDS_CUDA_ERROR, "memset");
    const void *dptr = data;
    if (!data_dev && n_obs) {
        const char *src = (const char *)data + min_off * (long long)esize;
        if (stage_h2d(ctx, arena + o_in, src, span * esize) != cudaSuccess)
            return fail(NDS_CUDA_ERROR, "H2D copy");
        dptr = arena + o_in - min_off * (long long)esize;
    }
    cudaError_t e = cudaSuccess;
    if (n_obs) {
        switch (dtype) {
        case NDS_I8: e = launch_hist<int8_t>(ctx, dptr, arena + o_edges, g, dcounts); break;
        case NDS_I16: e = launch_hist<int16_t>(ctx, dptr, arena + o_edges, g, dcounts); break;
        case NDS_I32: e = launch_hist<int32_t>(ctx, dptr, arena + o_edges, g, dcounts); break;
        case NDS_I64: e = launch_hist<int64_t>(ctx, dptr, arena + o_edges, g, dcounts); break;
        case NDS_U8: e = launch_hist<uint8_t>(ctx, dptr, arena + o_edges, g, dcounts); break;
        case NDS_U16: e = launch_hist<uint16_t>(ctx, dptr, arena + o_edges, g, dcounts); break;
        case NDS_U32: e = launch_hist<uint32_t>(ctx, dptr, arena + o_edges, g, dcounts); break;
        case NDS_U64: e = launch_hist<uint64_t>(ctx, dptr, arena + o_edges, g, dcounts); break;
        case NDS_F32: e = launch_hist<float>(ctx, dptr, arena + o_edges, g, dcounts); break;
        default: e = launch_hist<double>(ctx, dptr, arena + o_edges, g, dcounts); break;
        }
    }
    ctx->last_path = "histogram";
    if (e != cudaSuccess) return fail(NDS_CUDA_ERROR, cudaGetErrorString(e));
    if (!counts_dev && cudaMemcpyAsync(counts, dcounts, cells * 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
        return fail(NDS_CUDA_ERROR, "D2H copy");
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(NDS_CUDA_ERROR, "synchronize");
    return NDS_OK;
}
