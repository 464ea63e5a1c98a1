// Layout stage: lanes with an arbitrary stride (or contiguous lanes that are not 16-byte aligned) are gathered into an
// aligned, padded, lane-contiguous scratch buffer [lane][pitch] that the bit-sliced lane kernels then read with
// 16-byte loads.  Two extra passes over the data (one read, one write, then the kernel's own read) -- against the
// CTA-per-lane radix fallback that reads every lane key_bytes times per rank with a shared atomic per element.
//
//   transpose mode   neighbouring lanes are closer in memory than neighbouring elements of a lane (Axis(0) of a C-order
//                    matrix): 32 x 32 tiles through shared memory, reads coalesced across lanes, writes along lanes
//   direct mode      elements of a lane are the closer ones: rows are copied as they are
//
// The reference has no such stage: its lanes_mut(axis) walks any stride in place (src/quantile/mod.rs:470-472).
#include <algorithm>

#include "nds_internal.h"

namespace nds {

namespace {

template <typename U, bool TRANSPOSE>
__global__ void __launch_bounds__(256)
gather_lanes_kernel(const U *__restrict__ src, LaneGeom g, U *__restrict__ dst, unsigned long long pitch) {
    const unsigned long long n = g.axis_len;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const unsigned long long tiles_r = (n + 31) / 32, tiles_l = (g.nlanes + 31) / 32;
    const unsigned long long n_tiles = tiles_r * tiles_l;
    if (TRANSPOSE) {
        __shared__ U tile[32][33];
        for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const unsigned long long tl = t / tiles_r, tr = t - tl * tiles_r;
            const unsigned long long l_in = tl * 32 + tx;       // this thread's lane while reading
            long long in_off = 0, oo;
            if (l_in < g.nlanes) lane_offsets(g, l_in, in_off, oo);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned long long r = tr * 32 + ty + 8 * j;
                if (l_in < g.nlanes && r < n) tile[ty + 8 * j][tx] = src[in_off + (long long)r * g.axis_stride];
            }
            __syncthreads();
            const unsigned long long r_out = tr * 32 + tx;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned long long l_out = tl * 32 + ty + 8 * j;
                if (l_out < g.nlanes && r_out < n) dst[l_out * pitch + r_out] = tile[tx][ty + 8 * j];
            }
            __syncthreads();
        }
    } else {
        // one warp per (lane, run of 1024 elements)
        const unsigned long long runs = (n + 1023) / 1024;
        const unsigned long long n_items = g.nlanes * runs;
        for (unsigned long long it = (unsigned long long)blockIdx.x * 8 + ty; it < n_items; it += (unsigned long long)gridDim.x * 8) {
            const unsigned long long l = it / runs, run = it - l * runs;
            long long in_off, oo;
            lane_offsets(g, l, in_off, oo);
#pragma unroll 4
            for (int j = 0; j < 32; ++j) {
                const unsigned long long r = run * 1024 + (unsigned long long)j * 32 + tx;
                if (r < n) dst[l * pitch + r] = src[in_off + (long long)r * g.axis_stride];
            }
        }
    }
}

}  // namespace

// lanes of `c` (4- or 8-byte elements) -> dst[lane * pitch + r], lanes numbered in C order over the outer dims
cudaError_t launch_gather_lanes(nds_ctx *ctx, const DeviceCall &c, size_t esize, void *dst, unsigned long long pitch) {
    const LaneGeom &g = c.geom;
    long long lane_step = 0;   // smallest stride between neighbouring lanes
    for (int d = 0; d < g.n_outer; ++d)
        if (g.oshape[d] > 1) {
            const long long s = g.in_ostride[d] < 0 ? -g.in_ostride[d] : g.in_ostride[d];
            if (lane_step == 0 || s < lane_step) lane_step = s;
        }
    const long long as = g.axis_stride < 0 ? -g.axis_stride : g.axis_stride;
    const bool transpose = lane_step != 0 && lane_step < as;
    const unsigned long long tiles = ((g.axis_len + 31) / 32) * ((g.nlanes + 31) / 32);
    const unsigned long long items = g.nlanes * ((g.axis_len + 1023) / 1024);
    const unsigned long long want = transpose ? tiles : (items + 7) / 8;
    const unsigned grid = (unsigned)std::min<unsigned long long>(std::max<unsigned long long>(want, 1), (unsigned long long)ctx->sm_count * 16);
    if (esize == 4) {
        if (transpose) gather_lanes_kernel<uint32_t, true><<<grid, 256, 0, ctx->stream>>>((const uint32_t *)c.data, g, (uint32_t *)dst, pitch);
        else gather_lanes_kernel<uint32_t, false><<<grid, 256, 0, ctx->stream>>>((const uint32_t *)c.data, g, (uint32_t *)dst, pitch);
    } else {
        if (transpose) gather_lanes_kernel<uint64_t, true><<<grid, 256, 0, ctx->stream>>>((const uint64_t *)c.data, g, (uint64_t *)dst, pitch);
        else gather_lanes_kernel<uint64_t, false><<<grid, 256, 0, ctx->stream>>>((const uint64_t *)c.data, g, (uint64_t *)dst, pitch);
    }
    ctx->launches += 1;
    return cudaGetLastError();
}

}  // namespace nds
