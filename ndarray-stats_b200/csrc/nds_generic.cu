// General selection paths: correct for every dtype / stride / lane length / rank count.
//
//   generic_lanes_kernel  one CTA per lane, MSB radix select (8-bit digits) over keys staged in
//                         shared memory (or streamed from global when the lane does not fit).
//   long_*_kernel         one very long lane spread over the whole grid: per pass every CTA builds
//                         privatised shared-memory digit histograms for all active key prefixes, the
//                         histograms are merged in global memory (and across GPUs with ncclAllReduce
//                         when the lane is sharded), and a one-CTA kernel picks the digit of every
//                         requested rank on the device — no host round trip between passes.
//
// These restate nothing from the reference's quickselect (src/sort.rs:133-283): order statistics are
// determined by the multiset, so any exact selection is bit-identical (SURVEY.md §8a notes).
#include <algorithm>

#include "nds_internal.h"

namespace nds {

constexpr int GEN_THREADS = 256;

__device__ inline unsigned warp_incl_scan_u32(unsigned v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += t;
    }
    return v;
}
__device__ inline unsigned long long warp_incl_scan_u64(unsigned long long v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += t;
    }
    return v;
}

// ------------------------------------------------------------------------------------------------
// CTA per lane
// ------------------------------------------------------------------------------------------------
template <typename T, bool MASKED>
__global__ void __launch_bounds__(GEN_THREADS)
generic_lanes_kernel(const T *__restrict__ data, const uint8_t *__restrict__ valid, LaneGeom g, QuantArgs qa,
                     T *__restrict__ out, uint8_t *__restrict__ out_valid, unsigned cap_elems) {
    using Tr = Traits<T>;
    using Key = typename Tr::Key;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Key *skeys = reinterpret_cast<Key *>(smem_raw);
    __shared__ unsigned hist[256];
    __shared__ unsigned s_digit, s_cum, s_invalid;
    const Key KEY_MAX = (Key) ~(Key)0;
    const int tid = threadIdx.x;
    const unsigned long long n = g.axis_len;
    const long long stride = g.axis_stride;
    const bool resident = n <= (unsigned long long)cap_elems;

    for (unsigned long long lane = blockIdx.x; lane < g.nlanes; lane += gridDim.x) {
        long long in_off, out_off;
        lane_offsets(g, lane, in_off, out_off);
        const T *lp = data + in_off;
        const uint8_t *vp = MASKED ? valid + in_off : nullptr;

        // ---- stage keys, count invalid (NaN / None) elements --------------------------------------
        if (tid == 0) s_invalid = 0;
        __syncthreads();
        unsigned my_invalid = 0;
        for (unsigned long long i = tid; i < n; i += GEN_THREADS) {
            T x = lp[(long long)i * stride];
            bool inv = MASKED ? (vp[(long long)i * stride] == 0) : Tr::is_nan(x);
            my_invalid += inv ? 1u : 0u;
            if (resident) skeys[i] = inv ? KEY_MAX : Tr::to_key(x);
        }
        my_invalid = __reduce_add_sync(0xffffffffu, my_invalid);
        if ((tid & 31) == 0 && my_invalid) atomicAdd(&s_invalid, my_invalid);
        __syncthreads();
        const unsigned invalid = s_invalid;
        const bool drop = MASKED || qa.skipnan;
        if (!drop && Tr::is_float && invalid && tid == 0) atomicOr(qa.status, STATUS_BIT_NAN);
        const unsigned long long n_eff = drop ? n - invalid : n;

        if (n_eff == 0) {  // lane without a single value: NaN / None (src/quantile/mod.rs:532-534)
            for (int t = tid; t < qa.nq; t += GEN_THREADS) {
                out[out_off + (long long)t * g.out_axis_stride] = canonical_nan<T>();
                if (out_valid) out_valid[out_off + (long long)t * g.out_axis_stride] = 0;
            }
            __syncthreads();
            continue;
        }

        // ---- CTA-wide radix select of one rank ----------------------------------------------------
        auto select = [&](unsigned long long r) -> Key {
            Key prefix = 0, maskbits = 0;
            for (int shift = Tr::key_bits - 8; shift >= 0; shift -= 8) {
                hist[tid] = 0;
                __syncthreads();
                for (unsigned long long i = tid; i < n; i += GEN_THREADS) {
                    Key k;
                    if (resident) {
                        k = skeys[i];
                    } else {
                        T x = lp[(long long)i * stride];
                        bool inv = MASKED ? (vp[(long long)i * stride] == 0) : Tr::is_nan(x);
                        k = inv ? KEY_MAX : Tr::to_key(x);
                    }
                    if ((Key)(k & maskbits) == prefix) atomicAdd(&hist[(unsigned)(k >> shift) & 0xFFu], 1u);
                }
                __syncthreads();
                if (tid < 32) {
                    unsigned c[8], sum = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        c[j] = hist[tid * 8 + j];
                        sum += c[j];
                    }
                    unsigned incl = warp_incl_scan_u32(sum);
                    unsigned excl = incl - sum;
                    if (r >= excl && r < incl) {
                        unsigned run = excl;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (r >= run && r < (unsigned long long)run + c[j]) {
                                s_digit = tid * 8 + j;
                                s_cum = run;
                            }
                            run += c[j];
                        }
                    }
                }
                __syncthreads();
                prefix = (Key)(prefix | (Key)((Key)s_digit << shift));
                maskbits = (Key)(maskbits | (Key)((Key)0xFFu << shift));
                r -= s_cum;
            }
            return prefix;
        };

        // ---- every requested quantile of this lane (src/quantile/mod.rs:475-487) ------------------
        unsigned long long cache_rank[2] = {~0ull, ~0ull};
        Key cache_key[2] = {0, 0};
        auto cached_select = [&](unsigned long long r) -> Key {
            if (r == cache_rank[0]) return cache_key[0];
            if (r == cache_rank[1]) return cache_key[1];
            Key k = select(r);
            cache_rank[1] = cache_rank[0];
            cache_key[1] = cache_key[0];
            cache_rank[0] = r;
            cache_key[0] = k;
            return k;
        };
        for (int t = 0; t < qa.nq; ++t) {
            RankMath rm = slot_rank_math(qa, t, n_eff);
            Key kl = 0, kh = 0;
            if (needs_lower(qa.interp, rm.frac)) kl = cached_select(rm.lower);
            if (needs_higher(qa.interp, rm.frac)) kh = cached_select(rm.higher);
            if (tid == 0) {
                T r = interpolate<T>(qa.interp, Tr::from_key(kl), Tr::from_key(kh), rm.frac, qa.status);
                out[out_off + (long long)t * g.out_axis_stride] = r;
                if (out_valid) out_valid[out_off + (long long)t * g.out_axis_stride] = 1;
            }
        }
        __syncthreads();
    }
}

template <typename T>
static cudaError_t launch_generic_typed(nds_ctx *ctx, const DeviceCall &c) {
    using Key = typename Traits<T>::Key;
    // Stage lanes in shared memory when they fit: up to ~96 KiB so two CTAs share an SM.
    const size_t cap_bytes = 96 * 1024;
    unsigned cap_elems = (unsigned)(cap_bytes / sizeof(Key));
    size_t smem = 0;
    if (c.geom.axis_len <= cap_elems) {
        smem = (size_t)c.geom.axis_len * sizeof(Key);
        smem = (smem + 15) & ~(size_t)15;
    } else {
        cap_elems = 0;
    }
    unsigned long long want = c.geom.nlanes;
    unsigned grid = (unsigned)std::min<unsigned long long>(want, (unsigned long long)ctx->sm_count * 8);
    if (grid == 0) return cudaSuccess;
    auto launch = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap_bytes);
        if (e != cudaSuccess) return e;
        kernel<<<grid, GEN_THREADS, smem, ctx->stream>>>((const T *)c.data, c.valid, c.geom, c.q, (T *)c.out,
                                                          c.out_valid, cap_elems);
        ctx->launches += 1;
        return cudaGetLastError();
    };
    if (c.valid) return launch(generic_lanes_kernel<T, true>);
    return launch(generic_lanes_kernel<T, false>);
}

#define NDS_DISPATCH_DTYPE(dtype, FN, ...)                               \
    switch (dtype) {                                                     \
    case NDS_I8: return FN<int8_t>(__VA_ARGS__);                         \
    case NDS_I16: return FN<int16_t>(__VA_ARGS__);                       \
    case NDS_I32: return FN<int32_t>(__VA_ARGS__);                       \
    case NDS_I64: return FN<int64_t>(__VA_ARGS__);                       \
    case NDS_U8: return FN<uint8_t>(__VA_ARGS__);                        \
    case NDS_U16: return FN<uint16_t>(__VA_ARGS__);                      \
    case NDS_U32: return FN<uint32_t>(__VA_ARGS__);                      \
    case NDS_U64: return FN<uint64_t>(__VA_ARGS__);                      \
    case NDS_F32: return FN<float>(__VA_ARGS__);                         \
    case NDS_F64: return FN<double>(__VA_ARGS__);                        \
    }                                                                    \
    return cudaErrorInvalidValue;

cudaError_t launch_generic_cta(nds_ctx *ctx, const DeviceCall &c) {
    ctx->last_path = "generic_cta";
    NDS_DISPATCH_DTYPE(c.dtype, launch_generic_typed, ctx, c)
}

// ------------------------------------------------------------------------------------------------
// One long lane over the whole grid
// ------------------------------------------------------------------------------------------------
// Scratch layout (all 8-byte words), R = 2 * nq rank slots:
//   red[0]            element count of this shard (summed over ranks when sharded)
//   red[1]            invalid (NaN / None) count
//   red[2 ...]        hist[R][256]  u64 digit histograms, one row per prefix group
//   --- not reduced ---
//   n_eff, nuniq, ngroups, empty
//   slot_rank[R]      requested rank per (q, lower/higher) slot
//   slot_to_u[R]      slot -> unique-rank index
//   urem[R]           remaining rank of each unique rank inside its prefix
//   uprefix[R]        key prefix decided so far
//   ugroup[R]         prefix group (row of hist) of each unique rank
//   gprefix[R]        ascending distinct prefixes (what the histogram kernel searches)
struct LongView {
    unsigned long long *red;
    unsigned long long *hist;
    unsigned long long *hdr;  // [0]=n_eff [1]=nuniq [2]=ngroups [3]=empty [4]=n_total
    unsigned long long *slot_rank, *slot_to_u, *urem, *uprefix, *ugroup, *gprefix;
    int R;
};
__host__ __device__ inline LongView long_view(void *scratch, int nq) {
    LongView v;
    v.R = 2 * nq;
    unsigned long long *p = (unsigned long long *)scratch;
    v.red = p;
    v.hist = p + 2;
    p += 2 + (size_t)v.R * 256;
    v.hdr = p;
    p += 8;
    v.slot_rank = p; p += v.R;
    v.slot_to_u = p; p += v.R;
    v.urem = p; p += v.R;
    v.uprefix = p; p += v.R;
    v.ugroup = p; p += v.R;
    v.gprefix = p; p += v.R;
    return v;
}
size_t long_scratch_bytes(int nq) {
    size_t R = 2 * (size_t)nq;
    return (2 + R * 256 + 8 + 6 * R) * sizeof(unsigned long long);
}
const void *long_total_count_ptr(const void *scratch, int nq) {
    return long_view(const_cast<void *>(scratch), nq).hdr + 4;
}
static size_t long_reduce_words(int nq) { return 2 + 2 * (size_t)nq * 256; }

constexpr int LONG_THREADS = 512;

template <typename T, bool MASKED>
__global__ void __launch_bounds__(LONG_THREADS)
long_hist_kernel(const T *__restrict__ data, const uint8_t *__restrict__ valid, unsigned long long n,
                 long long stride, void *scratch, int nq, int shift, int first_pass, unsigned gmax) {
    using Tr = Traits<T>;
    using Key = typename Tr::Key;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned *hist_s = reinterpret_cast<unsigned *>(smem_raw);                       // gmax * 256
    Key *gp_s = reinterpret_cast<Key *>(smem_raw + (size_t)gmax * 256 * sizeof(unsigned));  // gmax
    LongView v = long_view(scratch, nq);
    const unsigned ngroups = first_pass ? 1u : (unsigned)v.hdr[2];
    if (!first_pass && v.hdr[3]) return;  // no valid element anywhere
    const Key maskbits = first_pass ? (Key)0 : (Key)(~0ull << (shift + 8));
    const unsigned long long tid0 = (unsigned long long)blockIdx.x * LONG_THREADS + threadIdx.x;
    const unsigned long long nthreads = (unsigned long long)gridDim.x * LONG_THREADS;
    unsigned long long my_invalid = 0;

    for (unsigned g0 = 0; g0 < ngroups; g0 += gmax) {
        const unsigned gcount = min(gmax, ngroups - g0);
        for (unsigned j = threadIdx.x; j < gcount * 256; j += LONG_THREADS) hist_s[j] = 0;
        for (unsigned j = threadIdx.x; j < gcount; j += LONG_THREADS)
            gp_s[j] = first_pass ? (Key)0 : (Key)v.gprefix[g0 + j];
        __syncthreads();
        for (unsigned long long i = tid0; i < n; i += nthreads) {
            T x = data[(long long)i * stride];
            bool inv = MASKED ? (valid[(long long)i * stride] == 0) : Tr::is_nan(x);
            if (inv) {
                if (first_pass) my_invalid += 1;
                continue;
            }
            Key k = Tr::to_key(x);
            unsigned digit = (unsigned)(k >> shift) & 0xFFu;
            if (first_pass) {
                atomicAdd(&hist_s[digit], 1u);
            } else {
                Key pk = (Key)(k & maskbits);
                unsigned lo = 0, hi = gcount;  // lower_bound over ascending group prefixes
                while (lo < hi) {
                    unsigned mid = (lo + hi) >> 1;
                    if (gp_s[mid] < pk) lo = mid + 1;
                    else hi = mid;
                }
                if (lo < gcount && gp_s[lo] == pk) atomicAdd(&hist_s[lo * 256 + digit], 1u);
            }
        }
        __syncthreads();
        for (unsigned j = threadIdx.x; j < gcount * 256; j += LONG_THREADS) {
            unsigned cnt = hist_s[j];
            if (cnt) atomicAdd(&v.hist[(size_t)g0 * 256 + j], (unsigned long long)cnt);
        }
        __syncthreads();
    }
    if (first_pass) {
        // warp-reduce then one atomic per warp
        for (int d = 16; d > 0; d >>= 1) my_invalid += __shfl_down_sync(0xffffffffu, my_invalid, d);
        if ((threadIdx.x & 31) == 0 && my_invalid) atomicAdd(&v.red[1], my_invalid);
        if (tid0 == 0) atomicAdd(&v.red[0], n);
    }
}

// One CTA: after the first histogram set up the rank slots, then for every unique rank pick its
// digit from its group's histogram, then regroup equal prefixes for the next pass.
template <typename T, bool MASKED>
__global__ void __launch_bounds__(256)
long_decide_kernel(void *scratch, QuantArgs qa, int shift, int first_pass) {
    using Tr = Traits<T>;
    LongView v = long_view(scratch, qa.nq);
    const int R = v.R;
    const int tid = threadIdx.x;
    if (first_pass) {
        const unsigned long long n_total = v.red[0], n_invalid = v.red[1];
        const bool drop = MASKED || qa.skipnan;
        const unsigned long long n_eff = drop ? n_total - n_invalid : n_total;
        if (tid == 0) {
            if (!drop && Tr::is_float && n_invalid) atomicOr(qa.status, STATUS_BIT_NAN);
            v.hdr[0] = n_eff;
            v.hdr[3] = (n_eff == 0);
            v.hdr[4] = n_total;
        }
        if (n_eff == 0) {
            if (tid == 0) v.hdr[1] = v.hdr[2] = 0;
            return;
        }
        for (int t = tid; t < qa.nq; t += blockDim.x) {
            RankMath rm = slot_rank_math(qa, t, n_eff);
            bool nl = needs_lower(qa.interp, rm.frac), nh = needs_higher(qa.interp, rm.frac);
            v.slot_rank[2 * t] = nl ? rm.lower : rm.higher;
            v.slot_rank[2 * t + 1] = nh ? rm.higher : rm.lower;
        }
        __syncthreads();
        // unique ascending ranks (R is small: O(R^2) counting); ugroup doubles as the "first
        // occurrence" flag until the groups are initialised below
        for (int i = tid; i < R; i += blockDim.x) {
            unsigned long long ri = v.slot_rank[i];
            bool first_i = true;
            for (int k = 0; k < i; ++k)
                if (v.slot_rank[k] == ri) { first_i = false; break; }
            v.ugroup[i] = first_i ? 1 : 0;
        }
        __syncthreads();
        for (int i = tid; i < R; i += blockDim.x) {
            unsigned long long ri = v.slot_rank[i];
            unsigned u = 0;
            for (int j = 0; j < R; ++j)
                if (v.ugroup[j] && v.slot_rank[j] < ri) ++u;
            v.slot_to_u[i] = u;
            v.urem[u] = ri;  // every slot with this rank writes the same value
            v.uprefix[u] = 0;
        }
        __syncthreads();
        for (int i = tid; i < R; i += blockDim.x) v.ugroup[i] = 0;
        if (tid == 0) {
            unsigned long long mx = 0;
            for (int i = 0; i < R; ++i) mx = max(mx, v.slot_to_u[i]);
            v.hdr[1] = mx + 1;
            v.hdr[2] = 1;
            v.gprefix[0] = 0;
        }
        __syncthreads();
    } else if (v.hdr[3]) {
        return;
    }
    const int nuniq = (int)v.hdr[1];
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    for (int u = warp; u < nuniq; u += nwarps) {
        const unsigned long long *row = v.hist + (size_t)v.ugroup[u] * 256;
        const unsigned long long r = v.urem[u];
        unsigned long long c[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            c[j] = row[lane * 8 + j];
            sum += c[j];
        }
        unsigned long long incl = warp_incl_scan_u64(sum);
        unsigned long long run = incl - sum;
        if (r >= run && r < incl) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (r >= run && r < run + c[j]) {
                    v.urem[u] = r - run;
                    v.uprefix[u] |= (unsigned long long)(lane * 8 + j) << shift;
                }
                run += c[j];
            }
        }
    }
    __syncthreads();
    if (tid == 0) {  // ranks ascend => prefixes are non-decreasing: run-length regroup
        unsigned long long gcount = 0;
        for (int u = 0; u < nuniq; ++u) {
            if (u == 0 || v.uprefix[u] != v.uprefix[u - 1]) {
                v.gprefix[gcount] = v.uprefix[u];
                ++gcount;
            }
            v.ugroup[u] = gcount - 1;
        }
        v.hdr[2] = gcount;
    }
}

template <typename T>
__global__ void long_finalize_kernel(void *scratch, QuantArgs qa, T *out, uint8_t *out_valid,
                                     long long out_axis_stride) {
    using Tr = Traits<T>;
    using Key = typename Tr::Key;
    LongView v = long_view(scratch, qa.nq);
    const unsigned long long n_eff = v.hdr[0];
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < qa.nq; t += gridDim.x * blockDim.x) {
        if (v.hdr[3]) {
            out[(long long)t * out_axis_stride] = canonical_nan<T>();
            if (out_valid) out_valid[(long long)t * out_axis_stride] = 0;
            continue;
        }
        RankMath rm = slot_rank_math(qa, t, n_eff);
        T lower = Tr::from_key((Key)v.uprefix[v.slot_to_u[2 * t]]);
        T higher = Tr::from_key((Key)v.uprefix[v.slot_to_u[2 * t + 1]]);
        out[(long long)t * out_axis_stride] = interpolate<T>(qa.interp, lower, higher, rm.frac, qa.status);
        if (out_valid) out_valid[(long long)t * out_axis_stride] = 1;
    }
}

template <typename T, bool MASKED>
static cudaError_t launch_long_lane(nds_ctx *ctx, const T *lane, const uint8_t *vlane, unsigned long long n,
                                    long long stride, const QuantArgs &qa, T *out, uint8_t *out_valid,
                                    long long out_axis_stride, void *scratch, bool sharded) {
    using Tr = Traits<T>;
    const int nq = qa.nq;
    const int R = 2 * nq;
    // shared memory: one 1 KiB histogram + one key per prefix group handled per sweep
    const size_t per_group = 256 * sizeof(unsigned) + sizeof(typename Tr::Key);
    size_t budget = (size_t)ctx->max_smem_optin > 8192 ? (size_t)ctx->max_smem_optin - 2048 : 46 * 1024;
    unsigned gmax = (unsigned)std::min<size_t>((size_t)R, budget / per_group);
    if (gmax == 0) gmax = 1;
    size_t smem = (size_t)gmax * per_group + 16;
    auto hist_k = long_hist_kernel<T, MASKED>;
    cudaError_t e = cudaFuncSetAttribute(hist_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hist_k, LONG_THREADS, smem);
    if (occ < 1) occ = 1;
    unsigned long long max_blocks = (n + LONG_THREADS - 1) / LONG_THREADS;
    unsigned grid = (unsigned)std::max<unsigned long long>(
        1, std::min<unsigned long long>(max_blocks, (unsigned long long)ctx->sm_count * occ));
    const size_t red_bytes = long_reduce_words(nq) * sizeof(unsigned long long);

    bool first = true;
    for (int shift = Tr::key_bits - 8; shift >= 0; shift -= 8) {
        e = cudaMemsetAsync(scratch, 0, red_bytes, ctx->stream);
        if (e != cudaSuccess) return e;
        hist_k<<<grid, LONG_THREADS, smem, ctx->stream>>>(lane, vlane, n, stride, scratch, nq, shift, first ? 1 : 0,
                                                          gmax);
        ctx->launches += 1;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (sharded && ctx->nccl_comm) {
            // first pass: counts + one histogram row; later passes: only the histogram rows (element
            // counts are not needed again, but keeping one buffer keeps this to ONE collective per pass)
            size_t words = first ? 2 + 256 : long_reduce_words(nq);
            std::string err;
            if (nccl_allreduce_sum_u64(ctx, scratch, words, ctx->stream, &err) != 0) {
                ctx->last_error = err;
                return cudaErrorUnknown;
            }
        }
        long_decide_kernel<T, MASKED><<<1, 256, 0, ctx->stream>>>(scratch, qa, shift, first ? 1 : 0);
        ctx->launches += 1;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        first = false;
    }
    long_finalize_kernel<T><<<(nq + 127) / 128, 128, 0, ctx->stream>>>(scratch, qa, out, out_valid, out_axis_stride);
    ctx->launches += 1;
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_long_typed(nds_ctx *ctx, const DeviceCall &c, void *scratch, bool sharded) {
    for (unsigned long long lane = 0; lane < c.geom.nlanes; ++lane) {
        long long in_off, out_off;
        lane_offsets(c.geom, lane, in_off, out_off);
        cudaError_t e;
        if (c.valid)
            e = launch_long_lane<T, true>(ctx, (const T *)c.data + in_off, c.valid + in_off, c.geom.axis_len,
                                          c.geom.axis_stride, c.q, (T *)c.out + out_off,
                                          c.out_valid ? c.out_valid + out_off : nullptr, c.geom.out_axis_stride,
                                          scratch, sharded);
        else
            e = launch_long_lane<T, false>(ctx, (const T *)c.data + in_off, nullptr, c.geom.axis_len,
                                           c.geom.axis_stride, c.q, (T *)c.out + out_off,
                                           c.out_valid ? c.out_valid + out_off : nullptr, c.geom.out_axis_stride,
                                           scratch, sharded);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_long_radix(nds_ctx *ctx, const DeviceCall &c, void *scratch, bool sharded) {
    ctx->last_path = sharded ? "long_radix_sharded" : "long_radix";
    NDS_DISPATCH_DTYPE(c.dtype, launch_long_typed, ctx, c, scratch, sharded)
}

}  // namespace nds
