// Contiguous short lanes (up to 8192 elements of a 4- or 8-byte type): warp-per-lane, bit-sliced
// MSB-first selection, lanes staged through shared memory by 1-D TMA bulk copies.
//
// Why not the shared-memory digit histogram: a shared atomic costs ~2 cycles per element per SM
// (B300_MICROARCH.md "ATOMS"), i.e. <= 0.5 element/cycle/SM, while streaming f32 lanes at the HBM
// roofline needs ~5.8 elements/cycle/SM.  Here every warp owns one lane:
//   1. the lane arrives in the warp's shared-memory buffer by cp.async.bulk (TMA) + mbarrier, the
//      next lane's copy is issued as soon as this one has been consumed, so HBM always has
//      (warps per SM) x (lane bytes) in flight;
//   2. each thread transposes 32 elements x 16 key bits into 16 "plane" words (bit b of 32 elements)
//      with byte-permute / shift-mask butterflies (~4 instructions per element);
//   3. selection walks the key bits MSB-first: count = popc(candidates & plane) summed over the warp
//      with one REDUX, keep the side that contains the rank.  One step costs ~20 instructions for
//      4096 elements instead of 2 per element;
//   4. as soon as <= 32 candidates remain they are gathered (from L2) and ranked directly.
// Key bits beyond the first 16 are handled by further rounds that re-read the lane (L2-resident).
// No key transform is applied: the walk interprets IEEE sign / two's-complement sign on the fly.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "nds_bitslice.cuh"
#include "nds_internal.h"

namespace nds {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "NDS_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra NDS_DONE;\n"
        "bra NDS_WAIT;\n"
        "NDS_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk tensor copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Per-warp shared memory:
//   [ring: S stages x (1024 elements)] [planes: 16 x NB x 32 words] [vm: NB x 32] [im: NB x 32]
// followed (after all warps) by the mbarriers, the candidate index lists and the candidate keys.
constexpr int SHORT_MAX_STAGES = 4;
template <int NB> __host__ __device__ constexpr unsigned short_warp_fixed_bytes() { return 16u * NB * 128u + 2u * NB * 128u; }
constexpr unsigned SHORT_WARP_TAIL_BYTES = SHORT_MAX_STAGES * 8 + 128 + 256;

template <typename T, int NB>
__global__ void __launch_bounds__(NB == 8 ? 256 : 512, 1)
short_bitslice_kernel(const T *__restrict__ data, LaneGeom g, QuantArgs qa, T *__restrict__ out, int S) {
    using Tr = Traits<T>;
    using Key = typename Tr::Key;   // order-preserving key (only used to rank the final <= 32 candidates)
    using Raw = typename RawBits<T>::U;
    constexpr int ES = (int)sizeof(T);
    constexpr int VEC = 16 / ES;          // elements per 16-byte chunk
    constexpr int LD_PER_BLK = 32 / VEC;  // 16-byte loads per thread per 32-element block
    constexpr int ROUNDS = ES * 8 / 16;   // 16 key bits per round
    constexpr int KEYBITS = ES * 8;
    constexpr unsigned BLK_BYTES = 1024u * ES;  // one block = 32 elements per thread = 1024 elements per warp
    constexpr bool IS_FLOAT = Tr::is_float;
    constexpr bool TOP_INVERTED = IS_FLOAT || std::is_signed<T>::value;  // sign bit set => smaller
    constexpr int EXP_LO = (ES == 4) ? 7 : 4;  // exponent bits inside the top 16-bit field: [EXP_LO, 14]
    constexpr unsigned FULL = 0xffffffffu;
    // SMART (8-byte types, 4-byte integers): the 16-bit window of a round is not a fixed slice of the key.  Bits on which
    // the whole lane agrees -- and, for signed integers, bits that merely repeat the sign -- tell no two elements apart, so
    // round 0 takes the sign bit plus the 15 bits below the highest INFORMATIVE bit of the lane, and later rounds start at
    // the next informative bit.  Small integers in a wide type, doubles of one binade, tie-heavy lanes: one round instead
    // of up to four.  (f32 keeps the fixed top half: sign, exponent and 7 mantissa bits nearly always do.)
    constexpr bool SMART = (ES == 8) || !IS_FLOAT;
    constexpr bool SIGNED_INT = !IS_FLOAT && std::is_signed<T>::value;
    constexpr Raw SIGN = (Raw)1 << (KEYBITS - 1);

    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    const size_t per_warp = (size_t)S * BLK_BYTES + short_warp_fixed_bytes<NB>();
    unsigned char *ring = smem + (size_t)warp * per_warp;
    uint32_t *planes = reinterpret_cast<uint32_t *>(ring + (size_t)S * BLK_BYTES) + lane;  // [b][blk][lane]
    uint32_t *vm = planes + 16 * NB * 32;                                                  // valid positions
    uint32_t *im = vm + NB * 32;                                                           // valid and not NaN
    unsigned char *tail = smem + (size_t)W * per_warp + (size_t)warp * SHORT_WARP_TAIL_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(tail);
    unsigned *clist = reinterpret_cast<unsigned *>(tail + SHORT_MAX_STAGES * 8);
    Key *ckeys = reinterpret_cast<Key *>(tail + SHORT_MAX_STAGES * 8 + 128);

    const unsigned n = (unsigned)g.axis_len;
    const unsigned lane_bytes = n * ES;
    const unsigned n_chunks = lane_bytes / 16;
    const int nbn = (int)((n + 1023u) / 1024u);  // blocks actually present in a lane (<= NB)

    // element index of slot s (0..31) of block blk held by this thread; plane bit p holds slot 2*(p&15)+(p>>4)
    auto elem_index = [&](int blk, int s) -> unsigned {
        if (ES == 4) return (unsigned)(((blk * LD_PER_BLK + (s >> 2)) * 32 + lane) * 4 + (s & 3));
        return (unsigned)(((blk * LD_PER_BLK + (s >> 1)) * 32 + lane) * 2 + (s & 1));
    };
    auto slot_of_pos = [](int p) -> int { return 2 * (p & 15) + (p >> 4); };

#pragma unroll 1
    for (int blk = 0; blk < NB; ++blk) {
        uint32_t v = 0;
#pragma unroll 1
        for (int p = 0; p < 32; ++p)
            if (elem_index(blk, slot_of_pos(p)) < n) v |= 1u << p;
        vm[blk * 32] = v;
    }

    if (lane == 0)
        for (int st = 0; st < S; ++st) mbar_init(bars + st, 1);
    fence_mbar_init();
    __syncwarp();

    const unsigned long long nw = (unsigned long long)gridDim.x * W;
    unsigned long long L = (unsigned long long)blockIdx.x * W + warp;

    // ---- TMA producer side of the per-warp block ring (warp-uniform state, lane 0 issues) ------------
    unsigned long long p_lane = L;  // lane whose block is loaded next
    int p_blk = 0, p_stage = 0;
    long long p_in_off = 0;
    auto produce = [&]() {
        if (p_lane >= g.nlanes) return;
        if (p_blk == 0) {
            long long oo;
            lane_offsets(g, p_lane, p_in_off, oo);
        }
        const unsigned off = (unsigned)p_blk * BLK_BYTES;
        const unsigned bytes = min(BLK_BYTES, lane_bytes - off);
        if (lane == 0) {
            fence_proxy_async();  // earlier generic-proxy reads of this stage precede the async-proxy write
            mbar_expect_tx(bars + p_stage, bytes);
            tma_load_1d(ring + (size_t)p_stage * BLK_BYTES,
                        reinterpret_cast<const unsigned char *>(data + p_in_off) + off, bytes, bars + p_stage);
        }
        if (++p_blk == nbn) {
            p_blk = 0;
            p_lane += nw;
        }
        p_stage = (p_stage + 1 == S) ? 0 : p_stage + 1;
    };
    for (int st = 0; st < S; ++st) produce();

    int c_stage = 0;          // consumer side
    unsigned c_parity = 0;    // bit s: parity to wait for on stage s

    // Transpose the 32 16-bit fields in w[] and store the 16 plane words of block blk.
    auto finish_block = [&](uint32_t (&w)[16], int blk, bool want_special) -> uint32_t {
        transpose16(w);
        uint32_t e = 0;
        if (IS_FLOAT && want_special) {
            e = vm[blk * 32];
#pragma unroll
            for (int b = EXP_LO; b <= 14; ++b) e &= w[b];
        }
#pragma unroll
        for (int b = 0; b < 16; ++b) planes[(b * NB + blk) * 32] = w[b];
        return e;
    };
    auto pack = [&](const uint4 &q, uint32_t (&w)[16], int j, int round) {
        if constexpr (ES == 4) {
            const unsigned sel = (round == 0) ? 0x7632u : 0x5410u;
            w[2 * j] = __byte_perm(q.x, q.y, sel);
            w[2 * j + 1] = __byte_perm(q.z, q.w, sel);
        } else {
            const unsigned sel = (round & 1) ? 0x5410u : 0x7632u;
            const uint32_t w0 = (round < 2) ? q.y : q.x, w1 = (round < 2) ? q.w : q.z;
            w[j] = __byte_perm(w0, w1, sel);
        }
    };

    // ---- SMART windows ------------------------------------------------------------------------------------
    // field of one element for the window that starts at bit `sh`: round 0 (`first`) = sign bit on top of 15 window bits,
    // later rounds = 16 window bits (the upper half of the returned word is garbage, the byte permute drops it)
    auto field64 = [&](uint32_t lo, uint32_t hi, int sh, bool first) -> uint32_t {
        uint32_t v = sh >= 32 ? (hi >> (sh - 32)) : __funnelshift_r(lo, hi, sh);
        if (first) v = bitsel(v, hi >> 16, 0x7fffu);
        return v;
    };
    auto field32 = [&](uint32_t x, int sh, bool first) -> uint32_t {
        uint32_t v = x >> sh;
        if (first) v = bitsel(v, x >> 16, 0x7fffu);
        return v;
    };
    auto pack_smart = [&](const uint4 &q, uint32_t (&w)[16], int j, int sh, bool first) {
        if constexpr (ES == 4) {
            w[2 * j] = __byte_perm(field32(q.x, sh, first), field32(q.y, sh, first), 0x5410u);
            w[2 * j + 1] = __byte_perm(field32(q.z, sh, first), field32(q.w, sh, first), 0x5410u);
        } else {
            w[j] = __byte_perm(field64(q.x, q.y, sh, first), field64(q.z, q.w, sh, first), 0x5410u);
        }
    };
    // lane statistics gathered while the lane streams through the registers: z = OR of (x ^ first element) -- the bits
    // that differ somewhere in the lane; o = OR of (x ^ sign copies) -- the bits that are not a repetition of the sign
    // (signed integers); am = largest |x| seen, high word (doubles: is an inf / NaN present?)
    uint32_t st_zlo = 0, st_zhi = 0, st_olo = 0, st_ohi = 0, st_x0lo = 0, st_x0hi = 0;
    auto stats = [&](const uint4 &q, uint32_t &am) {
        if constexpr (ES == 4) {
            const uint32_t e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                st_zlo |= e[u] ^ st_x0lo;
                if constexpr (SIGNED_INT) st_olo |= e[u] ^ (uint32_t)((int32_t)e[u] >> 31);
            }
        } else {
            st_zlo |= q.x ^ st_x0lo;
            st_zhi |= q.y ^ st_x0hi;
            st_zlo |= q.z ^ st_x0lo;
            st_zhi |= q.w ^ st_x0hi;
            if constexpr (SIGNED_INT) {
                const uint32_t s0 = (uint32_t)((int32_t)q.y >> 31), s1 = (uint32_t)((int32_t)q.w >> 31);
                st_olo |= q.x ^ s0;
                st_ohi |= q.y ^ s0;
                st_olo |= q.z ^ s1;
                st_ohi |= q.w ^ s1;
            }
            if constexpr (IS_FLOAT) am = max(am, max(q.y & 0x7fffffffu, q.w & 0x7fffffffu));
        }
    };
    // informative bits below the sign, over the whole warp
    auto informative = [&]() -> Raw {
        uint32_t lo = __reduce_or_sync(FULL, st_zlo), hi = 0;
        if constexpr (SIGNED_INT) lo &= __reduce_or_sync(FULL, st_olo);
        if constexpr (ES == 8) {
            hi = __reduce_or_sync(FULL, st_zhi);
            if constexpr (SIGNED_INT) hi &= __reduce_or_sync(FULL, st_ohi);
            return (((Raw)hi << 32) | (Raw)lo) & ~SIGN;
        } else {
            return (Raw)lo & ~SIGN;
        }
    };
    auto msb = [](Raw v) -> int { return ES == 8 ? 63 - __clzll((long long)v) : 31 - __clz((int)v); };
    auto low_mask = [](int pos) -> Raw { return pos >= KEYBITS ? ~(Raw)0 : (((Raw)1 << pos) - (Raw)1); };
    // shift of the round-0 window whose top informative bit is expected at msb(V) (+ margin when V covers only the
    // first block of the lane and the rest may reach a little higher)
    auto pick_shift = [&](Raw V, int margin) -> int {
        if (V == 0) return KEYBITS - 16;
        const int top = min(msb(V) + margin, KEYBITS - 2);
        return max(top - 14, 0);
    };
    Raw V_lane = 0;   // informative bits of the current lane (sign excluded)
    int sh0 = KEYBITS - 16;  // shift of the round-0 window the plane area holds (while !planes_dirty)
    unsigned smart_nan = 0;  // this thread's NaN count of the current lane (doubles)

    // Any round, from the lane in global memory (L2-resident), window at `sh`.
    auto build_from_global_smart = [&](const unsigned char *gsrc, int sh, bool first) {
#pragma unroll 1
        for (int blk = 0; blk < nbn; ++blk) {
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < LD_PER_BLK; ++j) {
                const unsigned chunk = min((unsigned)((blk * LD_PER_BLK + j) * 32 + lane), n_chunks - 1);
                const uint4 q = __ldg(reinterpret_cast<const uint4 *>(gsrc + (size_t)chunk * 16));
                pack_smart(q, w, j, sh, first);
            }
            transpose16(w);
#pragma unroll
            for (int b = 0; b < 16; ++b) planes[(b * NB + blk) * 32] = w[b];
        }
    };
    // Round 0 of a fresh lane from the ring.  The first block (already in registers) proposes the window, the whole lane
    // confirms it; a lane whose later blocks reach above the proposal is built again from L2 with the right window.
    auto build_from_ring_smart = [&](const unsigned char *gsrc) {
        st_zlo = st_zhi = st_olo = st_ohi = 0;
        smart_nan = 0;
#pragma unroll 1
        for (int blk = 0; blk < nbn; ++blk) {
            const unsigned nchunks_blk = min(BLK_BYTES, lane_bytes - (unsigned)blk * BLK_BYTES) / 16;
            const unsigned char *src = ring + (size_t)c_stage * BLK_BYTES;
            mbar_wait(bars + c_stage, (c_parity >> c_stage) & 1u);
            c_parity ^= 1u << c_stage;
            c_stage = (c_stage + 1 == S) ? 0 : c_stage + 1;
            uint4 q[LD_PER_BLK];
            if (nchunks_blk == BLK_BYTES / 16) {
#pragma unroll
                for (int j = 0; j < LD_PER_BLK; ++j)
                    q[j] = *reinterpret_cast<const uint4 *>(src + (size_t)(j * 32 + lane) * 16);
            } else {
#pragma unroll
                for (int j = 0; j < LD_PER_BLK; ++j) {
                    const unsigned chunk = min((unsigned)(j * 32 + lane), nchunks_blk - 1);
                    q[j] = *reinterpret_cast<const uint4 *>(src + (size_t)chunk * 16);
                }
            }
            __syncwarp();
            produce();
            if (blk == 0) {
                st_x0lo = __shfl_sync(FULL, q[0].x, 0);
                if constexpr (ES == 8) st_x0hi = __shfl_sync(FULL, q[0].y, 0);
            }
            uint32_t am = 0;
#pragma unroll
            for (int j = 0; j < LD_PER_BLK; ++j) stats(q[j], am);
            if (blk == 0) sh0 = pick_shift(informative(), nbn > 1 ? 2 : 0);
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < LD_PER_BLK; ++j) pack_smart(q[j], w, j, sh0, true);
            if constexpr (IS_FLOAT) {  // (doubles) NaN positions of the block, straight from the registers
                const uint32_t valid = vm[blk * 32];
                uint32_t nanm = 0;
                if (__any_sync(FULL, am >= 0x7ff00000u)) {
#pragma unroll
                    for (int j = 0; j < LD_PER_BLK; ++j) {
                        const double d0 = __hiloint2double((int)q[j].y, (int)q[j].x);
                        const double d1 = __hiloint2double((int)q[j].w, (int)q[j].z);
                        nanm |= (d0 != d0 ? 1u : 0u) << j;
                        nanm |= (d1 != d1 ? 1u : 0u) << (j + 16);
                    }
                }
                smart_nan += __popc(valid & nanm);
                im[blk * 32] = valid & ~nanm;
            }
            transpose16(w);
#pragma unroll
            for (int b = 0; b < 16; ++b) planes[(b * NB + blk) * 32] = w[b];
        }
        V_lane = informative();
        if ((V_lane >> (sh0 + 15)) != 0) {  // (rare) informative bits above the proposed window
            sh0 = pick_shift(V_lane, 0);
            __syncwarp();
            build_from_global_smart(gsrc, sh0, true);
        }
    };

    // Round 0 of a fresh lane: consume its blocks from the ring (refilling each stage as soon as it has
    // been read).  Returns (floats) the OR of the "exponent all ones" masks: is an inf/NaN present?
    auto build_from_ring = [&]() -> uint32_t {
        uint32_t special_any = 0;
#pragma unroll 1
        for (int blk = 0; blk < nbn; ++blk) {
            const unsigned nchunks_blk = min(BLK_BYTES, lane_bytes - (unsigned)blk * BLK_BYTES) / 16;
            const unsigned char *src = ring + (size_t)c_stage * BLK_BYTES;
            mbar_wait(bars + c_stage, (c_parity >> c_stage) & 1u);
            c_parity ^= 1u << c_stage;
            c_stage = (c_stage + 1 == S) ? 0 : c_stage + 1;
            uint4 q[LD_PER_BLK];
            if (nchunks_blk == BLK_BYTES / 16) {  // a full block (warp-uniform): no index clamping
#pragma unroll
                for (int j = 0; j < LD_PER_BLK; ++j)
                    q[j] = *reinterpret_cast<const uint4 *>(src + (size_t)(j * 32 + lane) * 16);
            } else {
#pragma unroll
                for (int j = 0; j < LD_PER_BLK; ++j) {
                    // chunks past the end of the lane are clamped onto the last one: their positions are
                    // masked out by vm, so any in-bounds value will do and no predicate is needed
                    const unsigned chunk = min((unsigned)(j * 32 + lane), nchunks_blk - 1);
                    q[j] = *reinterpret_cast<const uint4 *>(src + (size_t)chunk * 16);
                }
            }
            __syncwarp();
            produce();  // this stage is free again: fetch the block S positions ahead
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < LD_PER_BLK; ++j) pack(q[j], w, j, 0);
            special_any |= finish_block(w, blk, true);
        }
        return special_any;
    };
    // Any round, from the lane in global memory (L2-resident: it was streamed in moments ago).
    auto build_from_global = [&](const unsigned char *gsrc, int round) {
#pragma unroll 1
        for (int blk = 0; blk < nbn; ++blk) {
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < LD_PER_BLK; ++j) {
                const unsigned chunk = min((unsigned)((blk * LD_PER_BLK + j) * 32 + lane), n_chunks - 1);
                const uint4 q = __ldg(reinterpret_cast<const uint4 *>(gsrc + (size_t)chunk * 16));
                pack(q, w, j, round);
            }
            finish_block(w, blk, false);
        }
    };

    for (; L < g.nlanes; L += nw) {
        long long in_off, out_off;
        lane_offsets(g, L, in_off, out_off);
        const T *lp = data + in_off;

        uint32_t special_any = 0;
        if constexpr (SMART) build_from_ring_smart(reinterpret_cast<const unsigned char *>(lp));
        else special_any = build_from_ring();

        // ---- NaN handling (floats): elements with an all-ones exponent are inspected individually ----
        const uint32_t *init = vm;
        unsigned n_eff = n;
        if constexpr (SMART && IS_FLOAT) {  // (doubles) the build already wrote the NaN-free masks
            const unsigned nan_total = __reduce_add_sync(FULL, smart_nan);
            if (nan_total) {
                if (!qa.skipnan && lane == 0) atomicOr(qa.status, STATUS_BIT_NAN);
                n_eff = n - nan_total;
            }
            init = im;
        }
        if (!SMART && IS_FLOAT && __any_sync(FULL, special_any != 0)) {
            unsigned my_nan = 0;
#pragma unroll 1
            for (int blk = 0; blk < nbn; ++blk) {
                uint32_t keep = vm[blk * 32];
                uint32_t e = keep;
#pragma unroll
                for (int b = EXP_LO; b <= 14; ++b) e &= planes[(b * NB + blk) * 32];
                while (e) {
                    int p = __ffs(e) - 1;
                    e &= e - 1;
                    T x = lp[elem_index(blk, slot_of_pos(p))];
                    if (x != x) {
                        keep &= ~(1u << p);
                        my_nan += 1;
                    }
                }
                im[blk * 32] = keep;
            }
            const unsigned nan_total = __reduce_add_sync(FULL, my_nan);
            if (nan_total) {
                if (!qa.skipnan && lane == 0) atomicOr(qa.status, STATUS_BIT_NAN);
                n_eff = n - nan_total;  // non-skipnan: the result is meaningless (status is raised) but in bounds
                init = im;
            }
        }
        __syncwarp();

        if (n_eff == 0) {
            for (int t = lane; t < qa.nq; t += 32) out[out_off + (long long)t * g.out_axis_stride] = canonical_nan<T>();
            continue;
        }

        bool planes_dirty = false;  // the plane area no longer holds round 0

        // ---- select rank r (and r+1 when want_next): returns raw bit patterns --------------------------
        auto select = [&](unsigned r, bool want_next, Raw &raw_r, Raw &raw_next) {
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {  // pass 1 only when r+1 fell outside r's final bucket
                if (planes_dirty) {
                    __syncwarp();
                    build_from_global(reinterpret_cast<const unsigned char *>(lp), 0);
                    __syncwarp();
                    planes_dirty = false;
                }
                uint32_t m[NB];
#pragma unroll
                for (int blk = 0; blk < NB; ++blk) m[blk] = (blk < nbn) ? init[blk * 32] : 0u;
                unsigned total = n_eff, rem = r + (unsigned)pass;
                bool neg = false;
                Raw prefix = 0;
                int bits_done = 0;
                bool small = total <= 32;
                bool tie_done = false;
                Raw tie_raw = 0;
#pragma unroll 1
                for (int round = 0; round < ROUNDS && !small; ++round) {
                    if (round > 0) {
                        // Tie-heavy lanes: more than 32 candidates survive a whole 16-bit round.  Before the lane is read
                        // and transposed again for the next 16 bits: are the survivors all one value?  (They usually
                        // are once the bits that tell the lane's few distinct values apart have been walked.)  Every
                        // thread compares its own candidates with the first one; the warp stops at the first mismatch.
                        Raw mine = 0;
                        bool have = false;
#pragma unroll
                        for (int blk = 0; blk < NB; ++blk)
                            if (!have && m[blk]) {
                                mine = to_raw<T>(lp[elem_index(blk, slot_of_pos(__ffs(m[blk]) - 1))]);
                                have = true;
                            }
                        const unsigned owners = __ballot_sync(FULL, have);
                        const Raw ref = __shfl_sync(FULL, mine, __ffs(owners) - 1);
                        bool same = true, go = true;
#pragma unroll
                        for (int blk = 0; blk < NB; ++blk) {  // (unrolled: m[] must stay in registers)
                            if (go) {
                                uint32_t bits = m[blk];
                                while (bits) {
                                    const int p = __ffs(bits) - 1;
                                    bits &= bits - 1;
                                    same = same && (to_raw<T>(lp[elem_index(blk, slot_of_pos(p))]) == ref);
                                }
                                go = __all_sync(FULL, same);
                            }
                        }
                        if (go) {
                            tie_done = true;
                            tie_raw = ref;
                            break;
                        }
                        __syncwarp();
                        build_from_global(reinterpret_cast<const unsigned char *>(lp), round);
                        __syncwarp();
                        planes_dirty = true;
                    }
                    // Walk the 16 planes MSB-first.  `inv` is all-ones while a set bit means "smaller"
                    // (the sign bit itself; every later bit of a negative float): x ^ inv is then 1 on the
                    // LARGER side, so one popc gives the size of the smaller side as total - ones.
                    const uint32_t *pp = planes + 15 * NB * 32;
                    uint32_t inv = (round == 0) ? (TOP_INVERTED ? 0xffffffffu : 0u) : ((IS_FLOAT && neg) ? 0xffffffffu : 0u);
#pragma unroll 1
                    for (int b = 15; b >= 0; --b, pp -= NB * 32) {
                        uint32_t x[NB];
                        unsigned ones = 0;
#pragma unroll
                        for (int blk = 0; blk < NB; ++blk) {
                            x[blk] = pp[blk * 32];
                            ones += __popc(m[blk] & (x[blk] ^ inv));
                        }
                        ones = __reduce_add_sync(FULL, ones);
                        const unsigned small_cnt = total - ones;
                        const bool take_small = rem < small_cnt;
                        total = take_small ? small_cnt : ones;
                        rem = take_small ? rem : rem - small_cnt;
                        // smaller side = positions where (x ^ inv) == 0, i.e. x == inv; larger side x == ~inv
                        const uint32_t keep = take_small ? inv : ~inv;  // value of x on the kept side
                        prefix = (Raw)(prefix << 1) | (Raw)(keep & 1u);
#pragma unroll
                        for (int blk = 0; blk < NB; ++blk) m[blk] &= ~(x[blk] ^ keep);
                        if (round == 0 && b == 15) {
                            neg = IS_FLOAT && (keep & 1u);
                            inv = neg ? 0xffffffffu : 0u;
                        }
                        bits_done += 1;
                        if (total <= 32) {
                            small = true;
                            break;
                        }
                    }
                }
                // `prefix` holds the decided leading bits right-aligned; move them to the top of the key
                if (bits_done < KEYBITS) prefix <<= (KEYBITS - bits_done);
                Raw found, found_next = 0;
                bool have_next = false;
                if (tie_done) {
                    // every remaining candidate is the same value (found by the check below)
                    found = tie_raw;
                    if (rem + 1 < total) {
                        found_next = tie_raw;
                        have_next = true;
                    }
                } else if (total > 32) {
                    // every key bit decided: all remaining candidates are equal to `prefix`
                    found = prefix;
                    if (rem + 1 < total) {
                        found_next = prefix;
                        have_next = true;
                    }
                } else {
                    // gather the <= 32 candidates and rank them by full key
                    unsigned mycnt = 0;
#pragma unroll
                    for (int blk = 0; blk < NB; ++blk) mycnt += __popc(m[blk]);
                    unsigned incl = mycnt;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        unsigned tt = __shfl_up_sync(FULL, incl, d);
                        if (lane >= d) incl += tt;
                    }
                    unsigned base = incl - mycnt;
#pragma unroll
                    for (int blk = 0; blk < NB; ++blk) {
                        uint32_t bits = m[blk];
                        while (bits) {
                            int p = __ffs(bits) - 1;
                            bits &= bits - 1;
                            clist[base++] = elem_index(blk, slot_of_pos(p));
                        }
                    }
                    __syncwarp();
                    Raw myraw = 0;
                    Key mykey = (Key) ~(Key)0;  // padding lanes sort last and never count as "less"
                    if ((unsigned)lane < total) {
                        T x = lp[clist[lane]];
                        myraw = to_raw<T>(x);
                        mykey = Tr::to_key(x);
                    }
                    ckeys[lane] = mykey;
                    __syncwarp();
                    unsigned less = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {  // broadcast reads, fully unrolled: no shuffles, no branches,
                        const Key kj = ckeys[j];    // all loads in flight at once (a loop bounded by `total` was slower)
                        less += (kj < mykey || (kj == mykey && j < lane)) ? 1u : 0u;
                    }
                    __syncwarp();
                    unsigned hit = __ballot_sync(FULL, (unsigned)lane < total && less == rem);
                    found = __shfl_sync(FULL, myraw, __ffs(hit) - 1);
                    if (rem + 1 < total) {
                        unsigned hit2 = __ballot_sync(FULL, (unsigned)lane < total && less == rem + 1);
                        found_next = __shfl_sync(FULL, myraw, __ffs(hit2) - 1);
                        have_next = true;
                    }
                }
                if (pass == 0) {
                    raw_r = found;
                    if (!want_next) return;
                    if (have_next) {
                        raw_next = found_next;
                        return;
                    }
                } else {
                    raw_next = found;
                }
            }
        };

        // ---- SMART variant of select: informative-bit windows (see the top of the kernel) ------------------
        auto select_smart = [&](unsigned r, bool want_next, Raw &raw_r, Raw &raw_next) {
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                if (planes_dirty) {
                    __syncwarp();
                    build_from_global_smart(reinterpret_cast<const unsigned char *>(lp), sh0, true);
                    __syncwarp();
                    planes_dirty = false;
                }
                uint32_t m[NB];
#pragma unroll
                for (int blk = 0; blk < NB; ++blk) m[blk] = (blk < nbn) ? init[blk * 32] : 0u;
                unsigned total = n_eff, rem = r + (unsigned)pass;
                bool neg = false;
                bool small = total <= 32, alleq = false;
                int sh = sh0, topb = 15;
                bool first = true;
#pragma unroll 1
                while (!small) {
                    const uint32_t *pp = planes + topb * NB * 32;
                    uint32_t inv = first ? (TOP_INVERTED ? 0xffffffffu : 0u) : ((IS_FLOAT && neg) ? 0xffffffffu : 0u);
#pragma unroll 1
                    for (int b = topb; b >= 0; --b, pp -= NB * 32) {
                        uint32_t x[NB];
                        unsigned ones = 0;
#pragma unroll
                        for (int blk = 0; blk < NB; ++blk) {
                            x[blk] = pp[blk * 32];
                            ones += __popc(m[blk] & (x[blk] ^ inv));
                        }
                        ones = __reduce_add_sync(FULL, ones);
                        const unsigned small_cnt = total - ones;
                        const bool take_small = rem < small_cnt;
                        total = take_small ? small_cnt : ones;
                        rem = take_small ? rem : rem - small_cnt;
                        const uint32_t keep = take_small ? inv : ~inv;
#pragma unroll
                        for (int blk = 0; blk < NB; ++blk) m[blk] &= ~(x[blk] ^ keep);
                        if (first && b == 15) {
                            neg = IS_FLOAT && (keep & 1u);
                            inv = neg ? 0xffffffffu : 0u;
                        }
                        if (total <= 32) {
                            small = true;
                            break;
                        }
                        // no informative bit below this one: the candidates are all one value
                        if ((V_lane & low_mask(sh + ((first && b == 15) ? 15 : b))) == 0) {
                            alleq = true;
                            break;
                        }
                    }
                    if (small || alleq) break;
                    // More than 32 candidates survived the window.  Before the lane is read and transposed again: are they
                    // all one value already?  (Tie-heavy lanes: usually yes once the few distinct values are told apart.)
                    {
                        Raw mine = 0;
                        bool have = false;
#pragma unroll
                        for (int blk = 0; blk < NB; ++blk)
                            if (!have && m[blk]) {
                                mine = to_raw<T>(lp[elem_index(blk, slot_of_pos(__ffs(m[blk]) - 1))]);
                                have = true;
                            }
                        const unsigned owners = __ballot_sync(FULL, have);
                        const Raw ref = __shfl_sync(FULL, mine, __ffs(owners) - 1);
                        bool same = true, go = true;
#pragma unroll
                        for (int blk = 0; blk < NB; ++blk) {
                            if (go) {
                                uint32_t bits = m[blk];
                                while (bits) {
                                    const int p = __ffs(bits) - 1;
                                    bits &= bits - 1;
                                    same = same && (to_raw<T>(lp[elem_index(blk, slot_of_pos(p))]) == ref);
                                }
                                go = __all_sync(FULL, same);
                            }
                        }
                        if (go) {
                            alleq = true;
                            break;
                        }
                    }
                    const int top = msb(V_lane & low_mask(sh));  // (non-zero: the walk would have stopped at b == 0)
                    const int nsh = max(top - 15, 0);
                    __syncwarp();
                    build_from_global_smart(reinterpret_cast<const unsigned char *>(lp), nsh, false);
                    __syncwarp();
                    planes_dirty = true;
                    sh = nsh;
                    topb = top - nsh;
                    first = false;
                }
                Raw found, found_next = 0;
                bool have_next = false;
                if (!small) {
                    // every remaining candidate is the same value: take the first one
                    Raw mine = 0;
                    bool have = false;
#pragma unroll
                    for (int blk = 0; blk < NB; ++blk)
                        if (!have && m[blk]) {
                            mine = to_raw<T>(lp[elem_index(blk, slot_of_pos(__ffs(m[blk]) - 1))]);
                            have = true;
                        }
                    const unsigned owners = __ballot_sync(FULL, have);
                    found = __shfl_sync(FULL, mine, __ffs(owners) - 1);
                    if (rem + 1 < total) {
                        found_next = found;
                        have_next = true;
                    }
                } else {
                    unsigned mycnt = 0;
#pragma unroll
                    for (int blk = 0; blk < NB; ++blk) mycnt += __popc(m[blk]);
                    unsigned incl = mycnt;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        unsigned tt = __shfl_up_sync(FULL, incl, d);
                        if (lane >= d) incl += tt;
                    }
                    unsigned base = incl - mycnt;
#pragma unroll
                    for (int blk = 0; blk < NB; ++blk) {
                        uint32_t bits = m[blk];
                        while (bits) {
                            int p = __ffs(bits) - 1;
                            bits &= bits - 1;
                            clist[base++] = elem_index(blk, slot_of_pos(p));
                        }
                    }
                    __syncwarp();
                    Raw myraw = 0;
                    Key mykey = (Key) ~(Key)0;
                    if ((unsigned)lane < total) {
                        T x = lp[clist[lane]];
                        myraw = to_raw<T>(x);
                        mykey = Tr::to_key(x);
                    }
                    ckeys[lane] = mykey;
                    __syncwarp();
                    unsigned less = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const Key kj = ckeys[j];
                        less += (kj < mykey || (kj == mykey && j < lane)) ? 1u : 0u;
                    }
                    __syncwarp();
                    unsigned hit = __ballot_sync(FULL, (unsigned)lane < total && less == rem);
                    found = __shfl_sync(FULL, myraw, __ffs(hit) - 1);
                    if (rem + 1 < total) {
                        unsigned hit2 = __ballot_sync(FULL, (unsigned)lane < total && less == rem + 1);
                        found_next = __shfl_sync(FULL, myraw, __ffs(hit2) - 1);
                        have_next = true;
                    }
                }
                if (pass == 0) {
                    raw_r = found;
                    if (!want_next) return;
                    if (have_next) {
                        raw_next = found_next;
                        return;
                    }
                } else {
                    raw_next = found;
                }
            }
        };

        // ---- every requested quantile of the lane (src/quantile/mod.rs:475-487) ------------------------
        unsigned cache_rank[2] = {0xffffffffu, 0xffffffffu};
        Raw cache_raw[2] = {0, 0};
        auto remember = [&](unsigned r, Raw v) {
            cache_rank[1] = cache_rank[0];
            cache_raw[1] = cache_raw[0];
            cache_rank[0] = r;
            cache_raw[0] = v;
        };
        auto lookup = [&](unsigned r, Raw &v) -> bool {
            if (r == cache_rank[0]) { v = cache_raw[0]; return true; }
            if (r == cache_rank[1]) { v = cache_raw[1]; return true; }
            return false;
        };
#pragma unroll 1
        for (int t = 0; t < qa.nq; ++t) {
            const RankMath rm = slot_rank_math(qa, t, n_eff);
            const bool nl = needs_lower(qa.interp, rm.frac), nh = needs_higher(qa.interp, rm.frac);
            const unsigned rl = (unsigned)rm.lower, rh = (unsigned)rm.higher;
            Raw vl = 0, vh = 0;
#pragma unroll 1
            for (;;) {
                const bool have_l = !nl || lookup(rl, vl);
                const bool have_h = !nh || lookup(rh, vh);
                if (have_l && have_h) break;
                const unsigned r = !have_l ? rl : rh;
                const bool want_next = !have_l && !have_h && rh == rl + 1;
                Raw a = 0, b = 0;
                if constexpr (SMART) select_smart(r, want_next, a, b);
                else select(r, want_next, a, b);
                remember(r, a);
                if (want_next) remember(r + 1, b);
            }
            if (lane == 0) {
                out[out_off + (long long)t * g.out_axis_stride] =
                    interpolate<T>(qa.interp, from_raw<T>(vl), from_raw<T>(vh), rm.frac, qa.status);
            }
        }
        __syncwarp();  // the plane area is rewritten by the next lane's build
    }
}

template <typename T> constexpr bool short_type_ok() {
    return std::is_same<T, float>::value || std::is_same<T, double>::value || std::is_same<T, int32_t>::value ||
           std::is_same<T, uint32_t>::value || std::is_same<T, int64_t>::value || std::is_same<T, uint64_t>::value;
}

struct ShortPlan {
    bool ok = false;
    int nb = 0, warps = 0, stages = 0;
    size_t smem = 0;
};

ShortPlan plan_short(const nds_ctx *ctx, const DeviceCall &c) {
    ShortPlan p;
    size_t es = 0;
    switch (c.dtype) {
    case NDS_F32: case NDS_I32: case NDS_U32: es = 4; break;
    case NDS_F64: case NDS_I64: case NDS_U64: es = 8; break;
    default: return p;
    }
    const LaneGeom &g = c.geom;
    if (c.valid || c.out_valid) return p;
    if (g.axis_stride != 1 || g.axis_len < 64 || g.axis_len > 8192) return p;
    if ((g.axis_len * es) % 16 != 0) return p;  // TMA bulk copies move multiples of 16 bytes from 16-byte aligned addresses
    if (((uintptr_t)c.data) % 16 != 0) return p;
    for (int d = 0; d < g.n_outer; ++d)
        if (g.oshape[d] > 1 && ((g.in_ostride[d] * (long long)es) % 16) != 0) return p;
    if (g.nlanes < 64) return p;  // too few lanes to fill the machine warp-per-lane
    p.nb = g.axis_len <= 1024 ? 1 : (g.axis_len <= 2048 ? 2 : (g.axis_len <= 4096 ? 4 : 8));
    const size_t fixed = (p.nb == 1 ? short_warp_fixed_bytes<1>() : (p.nb == 2 ? short_warp_fixed_bytes<2>()
                          : (p.nb == 4 ? short_warp_fixed_bytes<4>() : short_warp_fixed_bytes<8>()))) + SHORT_WARP_TAIL_BYTES;
    const size_t blk_bytes = 1024 * es;
    const size_t budget = (size_t)ctx->max_smem_optin > 4096 ? (size_t)ctx->max_smem_optin - 1024 : 0;
    // Warps hide the latency of the serial bit walk; ring stages keep HBM requests in flight.  Measured on
    // B200 (C2 workload, tools/sweep_short.sh): 16 warps x 1 stage 67 %, 12 x 2 62 %, 8 x 3 48 % of the HBM
    // roofline -- so take as many warps as fit first, then as many stages as still fit.
    p.stages = 1;
    p.warps = (int)std::min<size_t>(16, budget / (p.stages * blk_bytes + fixed));
    while (p.stages < 2 && budget / ((p.stages + 1) * blk_bytes + fixed) >= (size_t)p.warps) p.stages++;
    if (const char *e = getenv("NDS_SHORT_STAGES")) {  // tuning knobs
        p.stages = std::min(SHORT_MAX_STAGES, std::max(1, atoi(e)));
        p.warps = (int)std::min<size_t>(16, budget / (p.stages * blk_bytes + fixed));
    }
    p.warps = p.warps / 4 * 4 > 0 ? p.warps / 4 * 4 : p.warps;  // whole warps per scheduler
    // few lanes (BASELINE config 1: 1024 lanes): spread them over all SMs instead of filling 16 warps on a few
    {
        const unsigned long long per_sm = (g.nlanes + (unsigned long long)ctx->sm_count - 1) / (unsigned long long)ctx->sm_count;
        if (per_sm < (unsigned long long)p.warps) p.warps = (int)std::max<unsigned long long>(1, per_sm);
    }
    if (const char *e = getenv("NDS_SHORT_WARPS")) p.warps = std::min(p.warps, std::max(1, atoi(e)));
    if (p.warps < 1) return p;
    p.smem = (size_t)p.warps * (p.stages * blk_bytes + fixed);
    p.ok = true;
    return p;
}

template <typename T, int NB>
cudaError_t launch_short_nb(nds_ctx *ctx, const DeviceCall &c, const ShortPlan &p) {
    auto kernel = short_bitslice_kernel<T, NB>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return e;
    unsigned long long need = (c.geom.nlanes + p.warps - 1) / p.warps;
    unsigned grid = (unsigned)std::min<unsigned long long>(need, (unsigned long long)ctx->sm_count);
    kernel<<<grid, p.warps * 32, p.smem, ctx->stream>>>((const T *)c.data, c.geom, c.q, (T *)c.out, p.stages);
    ctx->launches += 1;
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_short_typed(nds_ctx *ctx, const DeviceCall &c, const ShortPlan &p) {
    switch (p.nb) {
    case 1: return launch_short_nb<T, 1>(ctx, c, p);
    case 2: return launch_short_nb<T, 2>(ctx, c, p);
    case 4: return launch_short_nb<T, 4>(ctx, c, p);
    default: return launch_short_nb<T, 8>(ctx, c, p);
    }
}

}  // namespace

bool short_bitslice_supported(const nds_ctx *ctx, const DeviceCall &c) { return plan_short(ctx, c).ok; }

cudaError_t launch_short_bitslice(nds_ctx *ctx, const DeviceCall &c) {
    ShortPlan p = plan_short(ctx, c);
    if (!p.ok) return cudaErrorNotSupported;
    ctx->last_path = "short_bitslice";
    switch (c.dtype) {
    case NDS_F32: return launch_short_typed<float>(ctx, c, p);
    case NDS_F64: return launch_short_typed<double>(ctx, c, p);
    case NDS_I32: return launch_short_typed<int32_t>(ctx, c, p);
    case NDS_U32: return launch_short_typed<uint32_t>(ctx, c, p);
    case NDS_I64: return launch_short_typed<int64_t>(ctx, c, p);
    case NDS_U64: return launch_short_typed<uint64_t>(ctx, c, p);
    default: return cudaErrorNotSupported;
    }
}

}  // namespace nds
