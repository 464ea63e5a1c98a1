// Contiguous lanes that are too long for one warp (8193 ... 786432 elements) -- and calls with too few lanes for the
// warp-per-lane kernel to fill the machine: one CTA of 12 warps per lane, or a thread-block CLUSTER of 2, 4 or 8 CTAs
// per lane beyond 98304 elements; bit-sliced MSB-first selection.
//
// The lane is read ONCE from HBM, straight into registers (16-byte loads, a warp takes every 12th block of 1024
// elements: 12 x 8 KB in flight per SM without any staging buffer), transposed into 16 bit planes in shared memory
// (2 bytes per element: 192 KB for 98304 elements; a cluster splits the lane's blocks over its CTAs) and the planes are
// then walked by all threads: one popc per plane word and thread, REDUX per warp, one barrier per key bit -- a CTA
// barrier, or a cluster barrier with the per-warp counts read from the other CTAs' shared memory (DSMEM).  Up to three
// ranks are walked together (one read of the plane words, one barrier per bit for all of them).  The 16-bit window of
// a round is chosen from the bits that tell the lane's elements apart (see nds_short.cu, SMART; here with a 16-bit
// no-sign mode when the whole lane has one sign): the first block of every warp proposes it, the whole lane confirms
// it.  At most 128 survivors per rank are collected in the first CTA and ranked by full key.  Later windows (rare: more
// than 128 elements agree on the window's informative bits) re-read the lane from L2.
//
// Replaces generic_lanes_kernel (one digit pass per rank and key byte, one shared atomic per element) for these shapes:
// src/quantile/mod.rs:470-488 treats every lane alike, so must we.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "nds_bitslice.cuh"
#include "nds_internal.h"

namespace cg = cooperative_groups;

namespace nds {

namespace {

constexpr int MID_THREADS = 384;
constexpr int MID_WARPS = MID_THREADS / 32;
constexpr int MID_MAXW = 8;                       // plane words per thread
constexpr unsigned MID_CTA_N = MID_MAXW * MID_THREADS * 32;   // 98304 elements per CTA
constexpr int MID_MAX_CLUSTER = 8;
constexpr unsigned MID_MAX_N = MID_CTA_N * MID_MAX_CLUSTER;
constexpr int MID_CAND = 128;                     // survivors ranked directly
constexpr int MID_K = 3;                          // ranks walked together (MID_K * MID_CAND == MID_THREADS)

struct MidShared {          // small fixed part in front of the planes
    unsigned partial[2][MID_K][MID_MAX_CLUSTER * MID_WARPS];   // per-warp counts of one key bit, of EVERY CTA of the cluster
    unsigned stat[2][6];    // [stage][zlo, zhi, olo, ohi, (unused), nan count]
    unsigned ncand[MID_K];
    unsigned sel[MID_K][2];
    unsigned long long res[MID_K][2];   // first CTA: the raw values at ranks r and r + 1 of every rank of the group
    unsigned clist[MID_K][MID_CAND];
    unsigned long long ckeys[MID_K][MID_CAND];
    unsigned long long craw[MID_K][MID_CAND];
};
static_assert(MID_K * MID_CAND == MID_THREADS, "one thread per survivor of a group");

// CL: the lane is shared by the CTAs of a cluster (CTA c owns blocks [c * bpc, (c + 1) * bpc) of 1024 elements)
template <typename T, bool CL>
__global__ void __launch_bounds__(MID_THREADS, 1)
mid_bitslice_kernel(const T *__restrict__ data, LaneGeom g, QuantArgs qa, T *__restrict__ out, unsigned nwords, unsigned bpc) {
    using Tr = Traits<T>;
    using Key = typename Tr::Key;
    using Raw = typename RawBits<T>::U;
    constexpr int ES = (int)sizeof(T);
    constexpr int VEC = 16 / ES;
    constexpr int LD_PER_BLK = 32 / VEC;
    constexpr int KEYBITS = ES * 8;
    constexpr bool IS_FLOAT = Tr::is_float;
    constexpr bool SIGNED_INT = !IS_FLOAT && std::is_signed<T>::value;
    constexpr bool TOP_INVERTED = IS_FLOAT || std::is_signed<T>::value;
    constexpr Raw SIGN = (Raw)1 << (KEYBITS - 1);
    constexpr unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    MidShared &sh = *reinterpret_cast<MidShared *>(smem_raw);
    uint32_t *planes = reinterpret_cast<uint32_t *>(smem_raw + sizeof(MidShared));   // [16][nwords]
    uint32_t *im = planes + 16u * nwords;                                              // [nwords] valid and not NaN

    cg::cluster_group cluster = cg::this_cluster();
    const unsigned csize = CL ? cluster.num_blocks() : 1u;
    const unsigned crank = CL ? cluster.block_rank() : 0u;
    auto sync_all = [&]() {
        if constexpr (CL) cluster.sync();
        else __syncthreads();
    };
    auto peer = [&](unsigned r) -> MidShared * {   // the fixed part of CTA r's shared memory
        if constexpr (CL) return cluster.map_shared_rank(&sh, r);
        else return &sh;
    };

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned n = (unsigned)g.axis_len;
    const unsigned n_chunks = (n * ES + 15) / 16;   // (a ragged last chunk only with padded lanes, see plan_mid)
    const int nb = (int)((n + 1023u) / 1024u);
    const int blk0 = (int)(crank * bpc);                                    // first block of this CTA
    const int nbl = max(0, min((int)bpc, nb - blk0));                       // its number of blocks
    const unsigned long long n_groups = gridDim.x / csize, my_group = blockIdx.x / csize;

    auto slot_of_pos = [](int p) -> int { return 2 * (p & 15) + (p >> 4); };
    // element index of slot s of the block-thread (blk, ln)
    auto elem_index = [&](int blk, int ln, int s) -> unsigned {
        if (ES == 4) return (unsigned)(((blk * LD_PER_BLK + (s >> 2)) * 32 + ln) * 4 + (s & 3));
        return (unsigned)(((blk * LD_PER_BLK + (s >> 1)) * 32 + ln) * 2 + (s & 1));
    };
    // element index of bit p of this CTA's plane word w
    auto elem_of = [&](unsigned w, int p) -> unsigned { return elem_index(blk0 + (int)(w >> 5), (int)(w & 31u), slot_of_pos(p)); };
    auto valid_mask = [&](int blk, int ln) -> uint32_t {
        if ((unsigned)(blk + 1) * 1024u <= n) return FULL;
        uint32_t v = 0;
        for (int p = 0; p < 32; ++p)
            if (elem_index(blk, ln, slot_of_pos(p)) < n) v |= 1u << p;
        return v;
    };
    auto field64 = [&](uint32_t lo, uint32_t hi, int s, bool first) -> uint32_t {
        uint32_t v = s >= 32 ? (hi >> (s - 32)) : __funnelshift_r(lo, hi, s);
        if (first) v = bitsel(v, hi >> 16, 0x7fffu);
        return v;
    };
    auto field32 = [&](uint32_t x, int s, bool first) -> uint32_t {
        uint32_t v = x >> s;
        if (first) v = bitsel(v, x >> 16, 0x7fffu);
        return v;
    };
    auto pack = [&](const uint4 &q, uint32_t (&w)[16], int j, int s, bool first) {
        if constexpr (ES == 4) {
            w[2 * j] = __byte_perm(field32(q.x, s, first), field32(q.y, s, first), 0x5410u);
            w[2 * j + 1] = __byte_perm(field32(q.z, s, first), field32(q.w, s, first), 0x5410u);
        } else {
            w[j] = __byte_perm(field64(q.x, q.y, s, first), field64(q.z, q.w, s, first), 0x5410u);
        }
    };
    auto msb = [](Raw v) -> int { return ES == 8 ? 63 - __clzll((long long)v) : 31 - __clz((int)v); };
    auto low_mask = [](int pos) -> Raw { return pos >= KEYBITS ? ~(Raw)0 : (((Raw)1 << pos) - (Raw)1); };
    // shift of the round-0 window: sign bit + 15 bits, or -- when the whole lane has one sign -- 16 bits, below the highest
    // informative bit (+ margin while only a part of the lane has been seen)
    auto pick_shift = [&](Raw V, int margin, bool no_sign) -> int {
        if (V == 0) return KEYBITS - 17;
        const int top = min(msb(V) + margin, KEYBITS - 2);
        return max(top - (no_sign ? 15 : 14), 0);
    };

    for (unsigned long long L = my_group; L < g.nlanes; L += n_groups) {
        long long in_off, out_off;
        lane_offsets(g, L, in_off, out_off);
        const T *lp = data + in_off;
        const unsigned char *gsrc = reinterpret_cast<const unsigned char *>(lp);

        // ---- statistics of the lane (see nds_short.cu): z = bits that differ from the first element somewhere,
        //      o = bits that are not a copy of the sign (signed integers), am = an all-ones exponent is present ----
        uint32_t x0lo, x0hi = 0;
        {
            const Raw x0 = to_raw<T>(lp[0]);
            x0lo = (uint32_t)x0;
            if constexpr (ES == 8) x0hi = (uint32_t)((unsigned long long)x0 >> 32);
        }
        uint32_t zlo = 0, zhi = 0, olo = 0, ohi = 0;
        auto stats = [&](const uint4 &q, uint32_t &am) {
            if constexpr (ES == 4) {
                const uint32_t e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    zlo |= e[u] ^ x0lo;
                    if constexpr (SIGNED_INT) olo |= e[u] ^ (uint32_t)((int32_t)e[u] >> 31);
                    if constexpr (IS_FLOAT) am = max(am, e[u] & 0x7fffffffu);
                }
            } else {
                zlo |= (q.x ^ x0lo) | (q.z ^ x0lo);
                zhi |= (q.y ^ x0hi) | (q.w ^ x0hi);
                if constexpr (SIGNED_INT) {
                    const uint32_t s0 = (uint32_t)((int32_t)q.y >> 31), s1 = (uint32_t)((int32_t)q.w >> 31);
                    olo |= (q.x ^ s0) | (q.z ^ s1);
                    ohi |= (q.y ^ s0) | (q.w ^ s1);
                }
                if constexpr (IS_FLOAT) am = max(am, max(q.y & 0x7fffffffu, q.w & 0x7fffffffu));
            }
        };
        constexpr uint32_t SPECIAL = ES == 8 ? 0x7ff00000u : 0x7f800000u;   // |x| (high word) with an all-ones exponent
        auto publish_stats = [&](int stage) {  // warp OR -> shared OR (this CTA's share)
            const uint32_t a = __reduce_or_sync(FULL, zlo), b = __reduce_or_sync(FULL, zhi);
            const uint32_t c = __reduce_or_sync(FULL, olo), d = __reduce_or_sync(FULL, ohi);
            if (lane == 0) {
                atomicOr(&sh.stat[stage][0], a);
                atomicOr(&sh.stat[stage][1], b);
                atomicOr(&sh.stat[stage][2], c);
                atomicOr(&sh.stat[stage][3], d);
            }
        };
        // the statistics of the whole lane: OR (sum for the NaN count) over the CTAs of the cluster
        uint32_t cs[6];
        auto combine_stats = [&](int stage) {
#pragma unroll
            for (int i = 0; i < 6; ++i) cs[i] = 0;
            for (unsigned r = 0; r < csize; ++r) {
                const MidShared *p = peer(r);
#pragma unroll
                for (int i = 0; i < 4; ++i) cs[i] |= p->stat[stage][i];
                cs[5] += p->stat[stage][5];
            }
        };
        auto informative = [&]() -> Raw {
            uint32_t lo = cs[0], hi = cs[1];
            if constexpr (SIGNED_INT) {
                lo &= cs[2];
                hi &= cs[3];
            }
            if constexpr (ES == 8) return (((Raw)hi << 32) | (Raw)lo) & ~SIGN;
            else return (Raw)lo & ~SIGN;
        };
        auto sign_constant = [&]() -> bool { return ((ES == 8 ? cs[1] : cs[0]) & 0x80000000u) == 0u; };
        auto load_block = [&](int blk, uint4 (&q)[LD_PER_BLK]) {
#pragma unroll
            for (int j = 0; j < LD_PER_BLK; ++j) {
                const unsigned chunk = min((unsigned)((blk * LD_PER_BLK + j) * 32 + lane), n_chunks - 1);
                q[j] = __ldg(reinterpret_cast<const uint4 *>(gsrc + (size_t)chunk * 16));
            }
        };
        auto store_block = [&](uint32_t (&w)[16], int lb) {
            transpose16(w);
#pragma unroll
            for (int b = 0; b < 16; ++b) planes[(size_t)b * nwords + lb * 32 + lane] = w[b];
        };
        // this CTA's planes of one window from global memory (the lane is L2-resident after the first pass)
        auto rebuild = [&](int s, bool first) {
            for (int lb = warp; lb < nbl; lb += MID_WARPS) {
                uint4 q[LD_PER_BLK];
                load_block(blk0 + lb, q);
                uint32_t w[16];
#pragma unroll
                for (int j = 0; j < LD_PER_BLK; ++j) pack(q[j], w, j, s, first);
                store_block(w, lb);
            }
            __syncthreads();
        };

        if (tid < 12) (&sh.stat[0][0])[tid] = 0;
        __syncthreads();

        // ---- pass over the lane: the first block of every warp proposes the window, all blocks confirm it ----
        int sh0;
        bool signless;   // the window carries no sign bit (every element of the lane has the sign of the first one)
        unsigned my_nan = 0;
        {
            uint4 q[LD_PER_BLK];
            uint32_t am = 0;
            if (warp < nbl) {
                load_block(blk0 + warp, q);
#pragma unroll
                for (int j = 0; j < LD_PER_BLK; ++j) stats(q[j], am);
            }
            publish_stats(0);
            sync_all();
            combine_stats(0);
            signless = sign_constant();
            sh0 = pick_shift(informative(), bpc > (unsigned)MID_WARPS ? 1 : 0, signless);
            for (int lb = warp; lb < nbl; lb += MID_WARPS) {
                if (lb != warp) {
                    load_block(blk0 + lb, q);
                    am = 0;
#pragma unroll
                    for (int j = 0; j < LD_PER_BLK; ++j) stats(q[j], am);
                }
                uint32_t w[16];
#pragma unroll
                for (int j = 0; j < LD_PER_BLK; ++j) pack(q[j], w, j, sh0, !signless);
                uint32_t keep = valid_mask(blk0 + lb, lane);
                if constexpr (IS_FLOAT) {
                    uint32_t nanm = 0;
                    if (__any_sync(FULL, am >= SPECIAL)) {
#pragma unroll
                        for (int j = 0; j < LD_PER_BLK; ++j) {
                            if constexpr (ES == 8) {
                                const double d0 = __hiloint2double((int)q[j].y, (int)q[j].x);
                                const double d1 = __hiloint2double((int)q[j].w, (int)q[j].z);
                                nanm |= (d0 != d0 ? 1u : 0u) << j;
                                nanm |= (d1 != d1 ? 1u : 0u) << (j + 16);
                            } else {  // slots 4j .. 4j+3: positions 2j, 2j+16, 2j+1, 2j+17
                                const float f0 = __uint_as_float(q[j].x), f1 = __uint_as_float(q[j].y);
                                const float f2 = __uint_as_float(q[j].z), f3 = __uint_as_float(q[j].w);
                                nanm |= (f0 != f0 ? 1u : 0u) << (2 * j);
                                nanm |= (f1 != f1 ? 1u : 0u) << (2 * j + 16);
                                nanm |= (f2 != f2 ? 1u : 0u) << (2 * j + 1);
                                nanm |= (f3 != f3 ? 1u : 0u) << (2 * j + 17);
                            }
                        }
                    }
                    my_nan += __popc(keep & nanm);
                    keep &= ~nanm;
                }
                im[lb * 32 + lane] = keep;
                store_block(w, lb);
            }
            publish_stats(1);
            if constexpr (IS_FLOAT) {
                const unsigned wn = __reduce_add_sync(FULL, my_nan);
                if (lane == 0 && wn) atomicAdd(&sh.stat[1][5], wn);
            }
            sync_all();
            combine_stats(1);
        }
        const Raw V = informative();
        const unsigned nan_total = IS_FLOAT ? cs[5] : 0u;
        if ((V & ~low_mask(sh0 + (signless ? 16 : 15))) != 0 || (signless && !sign_constant())) {
            // (rare) informative bits above the proposed window, or a sign the first blocks did not show
            signless = sign_constant();
            sh0 = pick_shift(V, 0, signless);
            rebuild(sh0, !signless);
        }
        const bool lane_neg = IS_FLOAT && (((ES == 8 ? x0hi : x0lo) & 0x80000000u) != 0u);   // (meaningful when signless)
        if (nan_total && !qa.skipnan && tid == 0 && crank == 0) atomicOr(qa.status, STATUS_BIT_NAN);
        const unsigned n_eff = n - nan_total;   // (non-skipnan with NaNs: status raised, result meaningless but in bounds)

        if (n_eff == 0) {
            if (crank == 0)
                for (int t = tid; t < qa.nq; t += MID_THREADS) out[out_off + (long long)t * g.out_axis_stride] = canonical_nan<T>();
            sync_all();
            continue;
        }

        bool planes_dirty = false;
        unsigned step_no = 0;   // parity of the partial-sum buffer (identical in every thread of the cluster)

        // ---- up to MID_K ranks walked TOGETHER through the round-0 planes: one read of the plane words and one barrier
        //      per key bit serve all of them (get_many_from_sorted_mut's point, src/sort.rs:184-283).  hb[k]: rank r + 1
        //      was among r's survivors and is returned in b[k]. ---------------------------------------------------
        auto run_group = [&](const unsigned (&rk)[MID_K], int K, Raw (&a)[MID_K], Raw (&b)[MID_K], bool (&hb)[MID_K]) {
            if (planes_dirty) {
                __syncthreads();
                rebuild(sh0, !signless);
                planes_dirty = false;
            }
            uint32_t m[MID_K][MID_MAXW];
            unsigned total[MID_K], rem[MID_K];
            uint32_t inv[MID_K];
            bool open[MID_K], small[MID_K], alleq[MID_K];
#pragma unroll
            for (int k = 0; k < MID_K; ++k) {
#pragma unroll
                for (int i = 0; i < MID_MAXW; ++i) {
                    const unsigned w = (unsigned)tid + (unsigned)MID_THREADS * i;
                    m[k][i] = (k < K && w < (unsigned)nbl * 32u) ? im[w] : 0u;
                }
                total[k] = n_eff;
                rem[k] = k < K ? rk[k] : 0u;
                inv[k] = signless ? (lane_neg ? FULL : 0u) : (TOP_INVERTED ? FULL : 0u);
                small[k] = k < K && total[k] <= (unsigned)MID_CAND;
                alleq[k] = false;
                open[k] = k < K && !small[k];
                hb[k] = false;
            }
            // one key bit for the ranks in `act`: counts of this CTA, exchange, decision
            auto step = [&](int bit, int s, bool sign_step, const bool (&act)[MID_K]) {
                const uint32_t *pp = planes + (size_t)bit * nwords + tid;
                uint32_t x[MID_MAXW];
#pragma unroll
                for (int i = 0; i < MID_MAXW; ++i) x[i] = ((unsigned)tid + (unsigned)MID_THREADS * i < (unsigned)nbl * 32u) ? pp[MID_THREADS * i] : 0u;
                const unsigned par = step_no & 1u;
                step_no += 1;
#pragma unroll
                for (int k = 0; k < MID_K; ++k) {
                    if (!act[k]) continue;   // (uniform over the cluster)
                    unsigned ones = 0;
#pragma unroll
                    for (int i = 0; i < MID_MAXW; ++i) ones += __popc(m[k][i] & (x[i] ^ inv[k]));
                    ones = __reduce_add_sync(FULL, ones);
                    // pushed into every CTA's table (lane r stores to CTA r): after the barrier everybody sums locally
                    if ((unsigned)lane < csize) peer((unsigned)lane)->partial[par][k][crank * (unsigned)MID_WARPS + (unsigned)warp] = ones;
                }
                sync_all();
#pragma unroll
                for (int k = 0; k < MID_K; ++k) {
                    if (!act[k]) continue;
                    unsigned v = 0;
                    for (unsigned i = (unsigned)lane; i < csize * (unsigned)MID_WARPS; i += 32u) v += sh.partial[par][k][i];
                    const unsigned ones = __reduce_add_sync(FULL, v);
                    const unsigned small_cnt = total[k] - ones;
                    const bool take_small = rem[k] < small_cnt;
                    total[k] = take_small ? small_cnt : ones;
                    rem[k] = take_small ? rem[k] : rem[k] - small_cnt;
                    const uint32_t keep = take_small ? inv[k] : ~inv[k];
#pragma unroll
                    for (int i = 0; i < MID_MAXW; ++i) m[k][i] &= ~(x[i] ^ keep);
                    if (sign_step) inv[k] = (IS_FLOAT && (keep & 1u)) ? FULL : 0u;
                    if (total[k] <= (unsigned)MID_CAND) {
                        small[k] = true;
                        open[k] = false;
                    } else if ((V & low_mask(s + (sign_step ? 15 : bit))) == 0) {   // no informative bit below this one
                        alleq[k] = true;
                        open[k] = false;
                    }
                }
            };
            // round 0: all ranks of the group
#pragma unroll 1
            for (int bit = 15; bit >= 0; --bit) {
                bool any_open = false;
#pragma unroll
                for (int k = 0; k < MID_K; ++k) any_open = any_open || open[k];
                if (!any_open) break;
                step(bit, sh0, bit == 15 && !signless, open);
            }
            // (rare) a rank that more than MID_CAND elements still share: further windows, that rank alone
#pragma unroll
            for (int k = 0; k < MID_K; ++k) {
                int s = sh0;
#pragma unroll 1
                while (open[k]) {
                    const int top = msb(V & low_mask(s));   // (non-zero: the rank would have been closed as all-equal)
                    const int ns = max(top - 15, 0);
                    bool act[MID_K];
#pragma unroll
                    for (int kk = 0; kk < MID_K; ++kk) act[kk] = kk == k;
                    sync_all();                               // every CTA has finished reading the old planes' partials
                    rebuild(ns, false);
                    planes_dirty = true;
                    s = ns;
#pragma unroll 1
                    for (int bit = top - ns; bit >= 0 && open[k]; --bit) step(bit, s, false, act);
                }
            }
            // ---- survivors of all ranks at once, collected in the first CTA: thread (k, j) owns survivor j of rank k ----
            sync_all();
            if (crank == 0 && tid < MID_K) sh.ncand[tid] = 0;
            sync_all();
            MidShared *s0 = peer(0);
#pragma unroll
            for (int k = 0; k < MID_K; ++k) {
                if (small[k]) {
#pragma unroll
                    for (int i = 0; i < MID_MAXW; ++i) {
                        uint32_t bits = m[k][i];
                        if (bits) {
                            const unsigned w = (unsigned)tid + (unsigned)MID_THREADS * i;
                            unsigned base = atomicAdd(&s0->ncand[k], (unsigned)__popc(bits));
                            while (bits) {
                                const int p = __ffs(bits) - 1;
                                bits &= bits - 1;
                                s0->clist[k][base++] = elem_of(w, p);
                            }
                        }
                    }
                } else if (alleq[k]) {
                    bool have = false;
#pragma unroll
                    for (int i = 0; i < MID_MAXW; ++i)
                        if (!have && m[k][i]) {
                            const unsigned w = (unsigned)tid + (unsigned)MID_THREADS * i;
                            if (atomicAdd(&s0->ncand[k], 1u) == 0u) s0->clist[k][0] = elem_of(w, __ffs(m[k][i]) - 1);
                            have = true;
                        }
                }
            }
            sync_all();
            if (crank == 0) {
                const int mk = tid / MID_CAND;
                const unsigned mj = (unsigned)(tid % MID_CAND);
                unsigned tot_mine = 0, rem_mine = 0;
                bool small_mine = false, alleq_mine = false;
#pragma unroll
                for (int k = 0; k < MID_K; ++k)
                    if (k == mk) {
                        tot_mine = total[k];
                        rem_mine = rem[k];
                        small_mine = small[k];
                        alleq_mine = alleq[k];
                    }
                Key mykey = 0;
                if ((small_mine && mj < tot_mine) || (alleq_mine && mj == 0)) {
                    const T xv = lp[sh.clist[mk][mj]];
                    mykey = Tr::to_key(xv);
                    sh.ckeys[mk][mj] = (unsigned long long)mykey;
                    sh.craw[mk][mj] = (unsigned long long)to_raw<T>(xv);
                }
                __syncthreads();
                if (small_mine && mj < tot_mine) {
                    unsigned less = 0;
                    for (unsigned j = 0; j < tot_mine; ++j) {
                        const Key kj = (Key)sh.ckeys[mk][j];
                        less += (kj < mykey || (kj == mykey && j < mj)) ? 1u : 0u;
                    }
                    if (less == rem_mine) sh.res[mk][0] = sh.craw[mk][mj];
                    if (less == rem_mine + 1) sh.res[mk][1] = sh.craw[mk][mj];
                } else if (alleq_mine && mj == 0) {
                    sh.res[mk][0] = sh.res[mk][1] = sh.craw[mk][0];
                }
            }
            sync_all();
#pragma unroll
            for (int k = 0; k < MID_K; ++k)
                if (k < K) {
                    a[k] = (Raw)s0->res[k][0];
                    if (rem[k] + 1 < total[k]) {
                        b[k] = (Raw)s0->res[k][1];
                        hb[k] = true;
                    }
                }
            // (the next use of the shared lists starts with a barrier of its own)
        };

        // ---- every requested quantile of the lane (src/quantile/mod.rs:475-487), MID_K at a time ----------------
#pragma unroll 1
        for (int t0 = 0; t0 < qa.nq; t0 += MID_K) {
            const int K = min(MID_K, qa.nq - t0);
            RankMath rm[MID_K];
            unsigned rk[MID_K], rk2[MID_K];
            bool nl[MID_K], nh[MID_K], hb[MID_K], hb2[MID_K];
            Raw va[MID_K], vb[MID_K], va2[MID_K], vb2[MID_K];
            bool need2 = false;
#pragma unroll
            for (int k = 0; k < MID_K; ++k) {
                rk[k] = rk2[k] = 0;
                nl[k] = nh[k] = false;
                va[k] = vb[k] = va2[k] = vb2[k] = 0;
                if (k < K) {
                    rm[k] = slot_rank_math(qa, t0 + k, n_eff);
                    nl[k] = needs_lower(qa.interp, rm[k].frac);
                    nh[k] = needs_higher(qa.interp, rm[k].frac);
                    rk[k] = nl[k] ? (unsigned)rm[k].lower : (unsigned)rm[k].higher;
                }
            }
            run_group(rk, K, va, vb, hb);
            // the higher neighbour where it is wanted and was not among the lower rank's survivors
#pragma unroll
            for (int k = 0; k < MID_K; ++k) {
                const bool both = k < K && nl[k] && nh[k];
                const bool same = both && (unsigned)rm[k].higher == (unsigned)rm[k].lower;
                if (same) { vb[k] = va[k]; hb[k] = true; }
                rk2[k] = (both && !hb[k]) ? (unsigned)rm[k].higher : rk[k];
                need2 = need2 || (both && !hb[k]);
            }
            if (need2) {  // (uniform over the cluster)
                run_group(rk2, K, va2, vb2, hb2);
#pragma unroll
                for (int k = 0; k < MID_K; ++k)
                    if (k < K && nl[k] && nh[k] && !hb[k]) vb[k] = va2[k];
            }
            if (tid == 0 && crank == 0) {
#pragma unroll
                for (int k = 0; k < MID_K; ++k)
                    if (k < K) {
                        const Raw lo = va[k], hi = (nl[k] && nh[k]) ? vb[k] : va[k];
                        out[out_off + (long long)(t0 + k) * g.out_axis_stride] =
                            interpolate<T>(qa.interp, from_raw<T>(lo), from_raw<T>(hi), rm[k].frac, qa.status);
                    }
            }
        }
        sync_all();   // planes, masks and statistics are rewritten by the next lane
    }
}

struct MidPlan {
    bool ok = false;
    unsigned nwords = 0, bpc = 0;
    int csize = 1;
    size_t smem = 0;
};

MidPlan plan_mid(const nds_ctx *ctx, const DeviceCall &c) {
    MidPlan p;
    size_t es = 0;
    switch (c.dtype) {
    case NDS_F32: case NDS_I32: case NDS_U32: es = 4; break;
    case NDS_F64: case NDS_I64: case NDS_U64: es = 8; break;
    default: return p;
    }
    const LaneGeom &g = c.geom;
    if (c.valid || c.out_valid) return p;
    if (g.axis_stride != 1 || g.axis_len < 64 || g.axis_len > MID_MAX_N) return p;
    if ((g.axis_len * es) % 16 != 0 && !c.lanes_padded) return p;   // whole 16-byte chunks from 16-byte aligned lanes
    if (((uintptr_t)c.data) % 16 != 0) return p;
    for (int d = 0; d < g.n_outer; ++d)
        if (g.oshape[d] > 1 && ((g.in_ostride[d] * (long long)es) % 16) != 0) return p;
    const unsigned nb = (unsigned)((g.axis_len + 1023) / 1024);
    p.csize = 1;
    while ((unsigned long long)p.csize * MID_CTA_N < g.axis_len) p.csize *= 2;
    static const int max_cluster = getenv("NDS_MID_MAX_CLUSTER") ? atoi(getenv("NDS_MID_MAX_CLUSTER")) : MID_MAX_CLUSTER;
    if (p.csize > max_cluster) return p;
    if (const char *e = getenv("NDS_MID_CLUSTER")) {   // testing: a cluster for lanes that one CTA could hold
        const int want = atoi(e);
        if ((want == 2 || want == 4 || want == 8) && want > p.csize && nb >= (unsigned)want) p.csize = want;
    }
    p.bpc = (nb + p.csize - 1) / p.csize;
    p.nwords = p.bpc * 32u;
    p.smem = sizeof(MidShared) + (size_t)17 * p.nwords * 4;
    if (p.smem > (size_t)ctx->max_smem_optin) return p;
    p.ok = true;
    return p;
}

template <typename T> cudaError_t launch_mid_typed(nds_ctx *ctx, const DeviceCall &c, const MidPlan &p) {
    if (p.csize == 1) {
        auto kernel = mid_bitslice_kernel<T, false>;
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
        if (e != cudaSuccess) return e;
        const unsigned grid = (unsigned)std::min<unsigned long long>(c.geom.nlanes, (unsigned long long)ctx->sm_count);
        kernel<<<grid, MID_THREADS, p.smem, ctx->stream>>>((const T *)c.data, c.geom, c.q, (T *)c.out, p.nwords, p.bpc);
        ctx->launches += 1;
        return cudaGetLastError();
    }
    auto kernel = mid_bitslice_kernel<T, true>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return e;
    const unsigned long long groups = std::min<unsigned long long>(c.geom.nlanes, (unsigned long long)(ctx->sm_count / p.csize));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(groups * p.csize));
    cfg.blockDim = dim3(MID_THREADS);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)p.csize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kernel, (const T *)c.data, c.geom, c.q, (T *)c.out, p.nwords, p.bpc);
    ctx->launches += 1;
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace

bool mid_bitslice_supported(const nds_ctx *ctx, const DeviceCall &c) { return plan_mid(ctx, c).ok; }
int mid_bitslice_cluster(const nds_ctx *ctx, const DeviceCall &c) { return plan_mid(ctx, c).csize; }

cudaError_t launch_mid_bitslice(nds_ctx *ctx, const DeviceCall &c) {
    const MidPlan p = plan_mid(ctx, c);
    if (!p.ok) return cudaErrorNotSupported;
    ctx->last_path = "mid_bitslice";
    switch (c.dtype) {
    case NDS_F32: return launch_mid_typed<float>(ctx, c, p);
    case NDS_F64: return launch_mid_typed<double>(ctx, c, p);
    case NDS_I32: return launch_mid_typed<int32_t>(ctx, c, p);
    case NDS_U32: return launch_mid_typed<uint32_t>(ctx, c, p);
    case NDS_I64: return launch_mid_typed<int64_t>(ctx, c, p);
    case NDS_U64: return launch_mid_typed<uint64_t>(ctx, c, p);
    default: return cudaErrorNotSupported;
    }
}

}  // namespace nds
