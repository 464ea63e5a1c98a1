// One (or a few) very long contiguous lanes — BASELINE configs 4 and 5: "bracket select".
//
// An MSB radix select re-reads the whole lane once per digit; on a 16 GiB lane that caps the path at
// 1/8 of the HBM roofline, and uniform [0,1) doubles put half of the keys into ONE top-digit bucket, so
// the early digits filter nothing.  This path reads the lane ONCE:
//
//   1. sample      S1 = n/256 keys at hashed positions; S0 = 8192 of those, sorted by one CTA
//   2. bucket      count S1 against the sorted S0 (binary search): exact sample ranks of every S0 key
//   3. plan        per requested quantile pick a value bracket [lo, hi] of S0 keys whose sample-rank span
//                  is the expected position +- 5 sigma; merge overlapping brackets; build a two-level
//                  key -> class lookup table: 4096 top cells starting one cell BELOW the first bracket (keys
//                  below the origin are clamped onto it and read "below everything"; cells past the last
//                  bracket are "above everything"); a cell without a bracket edge carries its final entry,
//                  the cells that hold an edge are subdivided again
//   4. stream      THE pass over the lane: every element is classified (class = number of bracket lower
//                  edges <= key; one table read, a second, predicated one for the ~12 % of keys in a
//                  subdivided cell), counted in per-thread private 8-bit shared-memory counters (a plain
//                  LDS/ADD/STS on a conflict-free layout: shared atomics would cap the pass at 0.5
//                  element/cycle/SM), and the ~14 % of elements that fall inside a bracket are packed into
//                  a per-warp ring and appended, 32 at a time, to that bracket's candidate segment (this
//                  CTA's piece of it; the pieces of one CTA are contiguous)
//   5. resolve     exact counts below / inside every bracket (summed over ranks with ncclAllReduce when the
//                  lane is sharded) turn each requested rank into (bracket, rank inside the bracket); a rank
//                  that escaped its bracket raises FAIL and the caller falls back to the radix passes
//   6. refine      candidates only: 8-bit digits of (key - lo) with the same private counters, compaction of
//                  the chosen digit, until a segment holds <= 4096 keys, which one CTA sorts
//   7. output      Interpolate on device with the reference's formulas (nds_common.cuh)
//
// Every decision between the passes is taken on the device; the host only enqueues.
// Nothing here restates the reference's quickselect (src/sort.rs:133-283): order statistics are determined
// by the multiset, so the selected values are bit-identical.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "nds_bitslice.cuh"
#include "nds_internal.h"

namespace nds {
namespace {

constexpr int BRK_MAXB = 120;             // brackets per call
constexpr int BRK_NCLS = 128;             // counter classes of the streaming pass (127 = invalid: NaN)
constexpr int BRK_INVALID_CLS = 127;
constexpr int BRK_THREADS = 256;
constexpr int BRK_NC1 = 4096;             // top-level cells
constexpr int BRK_L2_PURE = 260;          // 2 * 128 pure (class, inside) combinations + 2 special entries
constexpr int BRK_L2_SUB = 16384;         // sub-cell entries
constexpr int BRK_L2N = BRK_L2_PURE + BRK_L2_SUB;
constexpr int BRK_S0 = 8192;
constexpr int BRK_MAXQ = 128;             // quantiles per call on this path
constexpr int BRK_MAXTASK = 2 * BRK_MAXQ; // unique ranks
constexpr int BRK_MAXSEG = 2 * BRK_MAXQ;
constexpr int BRK_FINAL_CAP = 4096;       // a segment this small is sorted by one CTA
constexpr int BRK_RCHUNK = 4096;          // refine passes: elements per chunk
constexpr int BRK_GATHER_CAP = 1 << 17;   // sharded lanes: keys of the small segments one rank may contribute per round
constexpr int BRK_MAX_RANKS = 8;
#ifndef NDS_STR_THREADS
#define NDS_STR_THREADS 1024
#endif
constexpr int STR_THREADS = NDS_STR_THREADS;  // streaming pass: 32 warps per SM (62 registers), the table-lookup chains are latency bound
constexpr int STR_WARPS = STR_THREADS / 32;
constexpr int STR_TILE = 4096;

constexpr unsigned ENT_COMPACT = 0x100u, ENT_MIXED = 0x200u;
// a table cell that holds exactly ONE bracket and that bracket is a single key v (a tie bracket b, the low byte): the
// class follows from two compares -- below v: gap b, equal: the tie's own class, above: gap b + 1 -- instead of the walk
// over the bracket edges.  Tie-heavy lanes send most of their keys through such cells.
constexpr unsigned ENT_TIE1 = 0x400u;
constexpr int BRK_LOWCARD_MAX = 63;       // distinct sample keys up to which a lane is counted per key (2 K + 1 classes)
// low-cardinality lanes: the streaming pass finds a key's class with ONE probe of a perfect hash table {key, class} in
// shared memory (4096 slots of 8 bytes for 4-byte keys, 2048 slots of 16 bytes for 8-byte keys: 32 KB either way)
template <typename Key> __host__ __device__ constexpr int brk_hash_bits() { return sizeof(Key) == 8 ? 11 : 12; }
template <typename Key> __device__ __forceinline__ unsigned brk_hash_fold(Key k) {
    if constexpr (sizeof(Key) == 8) return (unsigned)(k >> 32) ^ ((unsigned)k * 0x9E3779B1u);
    else return (unsigned)k;
}

enum : int {
    H_NTOTAL = 0, H_NINVALID, H_NEFF, H_FAIL, H_B, H_NSEG, H_NTASK, H_NACTIVE, H_S1VALID, H_EMPTY, H_NCHUNK_COUNT,
    H_NCHUNK_COMPACT, H_NCLS_USED, H_NTIE, H_CUR, H_NEXT_NCHUNK, H_NEXT_NSEG, H_OLD_NSEG, H_NW, H_ROUND, H_DID, H_FINALCAP, H_NFINAL, H_MULTI, H_LOCALFAIL, H_PSTRIDE,
    H_S1,           // keys this rank samples (sharded lanes: computed on the device from the total length)
    H_GATE_ROUND,   // 1 while refinement rounds have work (identical on every rank: derived from reduced counts)
    H_GATE_SMALL,   // sharded lanes: 1 when small segments wait to be gathered and sorted
    H_NEXT_ACTIVE, H_XWORDS, H_GWORDS, H_WORDS = 40
};
enum : int { FAIL_ESCAPED = 1, FAIL_CAPACITY = 2, FAIL_TOO_MANY = 4, FAIL_INCONSISTENT = 8, FAIL_REMOTE = 16 };
enum : int { TASK_ACTIVE = 0, TASK_DONE = 1 };
enum : int { SEG_TIE = 1, SEG_FINAL = 2, SEG_USED = 4, SEG_DONE = 8 };

template <typename Key> struct BrkSeg {
    Key lo, hi;
    int shift, flags;
    unsigned long long off, cnt_local, cnt_global, cursor;
    unsigned long long piece_cap;  // round 0: the segment is H_NW pieces of this capacity (one per streaming CTA), fills in buf.fill
    unsigned long long piece_stride;  // round 0: distance between the pieces of consecutive CTAs (all pieces of ONE CTA are contiguous)
};
struct BrkTask {
    unsigned long long r;      // rank inside its segment (over all ranks when sharded)
    unsigned long long value;  // the key, once known
    int seg, state;
};
template <typename Key> struct BrkLutConst {
    Key kmin;          // origin of the lookup cells (64-bit keys in hi32 mode: a multiple of 2^32)
    Key kmin_true;     // lower edge of the first bracket
    int shift1, subshift;
    unsigned submask;
    int B, ncls_used;
    int hi32;          // 64-bit keys: the tables are indexed with the high word only (shift1, subshift >= 32)
    int count_only;    // low-cardinality lane: every bracket is one distinct key, nothing is compacted (tie class = B + 1 + b)
    int hashed;        // count_only: `hmul` maps the distinct keys one-to-one onto the slots of a small hash table
    unsigned hmul;     // slot = (fold(key) * hmul) >> (32 - BRK_HASH_BITS)
};

template <typename Key> struct BrkBuf {
    unsigned long long *hdr;         // [H_WORDS]
    unsigned long long *red_n;       // [2]  local element count -> total            (allreduce 0)
    unsigned long long *s0hist;      // [2 (S0 + 2)] bucket counts of S1 against S0, then #{== S0[i]}        (allreduce 1)
    unsigned long long *red_stream;  // [NCLS + MAXB] class counts, bracket in-counts  (allreduce 2)
    unsigned long long *cur_local;   // [MAXB] local fill of the bracket segments (sum of the pieces)
    unsigned *fill;                  // [MAXB * nw_cap] fill of every (bracket, streaming CTA) piece
    unsigned long long *red_refine;  // [MAXSEG * 256]                               (allreduce per refine round)
    unsigned long long *loc_refine;  // [MAXSEG * 256] local copy (== red_refine on one GPU)
    Key *s0, *blo, *bhi;
    Key *s0all;                      // sharded lanes: the all-gathered S0
    unsigned long long *xsend, *xrecv;  // sharded lanes, first exchange: [n_local, 0, this rank's share of the S0 keys] / of all ranks
    unsigned long long *red_refine_fail;  // sharded lanes: the fail word that travels in front of red_refine
    unsigned long long *red_end;     // [2] final agreement on FAIL over ranks
    unsigned char *gsend, *grecv;    // sharded lanes: packed small segments of this rank / of all ranks (header + keys)
    unsigned long long *boff, *bcap;
    unsigned char *btie;             // tie class of a bracket (lo == hi: counted, never compacted) or 0xff
    unsigned *lut1;                  // [NC1]
    unsigned short *lut2;            // [L2N]
    BrkLutConst<Key> *lutc;
    BrkSeg<Key> *seg[2];
    BrkTask *task;
    int *slot_task;                  // [2 * nq]
    unsigned long long *slot_rank;   // [2 * nq]
    short *newid;                    // [MAXSEG * 256]
    unsigned long long *cprefix_count, *cprefix_compact;  // [MAXSEG + 1] chunk prefix sums
    Key *cand[2];
    unsigned long long cand_cap[2];
    unsigned s1cap, nw_cap;
};

__device__ __forceinline__ unsigned long long ld_now(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}
// A lane sharded over ranks must take every collective on every rank, so a failure seen by ONE rank must not change
// that rank's control flow on its own: it is parked in H_LOCALFAIL (which only mutes the data kernels), travels with the
// next collective, and comes back to every rank as FAIL_REMOTE in H_FAIL -- the word the host and the deciding kernels
// look at.  On one GPU the two are the same thing.
__device__ __forceinline__ void raise_fail(unsigned long long *hdr, int bits) {
    const bool park = hdr[H_MULTI] != 0 && !(bits & FAIL_REMOTE);
    atomicOr(&hdr[park ? H_LOCALFAIL : H_FAIL], (unsigned long long)bits);
}
__device__ __forceinline__ bool failed_any(const unsigned long long *hdr) { return (hdr[H_FAIL] | hdr[H_LOCALFAIL]) != 0; }
__device__ __forceinline__ unsigned mix32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
template <typename Key> __device__ __forceinline__ int key_bits_used(Key x) {
    if (sizeof(Key) == 8) return x ? 64 - __clzll((long long)x) : 0;
    return x ? 32 - __clz((int)x) : 0;
}
// conflict-free slot of a thread inside one class row of the private 16-bit counters: the 32 lanes of a
// warp land in 32 different banks, two warps share a 128-byte row (low / high half-word)
__device__ __forceinline__ unsigned counter_slot() {
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    return (warp >> 1) * 64 + lane * 2 + (warp & 1);
}

// floats: keys of NaNs lie outside [key(-inf), key(+inf)]
template <typename T> __device__ __forceinline__ bool key_is_nan(typename Traits<T>::Key k) {
    using Key = typename Traits<T>::Key;
    if (!Traits<T>::is_float) return false;
    if (sizeof(Key) == 8) return k > (Key)0xFFF0000000000000ull || k < (Key)0x000FFFFFFFFFFFFFull;
    return k > (Key)0xFF800000u || k < (Key)0x007FFFFFu;
}

// ------------------------------------------------------------------------------------------------
// 1. sample
// ------------------------------------------------------------------------------------------------
// sample size for a lane of n elements: `need` keys (what the requested quantiles ask for), between 2^16 and
// min(n / 8, 2^24), never more than the lane holds
__host__ __device__ inline unsigned pick_s1_from_need(unsigned long long n, double need) {
    unsigned long long hi = n / 8;
    if (hi < (1ull << 16)) hi = 1ull << 16;
    if (hi > (1ull << 24)) hi = 1ull << 24;
    unsigned long long s1 = need < 65536.0 ? 65536ull : (need > (double)hi ? hi : (unsigned long long)need);
    if (s1 > n) s1 = n;
    return (unsigned)s1;
}

// 0. this rank's element count (summed over ranks by the caller when the lane is sharded); final-sort threshold
template <typename Key>
__global__ void brk_begin_kernel(BrkBuf<Key> buf, unsigned long long n_local, unsigned final_cap, int multi, unsigned s1) {
    if (threadIdx.x == 0) {
        buf.red_n[0] = n_local;
        buf.hdr[H_FINALCAP] = final_cap;
        buf.hdr[H_MULTI] = multi ? 1ull : 0ull;
        buf.hdr[H_S1] = s1;  // (sharded lanes: overwritten by the sample kernel once the total length is known)
    }
}
// sharded lanes, first exchange: this rank's element count and its share of the S0 keys (drawn at hashed, evenly
// spread positions of the local shard), packed as [n_local, 0, keys ...]; the all-gathered keys are sorted identically
// on every rank and the counts add up to the total length -- ONE exchange step for both
template <typename T>
__global__ void brk_pick_local_s0_kernel(const T *__restrict__ data, unsigned long long n, BrkBuf<typename Traits<T>::Key> buf,
                                         unsigned count) {
    using Tr = Traits<T>;
    using Key = typename Tr::Key;
    Key *dst = reinterpret_cast<Key *>(buf.xsend + 2);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        buf.xsend[0] = n;
        buf.xsend[1] = 0;
    }
    const unsigned long long q = n / count, r = n % count;
    const double inv = 1.0 / (double)count;
    const unsigned long long width = q + (r ? 1ull : 0ull);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        Key k = (Key) ~(Key)0;
        if (n) {
            const unsigned long long p0 = (unsigned long long)i * q + (unsigned long long)((double)((unsigned long long)i * r) * inv);
            unsigned long long pos = p0 + (((unsigned long long)mix32(i ^ 0x5bd1e995u) * width) >> 32);
            if (pos >= n) pos = n - 1;
            k = Tr::to_key(data[pos]);
            if (key_is_nan<T>(k)) k = (Key) ~(Key)0;
        }
        dst[i] = k;
    }
}
template <typename Key> __global__ void brk_unpack_s0_kernel(BrkBuf<Key> buf, int nranks, unsigned count) {
    const size_t words_per_rank = 2 + ((size_t)count * sizeof(Key) + 7) / 8;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long total = 0;
        for (int r = 0; r < nranks; ++r) total += buf.xrecv[(size_t)r * words_per_rank];
        buf.red_n[0] = total;
    }
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)nranks * count; i += gridDim.x * blockDim.x) {
        const unsigned r = i / count, j = i - r * count;
        buf.s0all[i] = reinterpret_cast<const Key *>(buf.xrecv + (size_t)r * words_per_rank + 2)[j];
    }
}
// local fail bits travel with the counts of every exchange step, so that all ranks give up together
template <typename Key> __global__ void brk_merge_fail_kernel(BrkBuf<Key> buf, const unsigned long long *word) {
    if (threadIdx.x == 0 && *word) raise_fail(buf.hdr, FAIL_REMOTE);
}
template <typename Key> __global__ void brk_post_fail_kernel(BrkBuf<Key> buf, unsigned long long *word) {
    if (threadIdx.x == 0) *word = failed_any(buf.hdr) ? 1ull : 0ull;
}
template <typename Key> __global__ void brk_copy_counts_kernel(BrkBuf<Key> buf) {  // [fail word, red_refine = loc_refine]
    if (!buf.hdr[H_GATE_ROUND]) return;
    const int nseg = (int)buf.hdr[H_NSEG];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nseg * 256; i += gridDim.x * blockDim.x)
        buf.red_refine[i] = buf.loc_refine[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        buf.red_refine_fail[0] = failed_any(buf.hdr) ? 1ull : 0ull;
        buf.hdr[H_XWORDS] = 1ull + (unsigned long long)nseg * 256ull;
    }
}

// 2a. S0 = 8192 keys of the lane (brk_pick_local_s0_kernel; over ranks: the gathered shares), sorted by counting: the place of key i is
// #{ j : S0[j] < S0[i], or S0[j] == S0[i] and j < i }.  8192^2 comparisons spread over 128 CTAs (64 keys each, eight
// warps that each scan an eighth of S0 from shared memory with broadcast reads) take ~15 us; the bitonic network that
// one CTA ran before took 100 us -- a fixed cost that a lane sharded over eight GPUs cannot amortise.
constexpr int SORT_KEYS_PER_CTA = 64;
constexpr int SORT_CTAS = BRK_S0 / SORT_KEYS_PER_CTA;
constexpr int SORT_THREADS = 256;
static_assert(BRK_S0 % (SORT_THREADS / 32) == 0 && SORT_KEYS_PER_CTA == 64, "sort geometry");

template <typename Key>
__global__ void __launch_bounds__(SORT_THREADS) brk_sort_s0_kernel(BrkBuf<Key> buf, const Key *gathered) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Key *s = reinterpret_cast<Key *>(smem_raw);
    __shared__ unsigned place[SORT_KEYS_PER_CTA];
    for (unsigned i = threadIdx.x; i < BRK_S0; i += blockDim.x) s[i] = gathered[i];   // (an empty lane: all ~0)
    if (threadIdx.x < SORT_KEYS_PER_CTA) place[threadIdx.x] = 0;
    __syncthreads();
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned i0 = blockIdx.x * SORT_KEYS_PER_CTA + lane, i1 = i0 + 32;
    const Key a = s[i0], b = s[i1];
    constexpr unsigned SPAN = BRK_S0 / (SORT_THREADS / 32);
    const unsigned j0 = warp * SPAN;
    unsigned ca = 0, cb = 0;
#pragma unroll 8
    for (unsigned j = j0; j < j0 + SPAN; ++j) {
        const Key k = s[j];  // the same address for the whole warp: a broadcast
        ca += (k < a || (k == a && j < i0)) ? 1u : 0u;
        cb += (k < b || (k == b && j < i1)) ? 1u : 0u;
    }
    atomicAdd(&place[lane], ca);
    atomicAdd(&place[lane + 32], cb);
    __syncthreads();
    if (threadIdx.x < SORT_KEYS_PER_CTA) buf.s0[place[threadIdx.x]] = s[blockIdx.x * SORT_KEYS_PER_CTA + threadIdx.x];
}

// 2b. the sample, ranked against S0
constexpr int BRK_BLUT = 4096;  // cells of the search table over S0

// Sample and bucket in ONE kernel: every sampled key is ranked against the sorted splitters S0 (picked from the lane
// beforehand, brk_pick_local_s0_kernel) while the other scattered reads of the thread are still on their way -- the sample is
// never written out and read back, and the histogram atomics hide behind the read latency (C4: 287 us; 256 + 184 us as two
// kernels).  hist[i] = #{ sampled keys k : S0[i-1] < k <= S0[i] }, then #{ k == S0[i] } for the first key of every run.
// `sharded`: the lane is this rank's shard of a longer one whose total length red_n[0] is only known on the device; every
// rank samples at the same rate, so the sample size follows from the shard's share (and is published in H_S1).
// The strata are spread EVENLY over the lane, one hashed position in each: stratum i starts at floor(i n / s1) = i q +
// floor(i r / s1) (q = n / s1, r = n % s1) -- a sorted lane sampled at two different rates puts its quantiles in the wrong
// place.  The second term comes from one double multiplication (i r < 2^48 is exact; being off by one element does not
// matter), the offset inside the stratum from a multiply-high instead of a modulo (strata are narrower than 2^32).
template <typename T>
__global__ void __launch_bounds__(512) brk_sample_bucket_kernel(const T *__restrict__ data, unsigned long long n, unsigned s1,
                                                                 double need, int sharded, BrkBuf<typename Traits<T>::Key> buf) {
    using Tr = Traits<T>;
    using Key = typename Tr::Key;
    constexpr bool IS_FLOAT = Tr::is_float;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (sharded) {
        const unsigned long long n_total = buf.red_n[0];
        const unsigned s1_total = pick_s1_from_need(n_total, need);
        const double share = n_total ? (double)n / (double)n_total : 0.0;
        unsigned long long want = (unsigned long long)ceil(share * (double)s1_total);
        if (want > n) want = n;
        if (want > buf.s1cap) want = buf.s1cap;
        s1 = (unsigned)want;
        if (blockIdx.x == 0 && threadIdx.x == 0) buf.hdr[H_S1] = s1;
    }
    if (s1 == 0) return;
    Key *s = reinterpret_cast<Key *>(smem_raw);
    unsigned short *start = reinterpret_cast<unsigned short *>(s + BRK_S0);  // [BLUT + 1]: lower_bound of every cell start
    for (unsigned i = threadIdx.x; i < BRK_S0; i += blockDim.x) s[i] = buf.s0[i];
    __syncthreads();
    auto lower_bound = [&](Key k, unsigned lo, unsigned hi) -> unsigned {  // #{ S0[lo..hi) < k } + lo
        while (lo < hi) {
            const unsigned mid = (lo + hi) >> 1;
            if (s[mid] < k) lo = mid + 1;
            else hi = mid;
        }
        return lo;
    };
    // a key-linear table over [S0[0], S0[last]] narrows the 13-step binary search to the few splitters of one cell
    const Key k0 = s[0], k1 = s[BRK_S0 - 1];
    const Key range = k1 - k0;
    int sh = key_bits_used(range) - 12;
    if (sh < 0) sh = 0;
    while ((range >> sh) >= (Key)BRK_BLUT) ++sh;
    for (unsigned c = threadIdx.x; c <= BRK_BLUT; c += blockDim.x) {
        const Key off = (Key)c << sh;
        const bool past = c && (off >> sh) != (Key)c;  // shifted out of the key type
        start[c] = (unsigned short)((past || off > range) ? BRK_S0 : lower_bound((Key)(k0 + off), 0, BRK_S0));
    }
    __syncthreads();
    const unsigned long long q = n / s1, r = n % s1;
    const double inv_s1 = 1.0 / (double)s1;
    const unsigned long long width = q + (r ? 1ull : 0ull);
    const unsigned step = gridDim.x * blockDim.x;
    auto position = [&](unsigned i) -> unsigned long long {
        const unsigned long long p0 = (unsigned long long)i * q + (unsigned long long)((double)((unsigned long long)i * r) * inv_s1);
        unsigned long long pos = p0 + (((unsigned long long)mix32(i) * width) >> 32);
        return pos >= n ? n - 1 : pos;
    };
    constexpr int MLP = 8;   // independent scattered reads in flight per thread
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < s1; i += (unsigned long long)MLP * step) {
        T x[MLP];
#pragma unroll
        for (int u = 0; u < MLP; ++u) {
            const unsigned long long iu = i + (unsigned long long)u * step;
            if (iu < s1) x[u] = data[position((unsigned)iu)];
        }
#pragma unroll
        for (int u = 0; u < MLP; ++u) {
            const unsigned long long iu = i + (unsigned long long)u * step;
            if (iu >= s1) continue;
            const Key k = Tr::to_key(x[u]);
            if (IS_FLOAT && key_is_nan<T>(k)) continue;   // NaN samples are not counted
            unsigned lo;
            if (k <= k0) lo = 0;
            else if (k > k1) lo = BRK_S0;
            else {
                const unsigned c = (unsigned)((Key)(k - k0) >> sh);  // < BLUT; cell_start(c) <= k < cell_start(c + 1)
                lo = lower_bound(k, start[c], start[c + 1]);
            }
            atomicAdd(&buf.s0hist[lo], 1ull);
            if (lo < BRK_S0 && s[lo] == k) atomicAdd(&buf.s0hist[BRK_S0 + 2 + lo], 1ull);  // #{ sample == S0[lo] }, first of its run
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 3. plan: brackets + lookup tables
// ------------------------------------------------------------------------------------------------
template <typename Key> struct PointClass { int j; bool in; };
template <typename Key>
__device__ __forceinline__ PointClass<Key> classify_point(const Key *blo, const Key *bhi, int B, Key k) {
    int lo = 0, hi = B;  // j = #{ blo <= k }
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (blo[mid] <= k) lo = mid + 1;
        else hi = mid;
    }
    PointClass<Key> p;
    p.j = lo;
    p.in = lo > 0 && k <= bhi[lo - 1];
    return p;
}
__device__ __forceinline__ unsigned pure_entry(int j, bool in, const unsigned char *btie) {
    if (in) {
        const unsigned char t = btie[j - 1];
        if (t != 0xff) return (unsigned)t;     // tie bracket: own class, not compacted
        return (unsigned)j | ENT_COMPACT;
    }
    return (unsigned)j;
}

template <typename Key, bool IS_FLOAT>
__global__ void __launch_bounds__(1024)
brk_plan_kernel(BrkBuf<Key> buf, QuantArgs qa, double csigma, unsigned long long n_local, int sharded, unsigned nw,
                int lowcard_allowed) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned *cum = reinterpret_cast<unsigned *>(smem_raw);            // [S0 + 1] inclusive: #{ s1 <= S0[i] }
    Key *s0 = reinterpret_cast<Key *>(smem_raw + ((BRK_S0 + 1) * 4 + 15) / 16 * 16);  // [S0] the sorted sample
    unsigned short *mixcell = reinterpret_cast<unsigned short *>(s0 + BRK_S0);          // [NC1] cells that hold an edge
    unsigned *eqc = reinterpret_cast<unsigned *>(mixcell + BRK_NC1);                    // [S0 + 2] #{ s1 == S0[i] } (first of its run)
    __shared__ unsigned long long scap[BRK_MAXB];
    __shared__ Key qlo[BRK_MAXQ], qhi[BRK_MAXQ], slo[BRK_MAXQ], shi[BRK_MAXQ];
    __shared__ Key blo[BRK_MAXB], bhi[BRK_MAXB];
    __shared__ unsigned char btie[BRK_MAXB];
    __shared__ unsigned char qheavy[BRK_MAXQ];
    __shared__ Key hv[2 * BRK_MAXQ];
    __shared__ unsigned warp_tot[32];
    __shared__ int sB, sM, sShift1, sNmixed, sFail, sHi32, sLow, sK;
    __shared__ Key sKminLut;
    __shared__ Key dkey[BRK_LOWCARD_MAX + 1];
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const Key KMAX = (Key) ~(Key)0;
    const int nq = qa.nq;

    // ---- inclusive scan of the bucket counts (S0 + 1 entries, 1024 threads x 9) ----------------------
    {
        constexpr int PER = (BRK_S0 + 1 + 1023) / 1024;
        unsigned v[PER], sum = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int idx = tid * PER + i;
            v[i] = idx <= BRK_S0 ? (unsigned)buf.s0hist[idx] : 0u;
            sum += v[i];
        }
        unsigned incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
            if ((tid & 31) >= d) incl += t;
        }
        if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) {
            unsigned w = warp_tot[tid], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned t = __shfl_up_sync(0xffffffffu, wi, d);
                if (tid >= d) wi += t;
            }
            warp_tot[tid] = wi - w;
        }
        __syncthreads();
        unsigned run = warp_tot[tid >> 5] + incl - sum;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int idx = tid * PER + i;
            run += v[i];
            if (idx <= BRK_S0) cum[idx] = run;
        }
        if (tid == 0) sFail = 0;
        for (int i = tid; i < BRK_S0; i += nthreads) s0[i] = buf.s0[i];
        for (int i = tid; i < BRK_S0 + 2; i += nthreads) eqc[i] = (unsigned)buf.s0hist[BRK_S0 + 2 + i];
        __syncthreads();
    }
    const unsigned s1v = cum[BRK_S0];  // valid samples (over all ranks when sharded)
    const unsigned long long n_total = buf.red_n[0];
    if (tid == 0) {
        buf.hdr[H_NTOTAL] = n_total;
    }
    if (s1v == 0 || n_total == 0) {  // nothing but NaNs in the sample (or nothing at all): no brackets
        if (tid == 0) {
            buf.hdr[H_B] = 0;
            buf.lutc->B = 0;
            buf.lutc->kmin = 0;
            buf.lutc->kmin_true = 0;
            buf.lutc->hi32 = 0;
            buf.hdr[H_NW] = nw;
            buf.lutc->shift1 = 0;
            buf.lutc->subshift = 0;
            buf.lutc->submask = 0;
            buf.lutc->ncls_used = 1;
            buf.lutc->count_only = 0;
            buf.lutc->hashed = 0;
            buf.lutc->hmul = 1u;
        }
        for (int c = tid; c < BRK_NC1; c += nthreads) buf.lut1[c] = 0u;  // every number is class 0 (NaNs: invalid class)
        for (int i = tid; i < BRK_L2_PURE; i += nthreads) buf.lut2[i] = (unsigned short)0;
        return;
    }

    // ---- low-cardinality lanes: when the sorted sample S0 holds at most 63 distinct keys, every one of them becomes
    //      a bracket of its own (a tie: counted in its own class, never compacted) and the streaming pass only COUNTS.
    //      Any requested rank then resolves from the counts; a rank that falls between the known keys (a key the
    //      sample has not seen) escapes and the call falls back to the radix passes, as always. -------------------
    if (tid == 0) { sK = 0; sLow = 0; }
    __syncthreads();
    for (int i = tid; i < BRK_S0; i += nthreads) {
        const Key v = s0[i];
        if (IS_FLOAT && v == KMAX) continue;  // NaN samples
        if (i == 0 || s0[i - 1] != v) {
            const int slot = atomicAdd(&sK, 1);
            if (slot <= BRK_LOWCARD_MAX) dkey[slot] = v;
        }
    }
    __syncthreads();
    if (tid == 0) sLow = (sK >= 1 && sK <= BRK_LOWCARD_MAX && lowcard_allowed) ? 1 : 0;
    __syncthreads();
    const bool low = sLow != 0;
    if (low) {  // the distinct keys, ascending (rank sort of <= 63 keys)
        const int K = sK;
        if (tid < K) {
            const Key v = dkey[tid];
            int r = 0;
            for (int u = 0; u < K; ++u) r += dkey[u] < v ? 1 : 0;
            blo[r] = v;
            bhi[r] = v;
        }
    }
    // ---- otherwise: one bracket per requested quantile -------------------------------------------------
    for (int t = tid; t < (low ? 0 : nq); t += nthreads) {
        double p;
        if (qa.ranks) p = n_total > 1 ? (double)qa.ranks[t] / (double)(n_total - 1) : 0.0;
        else p = qa.inline_q ? qa.qv[t] : qa.qs[t];
        p = fmin(fmax(p, 0.0), 1.0);
        const double P = p * (double)(s1v - 1);
        const double sd = sqrt(fmax(p * (1.0 - p) * (double)s1v, 1.0));
        const double Plo = floor(P - csigma * sd - 2.0), Phi = ceil(P + csigma * sd + 2.0) + 1.0;
        Key lo = 0, hi = KMAX;
        if (Plo > 0.0) {  // largest i with #{ s1 < S0[i] } <= Plo   (strict: ties must not widen the bracket)
            const unsigned target = (unsigned)fmin(Plo, 4294967295.0);
            int a = 0, b = BRK_S0;  // first index with cum_lt > target
            while (a < b) {
                const int mid = (a + b) >> 1;
                const Key v = s0[mid];
                int f0 = 0, f1 = mid;  // first index of the run of keys equal to S0[mid]
                while (f0 < f1) {
                    const int m2 = (f0 + f1) >> 1;
                    if (s0[m2] < v) f0 = m2 + 1;
                    else f1 = m2;
                }
                const unsigned cum_lt = cum[f0] - eqc[f0];
                if (cum_lt <= target) a = mid + 1;
                else b = mid;
            }
            if (a > 0) lo = s0[a - 1];
        }
        if (Phi + 1.0 <= (double)cum[BRK_S0 - 1]) {  // smallest i with cum[i] >= Phi + 1
            const unsigned target = (unsigned)(Phi + 1.0);
            int a = 0, b = BRK_S0;
            while (a < b) {
                int mid = (a + b) >> 1;
                if (cum[mid] < target) a = mid + 1;
                else b = mid;
            }
            if (a < BRK_S0) hi = s0[a];
        }
        if (IS_FLOAT) {  // numbers live in [key(-inf), key(+inf)]; the keys outside are NaNs
            const Key ninf = sizeof(Key) == 8 ? (Key)0x000FFFFFFFFFFFFFull : (Key)0x007FFFFFu;
            const Key pinf = sizeof(Key) == 8 ? (Key)0xFFF0000000000000ull : (Key)0xFF800000u;
            if (lo < ninf) lo = ninf;
            if (hi > pinf) hi = pinf;
        }
        if (hi < lo) hi = lo;
        qlo[t] = lo;
        qhi[t] = hi;
        // an edge key that repeats in the sample is a heavy tie: it becomes a bracket of its own (counted, never
        // compacted) so that the candidates exclude it
        auto sample_multiplicity = [&](Key v) -> unsigned {
            int f0 = 0, f1 = BRK_S0;
            while (f0 < f1) {
                const int m2 = (f0 + f1) >> 1;
                if (s0[m2] < v) f0 = m2 + 1;
                else f1 = m2;
            }
            return (f0 < BRK_S0 && s0[f0] == v) ? eqc[f0] : 0u;
        };
        qheavy[t] = (unsigned char)((sample_multiplicity(lo) >= 4 ? 1 : 0) | (sample_multiplicity(hi) >= 4 ? 2 : 0));
    }
    __syncthreads();
    // sort by (lo, hi) with a rank sort (nq <= 128), then merge overlapping brackets sequentially
    for (int t = tid; t < (low ? 0 : nq); t += nthreads) {
        int rank = 0;
        const Key l = qlo[t], h = qhi[t];
        for (int u = 0; u < nq; ++u) {
            const Key lu = qlo[u], hu = qhi[u];
            if (lu < l || (lu == l && (hu < h || (hu == h && u < t)))) ++rank;
        }
        slo[rank] = l;
        shi[rank] = h;
    }
    __syncthreads();
    if (tid == 0) {
        // heavy edge keys, sorted and unique
        int nh = 0;
        for (int t = 0; t < (low ? 0 : nq); ++t) {
            for (int side = 0; side < 2; ++side) {
                if (!(qheavy[t] & (1 << side))) continue;
                const Key v = side ? qhi[t] : qlo[t];
                int pos = nh;
                bool dup = false;
                for (int i = 0; i < nh; ++i) {
                    if (hv[i] == v) { dup = true; break; }
                    if (hv[i] > v) { pos = i; break; }
                }
                if (dup) continue;
                for (int i = nh; i > pos; --i) hv[i] = hv[i - 1];
                hv[pos] = v;
                ++nh;
            }
        }
        // merge overlapping / adjacent raw intervals into spans, then cut every span at the heavy keys inside it
        int B = 0, ih = 0;
        auto emit = [&](Key l, Key h) {
            if (B < BRK_MAXB) { blo[B] = l; bhi[B] = h; }
            ++B;
        };
        auto emit_span = [&](Key l, Key h) {
            while (ih < nh && hv[ih] < l) ++ih;
            Key cur = l;
            bool open = true;  // [cur, h] still to emit
            while (open && ih < nh && hv[ih] <= h) {
                const Key v = hv[ih++];
                if (v > cur) emit(cur, v - 1);
                emit(v, v);
                if (v == h) open = false;
                else cur = v + 1;
            }
            if (open) emit(cur, h);
        };
        if (low) {
            B = sK;  // blo / bhi hold the distinct keys already
        } else {
            Key cl = slo[0], ch = shi[0];
            for (int t = 1; t < nq; ++t) {
                if (slo[t] <= ch || (ch != KMAX && slo[t] == ch + 1)) {
                    if (shi[t] > ch) ch = shi[t];
                } else {
                    emit_span(cl, ch);
                    cl = slo[t];
                    ch = shi[t];
                }
            }
            emit_span(cl, ch);
        }
        int fail = 0;
        if (B > BRK_MAXB) { fail |= FAIL_TOO_MANY; B = BRK_MAXB; }
        // tie brackets (a single key) get their own counter class while classes last
        int ntie = 0;
        for (int b = 0; b < B; ++b) {
            btie[b] = 0xff;
            if (blo[b] == bhi[b] && B + 1 + ntie < BRK_INVALID_CLS) btie[b] = (unsigned char)(B + 1 + ntie++);
        }
        if (B + 1 + ntie > BRK_INVALID_CLS) fail |= FAIL_TOO_MANY;
        sB = B;
        // top-level cell width.  64-bit keys whose brackets span many high words are indexed by the high word only
        // (half the integer work per key); the cell origin is then rounded down to a multiple of 2^32 so that
        // (k - kmin) >> 32 is exactly hi(k) - hi(kmin).
        Key kmin_lut = blo[0];
        int hi32 = 0;
        if (sizeof(Key) == 8 && key_bits_used((Key)(bhi[B - 1] - blo[0])) >= 44) {
            hi32 = 1;
            kmin_lut = (Key)((unsigned long long)blo[0] & 0xFFFFFFFF00000000ull);
        }
        // The origin of the cells lies one whole cell BELOW the first bracket: the streaming pass clamps every key
        // below the origin onto it, i.e. into sub-cell 0 of cell 0, which then holds no key of a bracket (or, when
        // the origin is clipped at key 0, no key below the origin exists).  Cells past the last bracket are pure
        // "above"; the last cell takes everything beyond the table.
        const Key kbase = kmin_lut;
        int shift1 = key_bits_used((Key)(bhi[B - 1] - kbase)) - 12;
        if (shift1 < 0) shift1 = 0;
        if (hi32 && shift1 < 32) shift1 = 32;
        for (;; ++shift1) {
            const Key step = (Key)1 << shift1;
            kmin_lut = kbase >= step ? kbase - step : (Key)0;
            if (((Key)(bhi[B - 1] - kmin_lut) >> shift1) <= (Key)(BRK_NC1 - 2)) break;
        }
        sShift1 = shift1;
        sHi32 = hi32;
        sKminLut = kmin_lut;
        buf.hdr[H_NW] = nw;
        sNmixed = 0;
        sFail = fail;
        buf.hdr[H_B] = (unsigned long long)B;
        buf.hdr[H_NTIE] = (unsigned long long)ntie;
        buf.hdr[H_NCLS_USED] = (unsigned long long)(B + 1 + ntie);
        if (fail) raise_fail(buf.hdr, fail);
    }
    __syncthreads();
    const int B = sB, shift1 = sShift1;
    // ---- low-cardinality lanes: a multiplier that hashes the distinct keys onto different slots (and, among those,
    //      one that spreads them over the shared-memory banks).  One candidate per warp and turn: lane i holds keys i and
    //      i + 32, duplicates show up in match.any and in one round of shuffles. -----------------------------------
    __shared__ unsigned hbest[32];
    if (low && !sFail) {
        constexpr int HB = brk_hash_bits<Key>();
        constexpr unsigned GROUPS = sizeof(Key) == 8 ? 8u : 16u;   // table entries per 128 bytes of shared memory
        const int lane = tid & 31, warp = tid >> 5;
        const bool va = lane < B, vb = lane + 32 < B;
        const unsigned fa = va ? brk_hash_fold<Key>(blo[lane]) : 0u, fb = vb ? brk_hash_fold<Key>(blo[lane + 32]) : 0u;
        unsigned best = 0xffffffffu;  // (score << 24 | candidate): smaller is better
        for (int turn = 0; turn < 16; ++turn) {
            const unsigned cand = (unsigned)(turn * 32 + warp);
            const unsigned mul = mix32(cand * 2654435761u + 12345u) | 1u;
            const unsigned sa = va ? (fa * mul) >> (32 - HB) : 0xffff0000u + (unsigned)lane;   // (idle lanes: unique dummies)
            const unsigned sb = vb ? (fb * mul) >> (32 - HB) : 0xfffe0000u + (unsigned)lane;
            const unsigned pa = __match_any_sync(0xffffffffu, sa), pb = __match_any_sync(0xffffffffu, sb);
            bool dup = (pa & (pa - 1)) != 0 || (pb & (pb - 1)) != 0;
            unsigned load = __popc(__match_any_sync(0xffffffffu, va ? (sa & (GROUPS - 1)) : 0x100u + (unsigned)lane));
            for (int j = 0; j < 32; ++j) {
                const unsigned o = __shfl_sync(0xffffffffu, sb, j);
                dup = dup || o == sa;
                load += (va && o < 0xfffe0000u && ((o ^ sa) & (GROUPS - 1)) == 0) ? 1u : 0u;
            }
            const bool bad = __any_sync(0xffffffffu, dup);
            const unsigned score = __reduce_max_sync(0xffffffffu, load);
            if (!bad) best = min(best, (score << 24) | cand);
        }
        if (lane == 0) hbest[warp] = best;
    }
    __syncthreads();
    if (tid == 0) {
        unsigned best = 0xffffffffu;
        if (low && !sFail)
            for (int w = 0; w < 32; ++w) best = min(best, hbest[w]);
        buf.lutc->hashed = best != 0xffffffffu ? 1 : 0;
        buf.lutc->hmul = mix32((best & 0xffffffu) * 2654435761u + 12345u) | 1u;
    }
    // candidate segments: predicted size from the sample, with head room (one bracket per thread)
    {
        const double scale = (double)n_total / (double)s1v;
        for (int b = tid; b < B; b += nthreads) {
            unsigned long long cap = 0;
            if (btie[b] == 0xff) {
                // upper bound of #{ s1 in [lo, hi] } from the S0 ranks
                int a = 0, e = BRK_S0;  // upper_bound(hi)
                while (a < e) { int mid = (a + e) >> 1; if (s0[mid] <= bhi[b]) a = mid + 1; else e = mid; }
                unsigned all_hi;
                if (a > 0 && s0[a - 1] == bhi[b]) all_hi = cum[a - 1];   // hi is an S0 key: exact
                else all_hi = cum[a] - (a < BRK_S0 ? eqc[a] : 0u);  // #{ s1 < next S0 key }, or everything
                a = 0; e = BRK_S0;      // lower_bound(lo)
                while (a < e) { int mid = (a + e) >> 1; if (s0[mid] < blo[b]) a = mid + 1; else e = mid; }
                const unsigned lt_lo = a > 0 ? cum[a - 1] : 0u;
                const double cnt = (double)(all_hi > lt_lo ? all_hi - lt_lo : 0u) + 1.0;
                double want = (cnt + 6.0 * sqrt(cnt) + 16.0) * scale * 1.05 + 2048.0;
                if (sharded) want = want * ((double)n_local / (double)n_total) * 1.5 + 4096.0;  // this rank's share
                if (want > (double)n_local) want = (double)n_local;
                // every streaming CTA owns one piece of the segment: its share, 6 sigma of head room
                const double share = want / (double)nw;
                const double stat_cap = share + 6.0 * sqrt(share) + 24.0;
                // sorted input: the bracket is one contiguous run of tiles, dealt round-robin to the CTAs; a warp
                // takes 256 keys of every tile of its CTA
                const double grid = (double)nw;
                const double run_cap = (double)STR_TILE * (ceil((want / (double)STR_TILE + 2.0) / grid) + 1.0);
                cap = (unsigned long long)fmax(stat_cap, fmin(run_cap, want + 24.0));
                cap = (cap + 3) & ~3ull;
            }
            scap[b] = cap;
        }
        __syncthreads();
        if (tid == 0) {
            // CTA-major layout: the pieces of ONE streaming CTA (one per bracket) are contiguous, so that a CTA appends
            // into a handful of 2 MB pages instead of one page per bracket (the scattered stores of 148 CTAs x ~100
            // brackets thrash the TLBs: measured as 40 SMs running 15 % behind the others)
            unsigned long long off = 0;
            for (int b = 0; b < B; ++b) {
                buf.boff[b] = off;              // offset of the bracket's piece inside a CTA's block
                buf.bcap[b] = scap[b];          // capacity of ONE piece
                off += scap[b];
            }
            buf.hdr[H_PSTRIDE] = off;           // block of one CTA
            if (off * (unsigned long long)nw > buf.cand_cap[0]) raise_fail(buf.hdr, FAIL_CAPACITY);
        }
    }
    const Key kmin = sKminLut;  // origin of the cells
    for (int b = tid; b < B; b += nthreads) {
        buf.blo[b] = blo[b];
        buf.bhi[b] = bhi[b];
        buf.btie[b] = btie[b];
    }
    // pure entries (NaN keys never reach the tables: the streaming pass finds them with a floating-point compare)
    for (int i = tid; i < BRK_L2_PURE; i += nthreads) {
        unsigned short e = 0;
        if (i < 256) {
            const int j = i >> 1;
            const bool in = i & 1;
            if (j <= B && (!in || j > 0)) e = (unsigned short)pure_entry(j, in, btie);
        }
        buf.lut2[i] = e;
    }
    // ---- level 1: which cells hold a bracket edge -----------------------------------------------------
    // cell c (0 .. NC1-2) covers keys [kmin + (c << shift1), +2^shift1); cell NC1-1 is everything beyond the table
    // (keys below kmin are clamped onto kmin by the streaming pass)
    const Key width1 = ((Key)1 << shift1) - 1;
    auto cell_ends = [&](int c, Key &a, Key &e) {
        a = kmin + ((Key)c << shift1);
        e = (a > KMAX - width1) ? KMAX : a + width1;
    };
    // Kind of the key range [a, e]: 0 = one class throughout (pa), 1 = holds a bracket edge
    auto range_kind = [&](Key a, Key e, PointClass<Key> &pa) -> int {
        pa = classify_point(blo, bhi, B, a);
        const PointClass<Key> pe = classify_point(blo, bhi, B, e);
        return (pa.j == pe.j && pa.in == pe.in) ? 0 : 1;
    };
    // index of the tie bracket (a single key with a class of its own) that is ALL the range [a, e] holds, or -1
    auto tie1_of = [&](Key a, Key e) -> int {
        int lo = 0, hi = B;  // first bracket that ends at or after a
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (bhi[mid] < a) lo = mid + 1;
            else hi = mid;
        }
        const int b = lo;
        if (b >= B || blo[b] != bhi[b] || btie[b] == 0xff) return -1;
        if (blo[b] < a || blo[b] > e) return -1;
        if (b + 1 < B && blo[b + 1] <= e) return -1;
        return b;
    };
    const Key max_cell = (KMAX - kmin) >> shift1;  // cells past the end of the key space can never be hit
    for (int c = tid; c < BRK_NC1; c += nthreads) {
        unsigned e1;
        if (c == BRK_NC1 - 1 || (Key)c > max_cell) {
            e1 = (unsigned)B;  // beyond the last bracket (or beyond the key space): pure "above"
        } else {
            Key a, e;
            cell_ends(c, a, e);
            PointClass<Key> pa;
            int tb;
            if (range_kind(a, e, pa) == 0) {
                e1 = pure_entry(pa.j, pa.in, btie);  // a pure cell: the final entry, no second table read
            } else if ((tb = tie1_of(a, e)) >= 0) {
                e1 = ENT_TIE1 | (unsigned)tb;         // one tie key and nothing else: final as well
            } else {
                const int slot = atomicAdd(&sNmixed, 1);
                mixcell[slot] = (unsigned short)c;
                e1 = 0x80000000u | (unsigned)slot;  // the table base is filled in once M is known
            }
        }
        buf.lut1[c] = e1;
    }
    __syncthreads();
    const int nmixed = sNmixed;
    if (tid == 0) {
        int M = 0;
        const int nm = nmixed > 0 ? nmixed : 1;
        const int m_max = sHi32 ? shift1 - 32 : shift1;
        while (M < m_max && M < 14 && ((nm << (M + 1)) <= BRK_L2_SUB)) ++M;
        sM = M;
        buf.lutc->kmin = kmin;
        buf.lutc->kmin_true = blo[0];
        buf.lutc->hi32 = sHi32;
        buf.lutc->shift1 = shift1;
        buf.lutc->subshift = shift1 - M;
        buf.lutc->submask = (1u << M) - 1u;
        buf.lutc->B = B;
        buf.lutc->ncls_used = (int)buf.hdr[H_NCLS_USED];
        buf.lutc->count_only = sLow;
    }
    __syncthreads();
    const int M = sM, subshift = shift1 - M;
    const Key width2 = ((Key)1 << subshift) - 1;
    for (int slot = tid; slot < nmixed; slot += nthreads)
        buf.lut1[mixcell[slot]] = 0x80000000u | (BRK_L2_PURE + ((unsigned)slot << M));
    // ---- level 2: the sub-cells of those cells, all threads ---------------------------------------------
    for (unsigned g = tid; g < ((unsigned)nmixed << M); g += nthreads) {
        const unsigned slot = g >> M, sidx = g & ((1u << M) - 1u);
        Key a, e;
        cell_ends((int)mixcell[slot], a, e);
        const Key a2 = a + ((Key)sidx << subshift);
        Key e2 = (a2 > KMAX - width2) ? KMAX : a2 + width2;
        if (e2 > e) e2 = e;
        PointClass<Key> p2;
        const int k2 = range_kind(a2, e2, p2);
        const int tb2 = k2 == 0 ? -1 : tie1_of(a2, e2);
        const unsigned short ent = k2 == 0 ? (unsigned short)pure_entry(p2.j, p2.in, btie)
                                           : (tb2 >= 0 ? (unsigned short)(ENT_TIE1 | (unsigned)tb2)
                                                       : (unsigned short)(ENT_MIXED | (unsigned)p2.j));
        buf.lut2[BRK_L2_PURE + g] = ent;
    }
}

// ------------------------------------------------------------------------------------------------
// 4. the streaming pass
// ------------------------------------------------------------------------------------------------
constexpr int STR_EPT = STR_TILE / STR_THREADS;  // keys per thread per tile
static_assert(STR_TILE == STR_THREADS * STR_EPT, "tile size");
constexpr int STR_SPILL_TILES = 255 / STR_EPT;  // 8-bit private counters
constexpr int STR_RING = 32 * STR_EPT;          // per-warp candidate ring (a power of two): one tile's worth
static_assert((STR_RING & (STR_RING - 1)) == 0, "ring size");

// shared-window accesses of the streaming pass (32-bit addresses: no generic-to-shared conversion in the hot loop)
__device__ __forceinline__ unsigned sm_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned lds_u8(unsigned a) {
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u8(unsigned a, unsigned v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// predicated pair of stores (candidate key + its bracket) without a branch
__device__ __forceinline__ void sts_key64_if(unsigned flag, unsigned akey, unsigned long long k, unsigned abrk, unsigned b) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.shared.u64 [%1], %2;\n\t@p st.shared.u8 [%3], %4;\n\t}"
                 ::"r"(flag), "r"(akey), "l"(k), "r"(abrk), "r"(b) : "memory");
}
__device__ __forceinline__ void sts_key32_if(unsigned flag, unsigned akey, unsigned k, unsigned abrk, unsigned b) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.shared.u32 [%1], %2;\n\t@p st.shared.u8 [%3], %4;\n\t}"
                 ::"r"(flag), "r"(akey), "r"(k), "r"(abrk), "r"(b) : "memory");
}

// this CTA's piece of one bracket's candidate segment: first element (index into cand[0]) and capacity
struct __align__(16) StrPiece { unsigned long long base; unsigned cap, pad; };

template <typename Key> __host__ __device__ constexpr size_t stream_smem_bytes() {
    return (size_t)BRK_NCLS * STR_THREADS + (size_t)BRK_NC1 * 4 + (((size_t)BRK_L2N * 2 + 15) & ~(size_t)15) +
           2 * (size_t)BRK_MAXB * sizeof(Key) + 128 + (size_t)BRK_MAXB * sizeof(StrPiece) +
           (size_t)BRK_MAXB * 4 + (size_t)BRK_NCLS * 4 + (size_t)STR_WARPS * STR_RING * (sizeof(Key) + 1) + 64;
}

// slow path of the classification: a sub-cell that holds a bracket edge.  `ent` carries the class at the lower end of
// the sub-cell; the edges inside it are walked -- two steps without a branch (tie-heavy lanes send a fifth of their keys
// here, and the lanes of a warp must not diverge), more in a loop.  Keys below the table origin were clamped into
// sub-cell 0 of cell 0 by the caller, whose entry says class 0: no bracket edge is <= such a key, so they stay there.
template <typename Key>
__device__ __forceinline__ unsigned brk_resolve_entry(Key k, unsigned ent, const Key *blo, const Key *bhi,
                                                      const unsigned char *btie, int B) {
    int j = (int)(ent & 0xffu);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int jj = j < B ? j : B - 1;
        j += (j < B && blo[jj] <= k) ? 1 : 0;
    }
    while (j < B && blo[j] <= k) ++j;
    const int jm = j > 0 ? j - 1 : 0;
    const bool in = j > 0 && k <= bhi[jm];
    const unsigned t = btie[jm];
    return in ? (t != 0xffu ? t : ((unsigned)j | ENT_COMPACT)) : (unsigned)j;
}
// a cell that holds one tie key and nothing else (ENT_TIE1): two compares
template <typename Key>
__device__ __forceinline__ unsigned brk_tie1_entry(Key k, unsigned ent, const Key *blo, const unsigned char *btie) {
    const unsigned b = ent & 0xffu;
    const Key v = blo[b];
    const unsigned t = btie[b];
    return k < v ? b : (k == v ? t : b + 1u);
}

// HI32: 64-bit keys, tables indexed with the high word (see brk_plan_kernel).
// COUNT: low-cardinality lanes (every bracket is one distinct key): classify and count only, nothing is compacted.
// One kernel holds all variants: which one runs is the plan's decision, read from device memory (brk_stream_kernel).
template <typename T, bool HI32, bool COUNT>
__device__ __forceinline__ void
brk_stream_body(const T *__restrict__ data, unsigned long long n, const BrkBuf<typename Traits<T>::Key> &buf) {
    using Tr = Traits<T>;
    using Key = typename Tr::Key;
    constexpr bool IS_FLOAT = Tr::is_float;
    constexpr int ES = (int)sizeof(T);
    constexpr int VEC = 16 / ES;             // elements per 16-byte load
    // keys per thread and tile: four -- eight for the 4-byte types when the pass only counts (then nothing but the loads in
    // flight limits it: two 16-byte loads per thread keep 2 x 32 KB per SM on their way, as the 8-byte types do)
    constexpr int EPT = (COUNT && ES == 4) ? 2 * STR_EPT : STR_EPT;
    constexpr int TILE = STR_THREADS * EPT;
    constexpr int SPILL_TILES = 255 / EPT;   // 8-bit private counters
    constexpr int NLD = EPT / VEC;           // 16-byte loads per thread per tile
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // ---- carve shared memory -----------------------------------------------------------------------
    unsigned char *p = smem_raw;
    unsigned char *__restrict__ counters = p; p += (size_t)BRK_NCLS * STR_THREADS;                 // [class][thread] 8-bit
    const unsigned *__restrict__ lut1 = reinterpret_cast<unsigned *>(p); p += (size_t)BRK_NC1 * 4;
    const unsigned short *__restrict__ lut2 = reinterpret_cast<unsigned short *>(p); p += ((size_t)BRK_L2N * 2 + 15) & ~(size_t)15;
    Key *blo = reinterpret_cast<Key *>(p); p += (size_t)BRK_MAXB * sizeof(Key);
    Key *bhi = reinterpret_cast<Key *>(p); p += (size_t)BRK_MAXB * sizeof(Key);
    unsigned char *btie = p; p += 128;
    StrPiece *piece = reinterpret_cast<StrPiece *>(p); p += (size_t)BRK_MAXB * sizeof(StrPiece);       // this CTA's pieces
    unsigned *pfill = reinterpret_cast<unsigned *>(p); p += (size_t)BRK_MAXB * 4;                        // their fills
    unsigned *cta_cnt = reinterpret_cast<unsigned *>(p); p += (size_t)BRK_NCLS * 4;                      // spilled private counters
    Key *__restrict__ wkey = reinterpret_cast<Key *>(p); p += (size_t)STR_WARPS * STR_RING * sizeof(Key);  // per-warp candidate ring
    unsigned char *wbrk = p;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const BrkLutConst<Key> lc = *buf.lutc;
#ifdef NDS_STR_TIMING  // diagnostic build (make strvar TIMING=1): per-CTA cycle counts of the streaming pass
    const long long t_begin = clock64();
    long long t_loop = 0;
#endif
    const int B = lc.B;
    const Key kmin = lc.kmin;
    const int shift1 = lc.shift1, subshift = lc.subshift;
    const unsigned submask = lc.submask;
    const int ncls_used = lc.ncls_used;
    const unsigned nw = (unsigned)buf.hdr[H_NW];        // pieces per bracket segment = streaming CTAs

    bool hashed = false;   // low-cardinality lanes with a perfect hash of their distinct keys: no lookup tables at all
    if constexpr (COUNT) hashed = lc.hashed != 0;
    if (!hashed) {
        for (int i = tid; i < BRK_NC1; i += STR_THREADS) const_cast<unsigned *>(lut1)[i] = buf.lut1[i];
        for (int i = tid; i < BRK_L2N; i += STR_THREADS) const_cast<unsigned short *>(lut2)[i] = buf.lut2[i];
    }
    for (int i = tid; i < BRK_MAXB; i += STR_THREADS) {
        blo[i] = i < B ? buf.blo[i] : (Key)0;
        bhi[i] = i < B ? buf.bhi[i] : (Key)0;
        btie[i] = i < B ? buf.btie[i] : (unsigned char)0xff;
        const unsigned cap = i < B ? (unsigned)buf.bcap[i] : 0u;
        StrPiece pc;
        pc.base = i < B ? buf.boff[i] + (unsigned long long)blockIdx.x * buf.hdr[H_PSTRIDE] : 0ull;
        pc.cap = cap;
        pc.pad = 0;
        piece[i] = pc;
        pfill[i] = 0;
    }
    for (int i = tid; i < BRK_NCLS; i += STR_THREADS) cta_cnt[i] = 0;
    for (int i = tid; i < BRK_NCLS * STR_THREADS / 4; i += STR_THREADS) reinterpret_cast<unsigned *>(counters)[i] = 0;
    __syncthreads();
    // the hash table {key, class} lives where the lookup tables would be.  Empty slots repeat the entry of key 0: a key
    // equal to it probes its own slot, so a copy elsewhere can never be hit by mistake.
    constexpr int HB = brk_hash_bits<Key>();
    constexpr int HENT = sizeof(Key) == 8 ? 16 : 8;   // bytes per entry
    static_assert(((size_t)HENT << HB) <= (size_t)BRK_NC1 * 4 + (size_t)BRK_L2N * 2, "hash table fits the lookup tables' room");
    const unsigned htab_a = sm_addr(lut1);
    const unsigned hmul = lc.hmul;
    auto hash_store = [&](unsigned slot, Key key, unsigned cls) {
        if constexpr (sizeof(Key) == 8)
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(htab_a + slot * 16u), "r"((unsigned)key),
                         "r"((unsigned)((unsigned long long)key >> 32)), "r"(cls), "r"(0u) : "memory");
        else
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(htab_a + slot * 8u), "r"((unsigned)key), "r"(cls) : "memory");
    };
    if (hashed) {
        for (unsigned i = tid; i < (1u << HB); i += STR_THREADS) hash_store(i, blo[0], (unsigned)btie[0]);
        __syncthreads();
        if (tid < B) hash_store((brk_hash_fold<Key>(blo[tid]) * hmul) >> (32 - HB), blo[tid], (unsigned)btie[tid]);
        __syncthreads();
    }
    // class of a key without any table (count-only lanes: every bracket is a single key): binary search over the keys
    auto classify_slow = [&](Key k) -> unsigned {
        if (key_is_nan<T>(k)) return (unsigned)BRK_INVALID_CLS;
        int a = 0, e = B;   // number of brackets with blo <= k
        while (a < e) {
            const int mid = (a + e) >> 1;
            if (blo[mid] <= k) a = mid + 1;
            else e = mid;
        }
        return (a > 0 && blo[a - 1] == k && btie[a - 1] != 0xffu) ? (unsigned)btie[a - 1] : (unsigned)a;
    };

    // conflict-free private counter of this thread in a class row of 512 bytes: the 32 lanes of a warp hit 32
    // different banks, four warps share a 128-byte line (byte 0..3 of each word)
    unsigned char *__restrict__ myctr = counters + (warp >> 2) * 128 + lane * 4 + (warp & 3);
    const unsigned myctr_a = sm_addr(myctr);
    Key *cand = buf.cand[0];
    Key *mykey = wkey + warp * STR_RING;
    unsigned char *mybrk = wbrk + warp * STR_RING;
    const unsigned mykey_a = sm_addr(mykey), mybrk_a = sm_addr(mybrk);
    unsigned whead = 0, wpend = 0;  // the warp's ring of candidates that wait for a full group of 32 (warp-uniform)

    // Two table reads per key, no side effects.  Keys below the table origin are clamped onto it: the origin lies a
    // whole cell below the first bracket (brk_plan_kernel), so they read a "below every bracket" entry; keys past the
    // last bracket fall into cells that are pure "above" (the last cell takes everything beyond the table).
    const unsigned lut2_a = sm_addr(lut2);
    // second table read only for the lanes whose cell is subdivided (e1 < 0): a predicated load, so that the others do
    // not take part in its bank conflicts
    auto second = [&](unsigned e1, unsigned sub) -> unsigned {
        unsigned ent;
        const unsigned a = lut2_a + 2u * ((e1 & 0x7fffffffu) + (sub & submask));
        asm("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %1, 0;\n\tmov.u32 %0, %1;\n\t@p ld.shared.u16 %0, [%2];\n\t}"
            : "=r"(ent) : "r"(e1), "r"(a));
        return ent;
    };
    auto lookup = [&](Key k) -> unsigned {
        if constexpr (HI32) {
            const unsigned hk = (unsigned)(k >> 32), h0 = (unsigned)(kmin >> 32);
            const unsigned d = umax(hk, h0) - h0;
            const unsigned cc = d >> (shift1 - 32);
            const unsigned cell = cc < (unsigned)(BRK_NC1 - 1) ? cc : (unsigned)(BRK_NC1 - 1);
            return second(lut1[cell], d >> (subshift - 32));
        } else {
            const Key d = (k > kmin ? k : kmin) - kmin;
            const Key cc = d >> shift1;
            const unsigned cell = (unsigned)(cc < (Key)(BRK_NC1 - 1) ? cc : (Key)(BRK_NC1 - 1));
            return second(lut1[cell], (unsigned)(d >> subshift));
        }
    };
    // One group of up to 32 candidates: lane i appends `k` to this CTA's piece of bracket `b` (if `has`).  Lanes with
    // the same bracket are grouped with match.any and the group leader reserves the positions with ONE shared-memory
    // atomic on the CTA's fill of that bracket.  Called by all 32 lanes; no CTA barrier anywhere.
    auto append = [&](bool has, Key k, unsigned b) {
        const unsigned bb = has ? b : 0xffffffffu;               // the idle lanes form a group of their own
        const unsigned peers = __match_any_sync(0xffffffffu, bb);
        const int leader = __ffs(peers) - 1;
        const unsigned rank = __popc(peers & ((1u << lane) - 1u));
        unsigned base = 0;
        if (has && lane == leader) base = atomicAdd(&pfill[b], (unsigned)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (has) {
            const StrPiece pc = piece[b];
            const unsigned pos = base + rank;
            if (pos < pc.cap) cand[pc.base + pos] = k;
            else raise_fail(buf.hdr, FAIL_CAPACITY);
        }
    };
    // full groups of the ring (`all`: the rest too)
    auto drain = [&](bool all) {
#pragma unroll 1
        while (wpend >= 32u || (all && wpend > 0u)) {
            const unsigned take = wpend < 32u ? wpend : 32u;
            const unsigned i = (whead + (unsigned)lane) & (unsigned)(STR_RING - 1);
            const bool has = (unsigned)lane < take;
            append(has, mykey[i], (unsigned)mybrk[i] - 1u);
            whead = (whead + take) & (unsigned)(STR_RING - 1);
            wpend -= take;
        }
        __syncwarp();
    };
    // private 8-bit counters -> 32-bit CTA counters (every STR_SPILL_TILES tiles)
    auto spill_counters = [&]() {
        __syncthreads();
        for (int c = warp; c < BRK_NCLS; c += STR_WARPS) {
            if (c >= ncls_used && c != BRK_INVALID_CLS) continue;
            unsigned *row = reinterpret_cast<unsigned *>(counters + (size_t)c * STR_THREADS);
            unsigned sum = 0;
#pragma unroll
            for (int i = 0; i < STR_THREADS / 4 / 32; ++i) {
                const unsigned w = row[i * 32 + lane];
                sum = __dp4a(w, 0x01010101u, sum);
                row[i * 32 + lane] = 0;
            }
            sum = __reduce_add_sync(0xffffffffu, sum);
            if (lane == 0 && sum) cta_cnt[c] += sum;
        }
        __syncthreads();
    };

    // ---- full, 16-byte aligned tiles ------------------------------------------------------------------
    const uintptr_t addr = reinterpret_cast<uintptr_t>(data);
    unsigned long long head = ((16 - (addr & 15)) & 15) / ES;  // elements before the first aligned one
    if (head > n) head = n;
    const unsigned long long n_tiles = (n - head) / TILE;
    const T *adata = data + head;
    const uint4 *vdata = reinterpret_cast<const uint4 *>(adata);

    auto load_tile = [&](unsigned long long tile, uint4 (&q)[NLD]) {
        const uint4 *src = vdata + tile * (TILE / VEC);
#pragma unroll
        for (int j = 0; j < NLD; ++j) q[j] = __ldg(src + j * STR_THREADS + tid);
    };
    auto process_tile = [&](const uint4 (&q)[NLD]) {
        Key k[EPT];
        bool nan = false;  // floats: NaNs are found with one floating-point compare per key (their keys lie outside
                           // [key(-inf), key(+inf)] and the tables do not know about them)
#pragma unroll
        for (int j = 0; j < NLD; ++j) {
            if constexpr (ES == 4) {
                const uint32_t w[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const T v = from_raw<T>(w[u]);
                    if constexpr (IS_FLOAT) nan |= v != v;
                    k[j * 4 + u] = Tr::to_key(v);
                }
            } else {
                const uint64_t w0 = (uint64_t)q[j].x | ((uint64_t)q[j].y << 32);
                const uint64_t w1 = (uint64_t)q[j].z | ((uint64_t)q[j].w << 32);
                const T v0 = from_raw<T>(w0), v1 = from_raw<T>(w1);
                if constexpr (IS_FLOAT) nan |= (v0 != v0) | (v1 != v1);
                k[j * 2] = Tr::to_key(v0);
                k[j * 2 + 1] = Tr::to_key(v1);
            }
        }
        // phase 1: all table reads of the tile are independent of each other
        unsigned ent[EPT];
        unsigned slow = 0;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            ent[e] = lookup(k[e]);
            slow |= ent[e];
        }
        if constexpr (IS_FLOAT) {
            if (__any_sync(0xffffffffu, nan)) {  // rare: NaNs get the invalid class, whatever the tables said
                slow = 0;
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    if (key_is_nan<T>(k[e])) ent[e] = (unsigned)BRK_INVALID_CLS;
                    slow |= ent[e];
                }
            }
        }
        // phase 2: keys in a cell with one tie key (two compares), the few keys in a sub-cell with a bracket edge (a walk)
        if (slow & (ENT_MIXED | ENT_TIE1)) {
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                if (ent[e] & ENT_TIE1) ent[e] = brk_tie1_entry<Key>(k[e], ent[e], blo, btie);
                else if (ent[e] & ENT_MIXED) ent[e] = brk_resolve_entry<Key>(k[e], ent[e], blo, bhi, btie, B);
            }
        }
        // phase 3: private counters (plain read-modify-write of this thread's own bytes, one after the other: two
        // keys of a thread may share a class), candidate mask
        unsigned cflag[EPT];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const unsigned a = myctr_a + (ent[e] & 0xffu) * (unsigned)STR_THREADS;
            sts_u8(a, lds_u8(a) + 1u);
            cflag[e] = (ent[e] >> 8) & 1u;
        }
        if constexpr (COUNT) return;  // nothing is ever compacted in this mode
        // phase 4: the ~10 % of keys inside a bracket are packed into the warp's ring; full groups of 32 are appended to
        // the brackets' segments, the rest waits for the next tile.  Ring positions from one ballot per key slot (the
        // ballots are independent of each other; a shuffle scan is a chain of five dependent shuffles): the candidates
        // of slot e of all lanes come before those of slot e + 1 -- the order inside the ring does not matter.
        const unsigned lt = (1u << lane) - 1u;
        unsigned total = 0, mypos[EPT];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const unsigned bal = __ballot_sync(0xffffffffu, cflag[e] != 0u);
            mypos[e] = total + __popc(bal & lt);
            total += __popc(bal);
        }
        if (total == 0) return;  // a tile without candidates (common when brackets are few)
        if (wpend + total > (unsigned)STR_RING) drain(true);  // rare: make room for a tile full of candidates
        const unsigned pos0 = whead + wpend;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {  // the ring keeps the class (= bracket + 1) of a candidate
            const unsigned i = (pos0 + mypos[e]) & (unsigned)(STR_RING - 1);
            if constexpr (sizeof(Key) == 8) sts_key64_if(cflag[e], mykey_a + i * 8u, (unsigned long long)k[e], mybrk_a + i, ent[e]);
            else sts_key32_if(cflag[e], mykey_a + i * 4u, (unsigned)k[e], mybrk_a + i, ent[e]);
        }
        wpend += total;
        __syncwarp();
        if (wpend >= 32u) drain(false);
    };

    // low-cardinality lanes with a perfect hash: one multiply, one probe, one compare per key; a key the table does not
    // hold (unseen by the sample, or a NaN) is classified by the binary search
    auto process_tile_hashed = [&](const uint4 (&q)[NLD]) {
        unsigned cls[EPT];
        Key k[EPT];
        bool miss = false;
#pragma unroll
        for (int j = 0; j < NLD; ++j) {
            if constexpr (ES == 4) {
                const uint32_t w[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const Key kk = Tr::to_key(from_raw<T>(w[u]));
                    k[j * 4 + u] = kk;
                    const unsigned slot = ((unsigned)kk * hmul) >> (32 - HB);
                    unsigned tk, tc;
                    asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(tk), "=r"(tc) : "r"(htab_a + slot * 8u));
                    const bool hit = tk == (unsigned)kk;
                    cls[j * 4 + u] = hit ? tc : 0xffffffffu;
                    miss |= !hit;
                }
            } else {
                const uint64_t w0 = (uint64_t)q[j].x | ((uint64_t)q[j].y << 32);
                const uint64_t w1 = (uint64_t)q[j].z | ((uint64_t)q[j].w << 32);
                const Key kk[2] = {Tr::to_key(from_raw<T>(w0)), Tr::to_key(from_raw<T>(w1))};
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    k[j * 2 + u] = kk[u];
                    const unsigned slot = (brk_hash_fold<Key>(kk[u]) * hmul) >> (32 - HB);
                    unsigned t0, t1, tc, tp;
                    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(t0), "=r"(t1), "=r"(tc), "=r"(tp) : "r"(htab_a + slot * 16u));
                    const bool hit = t0 == (unsigned)kk[u] && t1 == (unsigned)((unsigned long long)kk[u] >> 32);
                    cls[j * 2 + u] = hit ? tc : 0xffffffffu;
                    miss |= !hit;
                }
            }
        }
        if (miss) {
#pragma unroll
            for (int e = 0; e < EPT; ++e)
                if (cls[e] == 0xffffffffu) cls[e] = classify_slow(k[e]);
        }
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const unsigned a = myctr_a + cls[e] * (unsigned)STR_THREADS;
            sts_u8(a, lds_u8(a) + 1u);
        }
    };

    // (Measured and dropped: rotating the tile assignment per round, evict-first loads, L2 bulk prefetch two to eight
    // rounds ahead -- none of them moves the pass; what did was the CTA-major layout of the candidate pieces.)
    unsigned tiles_since_spill = 0;
    auto run_tiles = [&](auto &&process) {
        uint4 qa[NLD], qb[NLD];
        unsigned long long tile = blockIdx.x;
        if (tile < n_tiles) load_tile(tile, qa);
        while (tile < n_tiles) {
            const unsigned long long next = tile + gridDim.x;
            if (next < n_tiles) load_tile(next, qb);
            process(qa);
            if (++tiles_since_spill >= SPILL_TILES) {
                spill_counters();
                tiles_since_spill = 0;
            }
            tile = next;
            if (tile >= n_tiles) break;
            const unsigned long long next2 = tile + gridDim.x;
            if (next2 < n_tiles) load_tile(next2, qa);
            process(qb);
            if (++tiles_since_spill >= SPILL_TILES) {
                spill_counters();
                tiles_since_spill = 0;
            }
            tile = next2;
        }
    };
    if constexpr (COUNT) {
        if (hashed) run_tiles(process_tile_hashed);
        else run_tiles(process_tile);
    } else {
        run_tiles(process_tile);
    }
#ifdef NDS_STR_TIMING
    t_loop = clock64();
#endif
    drain(true);
    spill_counters();
    // ---- ragged head and tail: the last CTA, element by element ---------------------------------------
    if (blockIdx.x == gridDim.x - 1) {
        const unsigned long long tail0 = head + n_tiles * TILE;
        const unsigned long long n_ragged = head + (n - tail0);
        unsigned rounds = 0;
        for (unsigned long long i0 = 0; i0 < n_ragged; i0 += STR_THREADS) {
            const unsigned long long i = i0 + tid;
            unsigned ent = 0;
            Key k = 0;
            const bool have = i < n_ragged;
            if (have) {
                const unsigned long long idx = i < head ? i : tail0 + (i - head);
                k = Tr::to_key(data[idx]);
                if (hashed) {
                    ent = classify_slow(k);
                } else if (key_is_nan<T>(k)) {
                    ent = (unsigned)BRK_INVALID_CLS;
                } else {
                    ent = lookup(k);
                    if (ent & ENT_TIE1) ent = brk_tie1_entry<Key>(k, ent, blo, btie);
                    else if (ent & ENT_MIXED) ent = brk_resolve_entry<Key>(k, ent, blo, bhi, btie, B);
                }
                myctr[(ent & 0xffu) * STR_THREADS] += 1;
            }
            append(have && (ent & ENT_COMPACT), k, ((ent & 0xffu) - 1u) & 0xffu);
            if (++rounds >= 200) {
                spill_counters();
                rounds = 0;
            }
        }
        spill_counters();
    }
    __syncthreads();
    for (int c = tid; c < BRK_NCLS; c += STR_THREADS)
        if (cta_cnt[c]) atomicAdd(&buf.red_stream[c], (unsigned long long)cta_cnt[c]);
    // fills of this CTA's pieces; their sums are the brackets' in-counts
    for (int b = tid; b < B; b += STR_THREADS) {
        const unsigned f = pfill[b];
        buf.fill[(size_t)b * nw + blockIdx.x] = f;
        if (f) atomicAdd(&buf.cur_local[b], (unsigned long long)f);
    }
#ifdef NDS_STR_TIMING
    if (tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        printf("strtime cta %d sm %u loop %lld total %lld\n", (int)blockIdx.x, smid, t_loop - t_begin, clock64() - t_begin);
    }
#endif
}

template <typename T>
__global__ void __launch_bounds__(STR_THREADS, 1)
brk_stream_kernel(const T *__restrict__ data, unsigned long long n, BrkBuf<typename Traits<T>::Key> buf) {
    if (failed_any(buf.hdr)) return;  // the plan already gave up: the caller falls back
    const int hi32 = buf.lutc->hi32, count_only = buf.lutc->count_only;  // (uniform over the grid)
    if constexpr (sizeof(T) == 8) {
        if (hi32) {
            if (count_only) brk_stream_body<T, true, true>(data, n, buf);
            else brk_stream_body<T, true, false>(data, n, buf);
            return;
        }
    }
    if (count_only) brk_stream_body<T, false, true>(data, n, buf);
    else brk_stream_body<T, false, false>(data, n, buf);
}

// in-counts of the non-tie brackets are their segment fills
template <typename Key> __global__ void brk_publish_fill_kernel(BrkBuf<Key> buf) {
    const int B = (int)buf.hdr[H_B];
    for (int b = threadIdx.x; b < B; b += blockDim.x) buf.red_stream[BRK_NCLS + b] = buf.cur_local[b];
    if (threadIdx.x == 0) buf.red_stream[BRK_NCLS + BRK_MAXB] = failed_any(buf.hdr) ? 1ull : 0ull;  // summed over ranks
}

// ------------------------------------------------------------------------------------------------
// 5. resolve: ranks -> (bracket, rank inside)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
brk_resolve_kernel(BrkBuf<typename Traits<T>::Key> buf, QuantArgs qa, int masked_or_skip) {
    using Tr = Traits<T>;
    using Key = typename Tr::Key;
    __shared__ unsigned long long below[BRK_MAXB], incnt[BRK_MAXB];
    __shared__ int seg_used[BRK_MAXB];
    // everything the single-threaded parts look at is staged in shared memory first (this kernel sits between the
    // streaming pass and the refinement on the critical path of every call: one memory latency per table entry adds up)
    __shared__ unsigned long long s_red[BRK_NCLS + BRK_MAXB + 2];
    __shared__ unsigned char s_btie[BRK_MAXB], s_segkind[BRK_MAXB];
    __shared__ Key s_blo[BRK_MAXB], s_bhi[BRK_MAXB];
    __shared__ unsigned long long s_rank[BRK_MAXTASK];
    __shared__ int s_first[BRK_MAXTASK], s_id[BRK_MAXTASK];
    const int tid = threadIdx.x;
    if (buf.red_stream[BRK_NCLS + BRK_MAXB]) {  // some rank gave up in the streaming pass: everybody does
        if (tid == 0) raise_fail(buf.hdr, FAIL_REMOTE);
        return;
    }
    if (buf.hdr[H_FAIL]) return;
    const int B = (int)buf.hdr[H_B];
    const int R = 2 * qa.nq;
    const unsigned long long n_total = buf.hdr[H_NTOTAL];
    for (int i = tid; i < BRK_NCLS + BRK_MAXB + 2; i += blockDim.x) s_red[i] = buf.red_stream[i];
    for (int b = tid; b < BRK_MAXB; b += blockDim.x) {
        seg_used[b] = 0;
        s_btie[b] = b < B ? buf.btie[b] : (unsigned char)0xff;
        s_blo[b] = b < B ? buf.blo[b] : (Key)0;
        s_bhi[b] = b < B ? buf.bhi[b] : (Key)0;
    }
    __syncthreads();
    const unsigned long long n_invalid = s_red[BRK_INVALID_CLS];
    const bool drop = masked_or_skip != 0;
    const unsigned long long n_eff = drop ? n_total - n_invalid : n_total;
    if (tid == 0) {
        if (!drop && Tr::is_float && n_invalid) atomicOr(qa.status, STATUS_BIT_NAN);
        buf.hdr[H_NINVALID] = n_invalid;
        buf.hdr[H_NEFF] = n_eff;
        buf.hdr[H_EMPTY] = n_eff == 0;
        buf.hdr[H_NSEG] = (unsigned long long)B;
    }
    if (n_eff == 0) {
        if (tid == 0) {
            buf.hdr[H_NTASK] = 0;
            buf.hdr[H_NACTIVE] = 0;
            buf.hdr[H_NCHUNK_COUNT] = 0;
            buf.hdr[H_NCHUNK_COMPACT] = 0;
            buf.cprefix_count[0] = 0;
        }
        return;
    }
    if (tid == 0) {  // exact counts below / inside each bracket
        unsigned long long run = 0, check = 0;
        for (int b = 0; b < B; ++b) {
            const unsigned char t = s_btie[b];
            const unsigned long long in_b = t != 0xff ? s_red[t] : s_red[BRK_NCLS + b];
            // class j counts the gap before bracket j plus (non-tie) the inside of bracket j-1
            unsigned long long gap = s_red[b];
            if (b > 0 && s_btie[b - 1] == 0xff) gap -= incnt[b - 1];
            run += gap;
            below[b] = run;
            incnt[b] = in_b;
            run += in_b;
        }
        unsigned long long gap = s_red[B];
        if (B > 0 && s_btie[B - 1] == 0xff) gap -= incnt[B - 1];
        check = run + gap;
        if (check != n_total - n_invalid) raise_fail(buf.hdr, FAIL_INCONSISTENT);  // every non-NaN element is in one class
    }
    // ---- requested ranks (as long_decide does): slots -> unique ranks -> tasks -------------------------
    for (int t = tid; t < qa.nq; t += blockDim.x) {
        RankMath rm = slot_rank_math(qa, t, n_eff);
        const bool nl = needs_lower(qa.interp, rm.frac), nh = needs_higher(qa.interp, rm.frac);
        s_rank[2 * t] = nl ? rm.lower : rm.higher;
        s_rank[2 * t + 1] = nh ? rm.higher : rm.lower;
    }
    __syncthreads();
    if (ld_now(&buf.hdr[H_FAIL])) return;
    for (int i = tid; i < R; i += blockDim.x) {  // first occurrence of each rank, numbered in slot order
        const unsigned long long ri = s_rank[i];
        int first = i;
        for (int k = 0; k < i; ++k)
            if (s_rank[k] == ri) { first = k; break; }
        s_first[i] = first == i ? -1 - i : first;  // negative: "I am a first occurrence"
    }
    __syncthreads();
    for (int i = tid; i < R; i += blockDim.x) {  // dense task ids in slot order: every slot learns the id of its rank
        const int f = s_first[i] < 0 ? i : s_first[i];
        int id = 0;
        for (int k = 0; k < f; ++k) id += s_first[k] < 0 ? 1 : 0;
        s_id[i] = id;
    }
    __syncthreads();
    for (int i = tid; i < R; i += blockDim.x) {
        buf.slot_task[i] = s_first[i];
        buf.slot_rank[i] = (unsigned long long)s_id[i];  // slot -> task id (what the output kernel reads)
        if (s_first[i] < 0) {
            const unsigned long long r = s_rank[i];
            BrkTask tk;
            tk.r = 0;
            tk.value = 0;
            tk.seg = -1;
            tk.state = TASK_ACTIVE;
            // bracket holding rank r
            int lo = 0, hi = B;  // first b with below[b] + incnt[b] > r
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (below[mid] + incnt[mid] <= r) lo = mid + 1;
                else hi = mid;
            }
            if (lo >= B || below[lo] > r) {
                raise_fail(buf.hdr, FAIL_ESCAPED);
            } else {
                tk.seg = lo;
                tk.r = r - below[lo];
                if (s_btie[lo] != 0xff || s_blo[lo] == s_bhi[lo]) {
                    tk.state = TASK_DONE;
                    tk.value = (unsigned long long)s_blo[lo];
                } else {
                    seg_used[lo] = 1;
                }
            }
            buf.task[s_id[i]] = tk;
        }
    }
    if (tid == 0) {
        int nt = 0;
        for (int i = 0; i < R; ++i) nt += s_first[i] < 0 ? 1 : 0;
        buf.hdr[H_NTASK] = (unsigned long long)nt;
    }
    __syncthreads();
    // ---- round-0 segments are the brackets -------------------------------------------------------------
    const unsigned long long final_cap = buf.hdr[H_FINALCAP], pstride = buf.hdr[H_PSTRIDE];
    for (int b = tid; b < B; b += blockDim.x) {
        BrkSeg<Key> s;
        s.lo = s_blo[b];
        s.hi = s_bhi[b];
        const int bits = key_bits_used<Key>((Key)(s.hi - s.lo));
        s.shift = bits > 8 ? bits - 8 : 0;
        s.flags = 0;
        if (s_btie[b] != 0xff) s.flags |= SEG_TIE;
        if (seg_used[b]) s.flags |= SEG_USED;
        s.off = buf.boff[b];
        s.piece_cap = buf.bcap[b];
        s.piece_stride = pstride;
        s.cnt_local = (s.flags & SEG_TIE) ? 0 : buf.cur_local[b];  // sum over this GPU's pieces
        s.cnt_global = incnt[b];
        s.cursor = 0;
        if ((s.flags & SEG_USED) && s.cnt_global <= final_cap) s.flags |= SEG_FINAL;
        buf.seg[0][b] = s;
        unsigned char kind = 0;  // 1 = goes through the refinement rounds, 2 = small enough to be sorted at once
        if ((s.flags & SEG_USED) && !(s.flags & (SEG_FINAL | SEG_TIE))) kind = 1;
        if ((s.flags & SEG_USED) && (s.flags & SEG_FINAL) && !(s.flags & SEG_TIE)) kind = 2;
        s_segkind[b] = kind;
    }
    __syncthreads();
    if (tid == 0) {
        int active = 0, nfinal = 0;
        for (int b = 0; b < B; ++b) {
            active += s_segkind[b] == 1 ? 1 : 0;
            nfinal += s_segkind[b] == 2 ? 1 : 0;
        }
        buf.hdr[H_NFINAL] = (unsigned long long)nfinal;   // small segments waiting for brk_final_kernel
        buf.hdr[H_NCHUNK_COUNT] = 0;   // round 0 walks the pieces (brk_refine_count0_kernel), not a chunk table
        buf.hdr[H_NACTIVE] = (unsigned long long)active;
        buf.hdr[H_ROUND] = 0;
        // what the exchange steps of the refinement look at: identical on every rank (H_FAIL, not the local fail word)
        const bool ok = ld_now(&buf.hdr[H_FAIL]) == 0;
        buf.hdr[H_GATE_ROUND] = (ok && active != 0) ? 1ull : 0ull;
        buf.hdr[H_GATE_SMALL] = (ok && nfinal != 0 && buf.hdr[H_MULTI]) ? 1ull : 0ull;
    }
}

// ------------------------------------------------------------------------------------------------
// 6. refine rounds over the candidate segments
// ------------------------------------------------------------------------------------------------
// 16 keys per thread of one full, 16-byte aligned chunk of BRK_RCHUNK keys
template <typename Key>
__device__ __forceinline__ void load_chunk_keys(const Key *chunk, int tid, Key (&k)[BRK_RCHUNK / BRK_THREADS]) {
    constexpr int VEC = 16 / (int)sizeof(Key), NLD = (BRK_RCHUNK / BRK_THREADS) / VEC;
    const uint4 *v = reinterpret_cast<const uint4 *>(chunk);
#pragma unroll
    for (int j = 0; j < NLD; ++j) {
        const uint4 q = __ldg(v + j * BRK_THREADS + tid);
        if constexpr (sizeof(Key) == 4) {
            k[j * 4] = q.x; k[j * 4 + 1] = q.y; k[j * 4 + 2] = q.z; k[j * 4 + 3] = q.w;
        } else {
            k[j * 2] = (Key)q.x | ((Key)q.y << 32);
            k[j * 2 + 1] = (Key)q.z | ((Key)q.w << 32);
        }
    }
}

// linear chunk id -> segment (binary search over the chunk prefix)
__device__ __forceinline__ int seg_of_chunk(const unsigned long long *prefix, int nseg, unsigned long long chunk) {
    int lo = 0, hi = nseg;  // last s with prefix[s] <= chunk
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (prefix[mid] <= chunk) lo = mid;
        else hi = mid;
    }
    return lo;
}

template <typename Key>
__global__ void __launch_bounds__(BRK_THREADS, 1) brk_refine_count_kernel(BrkBuf<Key> buf) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned short *counters = reinterpret_cast<unsigned short *>(smem_raw);  // [256][THREADS]
    __shared__ unsigned long long sprefix[BRK_MAXSEG + 1];
    if (failed_any(buf.hdr) || buf.hdr[H_ROUND] == 0 || buf.hdr[H_NACTIVE] == 0) return;
    const unsigned long long nchunks = buf.hdr[H_NCHUNK_COUNT];
    if (nchunks == 0) return;
    const int cur = (int)buf.hdr[H_CUR];
    const int nseg = (int)buf.hdr[H_NSEG];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i <= nseg; i += BRK_THREADS) sprefix[i] = buf.cprefix_count[i];
    for (int i = tid; i < 256 * BRK_THREADS / 2; i += BRK_THREADS) reinterpret_cast<unsigned *>(counters)[i] = 0;
    __syncthreads();
    // contiguous range of chunks per CTA: a CTA stays inside one segment for as long as possible
    const unsigned long long per = (nchunks + gridDim.x - 1) / gridDim.x;
    const unsigned long long c0 = (unsigned long long)blockIdx.x * per;
    const unsigned long long c1 = c0 + per < nchunks ? c0 + per : nchunks;
    if (c0 >= c1) return;
    unsigned short *myctr = counters + counter_slot();
    const Key *src = buf.cand[cur];
    const BrkSeg<Key> *segs = buf.seg[cur];
    int cur_seg = -1;
    unsigned pending = 0;  // chunks counted since the last spill
    auto spill = [&](int s) {
        __syncthreads();
        for (int c = warp; c < 256; c += BRK_THREADS / 32) {
            unsigned sum = 0;
            for (int i = lane; i < BRK_THREADS; i += 32) {
                sum += counters[c * BRK_THREADS + i];
                counters[c * BRK_THREADS + i] = 0;
            }
            sum = __reduce_add_sync(0xffffffffu, sum);
            if (lane == 0 && sum) atomicAdd(&buf.loc_refine[(size_t)s * 256 + c], (unsigned long long)sum);
        }
        __syncthreads();
    };
    for (unsigned long long ch = c0; ch < c1; ++ch) {
        const int s = seg_of_chunk(sprefix, nseg, ch);
        if (s != cur_seg || pending >= 4000) {
            if (cur_seg >= 0 && pending) spill(cur_seg);
            cur_seg = s;
            pending = 0;
        }
        const BrkSeg<Key> sg = segs[s];
        const unsigned long long e0 = (ch - sprefix[s]) * BRK_RCHUNK;
        const unsigned long long e1 = e0 + BRK_RCHUNK < sg.cnt_local ? e0 + BRK_RCHUNK : sg.cnt_local;
        const Key *sp = src + sg.off;
        if (e1 - e0 == BRK_RCHUNK) {
            Key k[BRK_RCHUNK / BRK_THREADS];
            load_chunk_keys<Key>(sp + e0, tid, k);
#pragma unroll
            for (int g = 0; g < BRK_RCHUNK / BRK_THREADS; g += 4) {  // as in the streaming pass: 4 loads, merged, 4 stores
                const unsigned c0 = (unsigned)((Key)(k[g] - sg.lo) >> sg.shift) & 0xffu;
                const unsigned c1 = (unsigned)((Key)(k[g + 1] - sg.lo) >> sg.shift) & 0xffu;
                const unsigned c2 = (unsigned)((Key)(k[g + 2] - sg.lo) >> sg.shift) & 0xffu;
                const unsigned c3 = (unsigned)((Key)(k[g + 3] - sg.lo) >> sg.shift) & 0xffu;
                unsigned short *p0 = myctr + c0 * BRK_THREADS, *p1 = myctr + c1 * BRK_THREADS,
                               *p2 = myctr + c2 * BRK_THREADS, *p3 = myctr + c3 * BRK_THREADS;
                const unsigned x0 = *p0, x1 = *p1, x2 = *p2, x3 = *p3;
                const unsigned m0 = 1u + (c1 == c0) + (c2 == c0) + (c3 == c0), m1 = 1u + (c0 == c1) + (c2 == c1) + (c3 == c1);
                const unsigned m2 = 1u + (c0 == c2) + (c1 == c2) + (c3 == c2), m3 = 1u + (c0 == c3) + (c1 == c3) + (c2 == c3);
                *p0 = (unsigned short)(x0 + m0);
                *p1 = (unsigned short)(x1 + m1);
                *p2 = (unsigned short)(x2 + m2);
                *p3 = (unsigned short)(x3 + m3);
            }
        } else {
            for (unsigned long long i = e0 + tid; i < e1; i += BRK_THREADS) {
                const Key k = sp[i];
                const unsigned c = (unsigned)((Key)(k - sg.lo) >> sg.shift) & 0xffu;
                myctr[c * BRK_THREADS] += 1;
            }
        }
        ++pending;
    }
    if (cur_seg >= 0 && pending) spill(cur_seg);
}

// Round 0 of the refinement reads the brackets' segments as the streaming pass wrote them: one piece per streaming
// CTA (fills in buf.fill).  A group of `cpb` CTAs shares one bracket and walks its pieces, so a CTA never mixes
// brackets in its private counters.  `COMPACT` = second pass: keep the keys whose digit was chosen.
template <typename Key, bool COMPACT>
__global__ void __launch_bounds__(BRK_THREADS, 2) brk_refine_round0_kernel(BrkBuf<Key> buf) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *counters = smem_raw;  // [256 digits + 1][THREADS] private 8-bit counters (count pass only): 64 KB, 3 CTAs per SM
    __shared__ short sid[256];
    __shared__ int wanted[256];   // the digits of this bracket that some rank chose (usually one or two)
    __shared__ int n_wanted;
    __shared__ unsigned sfill[148];                       // fills of this bracket's pieces
    __shared__ unsigned long long wseg_off[4], wseg_cap[4];  // destination segments of the first wanted digits
    if (failed_any(buf.hdr) || buf.hdr[H_ROUND] != 0 || buf.hdr[H_NACTIVE] == 0) return;
    if (COMPACT && !buf.hdr[H_DID]) return;
    const int B = (int)(COMPACT ? buf.hdr[H_OLD_NSEG] : buf.hdr[H_NSEG]);
    const unsigned nw = (unsigned)buf.hdr[H_NW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int WARPS = BRK_THREADS / 32;
    constexpr int KPT = BRK_RCHUNK / BRK_THREADS;
    // The active brackets' candidates, bracket after bracket, are one long sequence of keys; CTA c takes the part
    // [c K / G, (c + 1) K / G) of it, cut at piece boundaries (the pieces of one bracket have nearly equal fills, the
    // brackets themselves differ by a factor of five between the median and the first percentile).  A CTA touches one or
    // two brackets, so its private counters are spilled once or twice.
    __shared__ unsigned char act[BRK_MAXB];
    __shared__ unsigned long long acnt[BRK_MAXB], apre[BRK_MAXB + 1];
    __shared__ unsigned actmask[4];
    __shared__ int n_act;
    if (tid < 4) actmask[tid] = 0;
    __syncthreads();
    unsigned long long my_cnt = 0;
    if (tid < B) {
        const BrkSeg<Key> &sgt = buf.seg[0][tid];
        const int fl = sgt.flags;
        if ((fl & SEG_USED) && !(fl & (SEG_FINAL | SEG_TIE))) {
            atomicOr(&actmask[tid >> 5], 1u << (tid & 31));
            my_cnt = sgt.cnt_local > 0 ? sgt.cnt_local : 1ull;
        }
    }
    __syncthreads();
    if (tid < B && (actmask[tid >> 5] >> (tid & 31) & 1u)) {
        int pos = __popc(actmask[tid >> 5] & ((1u << (tid & 31)) - 1u));
        for (int wd = 0; wd < (tid >> 5); ++wd) pos += __popc(actmask[wd]);
        act[pos] = (unsigned char)tid;
        acnt[pos] = my_cnt;
    }
    if (tid == 0) n_act = __popc(actmask[0]) + __popc(actmask[1]) + __popc(actmask[2]) + __popc(actmask[3]);
    __syncthreads();
    if (tid == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < n_act; ++i) {
            apre[i] = run;
            run += acnt[i];
        }
        apre[n_act] = run;
    }
    __syncthreads();
    const int nact = n_act;
    if (nact == 0) return;
    const unsigned long long keys = apre[nact];
    // position k of the sequence -> (active bracket, piece)
    auto locate = [&](unsigned long long k, int &ai, unsigned &w) {
        if (k >= keys) { ai = nact; w = 0; return; }
        int lo = 0, hi = nact;  // last ai with apre[ai] <= k
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (apre[mid] <= k) lo = mid;
            else hi = mid;
        }
        ai = lo;
        const unsigned long long inb = k - apre[lo];
        unsigned long long ww = inb * nw / acnt[lo];  // (inb < acnt <= 2^40, nw <= 148: no overflow)
        w = (unsigned)(ww < nw ? ww : nw - 1);
    };
    int ai0, ai1;
    unsigned w0, w1;
    locate(keys * blockIdx.x / gridDim.x, ai0, w0);
    locate(keys * (blockIdx.x + 1ull) / gridDim.x, ai1, w1);
    unsigned char *myctr = counters + (warp >> 2) * 128 + lane * 4 + (warp & 3);  // conflict-free (see the streaming pass)
    BrkSeg<Key> *nsegs = buf.seg[1];
    Key *dst = buf.cand[1];
    for (int ai = ai0; ai <= ai1 && ai < nact; ++ai) {
        const int b = (int)act[ai];
        const unsigned w_begin = ai == ai0 ? w0 : 0u;
        const unsigned w_end = ai == ai1 ? w1 : nw;
        if (w_begin >= w_end) continue;
        const BrkSeg<Key> sg = buf.seg[0][b];
        if (COMPACT) {
            __syncthreads();
            if (tid == 0) n_wanted = 0;
            __syncthreads();
            const short id = buf.newid[b * 256 + tid];
            sid[tid] = id;
            if (id >= 0) wanted[atomicAdd(&n_wanted, 1)] = tid;
            __syncthreads();
            if (n_wanted == 0) continue;
        } else {
            __syncthreads();
            for (int i = tid; i < 256 * BRK_THREADS / 4; i += BRK_THREADS) reinterpret_cast<unsigned *>(counters)[i] = 0;
            __syncthreads();
        }
        const unsigned *fills = buf.fill + (size_t)b * nw;
        const Key *src = buf.cand[0] + sg.off;
        auto digit = [&](Key k) -> unsigned { return (unsigned)((Key)(k - sg.lo) >> sg.shift) & 0xffu; };
        auto keep = [&](Key k) {
            const int id = sid[digit(k)];
            if (id >= 0) {
                const unsigned long long pos = atomicAdd(&nsegs[id].cursor, 1ull);
                if (pos < nsegs[id].cnt_local) dst[nsegs[id].off + pos] = k;
            }
        };
        unsigned chunks_since_spill = 0;
        auto spill = [&]() {
            __syncthreads();
            for (int c = warp; c < 256; c += WARPS) {
                unsigned *row = reinterpret_cast<unsigned *>(counters + (size_t)c * BRK_THREADS);
                unsigned sum = 0;
#pragma unroll
                for (int i = 0; i < BRK_THREADS / 4 / 32; ++i) {
                    sum = __dp4a(row[i * 32 + lane], 0x01010101u, sum);
                    row[i * 32 + lane] = 0;
                }
                sum = __reduce_add_sync(0xffffffffu, sum);
                if (lane == 0 && sum) atomicAdd(&buf.loc_refine[(size_t)b * 256 + c], (unsigned long long)sum);
            }
            __syncthreads();
        };
        // ---- the chunks of this CTA's pieces as one stream: the next chunk is loaded while this one is processed ----
        __syncthreads();
        for (unsigned w = tid; w < nw; w += BRK_THREADS) {
            unsigned f = fills[w];
            sfill[w] = f > sg.piece_cap ? (unsigned)sg.piece_cap : f;
        }
        if (COMPACT)
            for (int j = tid; j < n_wanted && j < 4; j += BRK_THREADS) {
                const int id = sid[wanted[j]];
                wseg_off[j] = nsegs[id].off;
                wseg_cap[j] = nsegs[id].cnt_local;
            }
        __syncthreads();
        unsigned cw = w_begin, ce0 = 0;   // current chunk: piece cw, first key ce0
        auto seek = [&](unsigned &w, unsigned &e0) -> bool {  // move (w, e0) to the next non-empty chunk at or after it
            while (w < w_end && e0 >= sfill[w]) {
                w += 1u;
                e0 = 0;
            }
            return w < w_end;
        };
        // keys of chunk (w, e0) into k[]; bit e of the result is set when k[e] is a real key
        auto load = [&](unsigned w, unsigned e0, Key (&k)[KPT]) -> unsigned {
            constexpr int VEC = 16 / (int)sizeof(Key);
            const unsigned cnt = sfill[w] - e0 < (unsigned)BRK_RCHUNK ? sfill[w] - e0 : (unsigned)BRK_RCHUNK;
            const Key *ptr = src + (unsigned long long)w * sg.piece_stride + e0;  // 16-byte aligned
            unsigned valid = 0;
#pragma unroll
            for (int j = 0; j < KPT / VEC; ++j) {
                const unsigned idx = ((unsigned)j * BRK_THREADS + tid) * VEC;
                if (idx + VEC <= cnt) {
                    const uint4 q = __ldg(reinterpret_cast<const uint4 *>(ptr + idx));
                    if constexpr (sizeof(Key) == 4) {
                        k[j * 4] = q.x; k[j * 4 + 1] = q.y; k[j * 4 + 2] = q.z; k[j * 4 + 3] = q.w;
                    } else {
                        k[j * 2] = (Key)q.x | ((Key)q.y << 32);
                        k[j * 2 + 1] = (Key)q.z | ((Key)q.w << 32);
                    }
                    valid |= ((1u << VEC) - 1u) << (j * VEC);
                } else {
#pragma unroll
                    for (int u = 0; u < VEC; ++u) {
                        k[j * VEC + u] = 0;
                        if (idx + u < cnt) {
                            k[j * VEC + u] = ptr[idx + u];
                            valid |= 1u << (j * VEC + u);
                        }
                    }
                }
            }
            return valid;
        };
        auto process = [&](const Key (&k)[KPT], unsigned valid) {
            if (COMPACT) {
                // one global atomic per warp, chunk and wanted digit instead of one per kept key (their latency would
                // serialise the warp)
                const int nwd = n_wanted;
                if (nwd > 4) {
#pragma unroll
                    for (int e = 0; e < KPT; ++e)
                        if (valid & (1u << e)) keep(k[e]);
                    return;
                }
                for (int j = 0; j < nwd; ++j) {
                    const unsigned wd = (unsigned)wanted[j];
                    unsigned m = 0;
#pragma unroll
                    for (int e = 0; e < KPT; ++e) m |= (digit(k[e]) == wd ? 1u : 0u) << e;
                    m &= valid;
                    const unsigned cnt = __popc(m);
                    unsigned incl = cnt;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += t;
                    }
                    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
                    if (total == 0) continue;
                    unsigned long long base = 0;
                    if (lane == 31) base = atomicAdd(&nsegs[sid[wd]].cursor, (unsigned long long)total);
                    base = __shfl_sync(0xffffffffu, base, 31);
                    unsigned long long pos = base + incl - cnt;
                    const unsigned long long cap = wseg_cap[j], off = wseg_off[j];
#pragma unroll
                    for (int e = 0; e < KPT; ++e) {
                        if (m & (1u << e)) {
                            if (pos < cap) dst[off + pos] = k[e];
                            ++pos;
                        }
                    }
                }
            } else {
                // Partial chunks as well (the last chunk of every piece; the pieces of a lane sharded over eight ranks are
                // shorter than one chunk): a key that is not there counts in a spare 257th row.  The key-by-key path this
                // replaces ran at a sixth of this one's rate.
#pragma unroll
                for (int g = 0; g < KPT; g += 4) {  // 4 loads, equal digits merged, 4 stores (see the count kernel)
                    const unsigned c0 = (valid >> g) & 1u ? digit(k[g]) : 256u, c1 = (valid >> (g + 1)) & 1u ? digit(k[g + 1]) : 256u;
                    const unsigned c2 = (valid >> (g + 2)) & 1u ? digit(k[g + 2]) : 256u, c3 = (valid >> (g + 3)) & 1u ? digit(k[g + 3]) : 256u;
                    unsigned char *p0 = myctr + c0 * BRK_THREADS, *p1 = myctr + c1 * BRK_THREADS,
                                  *p2 = myctr + c2 * BRK_THREADS, *p3 = myctr + c3 * BRK_THREADS;
                    const unsigned x0 = *p0, x1 = *p1, x2 = *p2, x3 = *p3;
                    const unsigned m0 = 1u + (c1 == c0) + (c2 == c0) + (c3 == c0), m1 = 1u + (c0 == c1) + (c2 == c1) + (c3 == c1);
                    const unsigned m2 = 1u + (c0 == c2) + (c1 == c2) + (c3 == c2), m3 = 1u + (c0 == c3) + (c1 == c3) + (c2 == c3);
                    *p0 = (unsigned char)(x0 + m0);
                    *p1 = (unsigned char)(x1 + m1);
                    *p2 = (unsigned char)(x2 + m2);
                    *p3 = (unsigned char)(x3 + m3);
                }
            }
        };
        if (seek(cw, ce0)) {
            Key ka[KPT], kb[KPT];
            unsigned va = load(cw, ce0, ka), vb = 0;
            for (;;) {
                unsigned nw_ = cw, ne0 = ce0 + BRK_RCHUNK;
                bool more = seek(nw_, ne0);
                if (more) vb = load(nw_, ne0, kb);
                process(ka, va);
                if (!COMPACT && ++chunks_since_spill >= 255 / KPT) {  // 8-bit counters; CTA-uniform (the chunk sequence is)
                    spill();
                    chunks_since_spill = 0;
                }
                if (!more) break;
                cw = nw_; ce0 = ne0;
                nw_ = cw; ne0 = ce0 + BRK_RCHUNK;
                more = seek(nw_, ne0);
                if (more) va = load(nw_, ne0, ka);
                process(kb, vb);
                if (!COMPACT && ++chunks_since_spill >= 255 / KPT) {
                    spill();
                    chunks_since_spill = 0;
                }
                if (!more) break;
                cw = nw_; ce0 = ne0;
            }
        }
        if (!COMPACT) spill();
    }
}

// One CTA: every active task picks its digit; the chosen (segment, digit) pairs become the next segments.
template <typename Key>
__global__ void __launch_bounds__(1024) brk_refine_decide_kernel(BrkBuf<Key> buf) {
    __shared__ int want[BRK_MAXTASK];                 // seg * 256 + digit of the task, or -1
    __shared__ int newseg_of[BRK_MAXTASK];            // id of the segment the task moves to
    __shared__ unsigned long long pad_cnt[BRK_MAXTASK];
    __shared__ unsigned char first[BRK_MAXTASK];
    __shared__ unsigned char seg_any[BRK_MAXSEG];
    __shared__ unsigned long long new_cnt_local[BRK_MAXSEG], old_cnt_local[BRK_MAXSEG];
    __shared__ unsigned char new_final[BRK_MAXSEG];
    __shared__ int s_nid;
    // the tasks and what they need of their segments, staged once by all threads: a warp (or a single thread) that
    // walks the tables in global memory pays one memory latency per entry, and this kernel sits on the critical path
    // of every refinement round
    __shared__ BrkTask s_task[BRK_MAXTASK];
    __shared__ Key s_lo[BRK_MAXTASK];
    __shared__ int s_shift[BRK_MAXTASK], s_flags[BRK_MAXTASK];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    if (buf.red_refine != buf.loc_refine && buf.hdr[H_GATE_ROUND] && buf.red_refine_fail[0]) {
        if (tid == 0) raise_fail(buf.hdr, FAIL_REMOTE);  // some rank gave up: everybody does
        return;
    }
    if (buf.hdr[H_FAIL]) return;
    if (buf.hdr[H_NACTIVE] == 0) {
        if (tid == 0) {
            buf.hdr[H_NCHUNK_COMPACT] = 0;
            buf.hdr[H_DID] = 0;
            buf.hdr[H_NFINAL] = 0;
        }
        return;
    }
    const int cur = (int)buf.hdr[H_CUR];
    const int nseg = (int)buf.hdr[H_NSEG], ntask = (int)buf.hdr[H_NTASK];
    const BrkSeg<Key> *segs = buf.seg[cur];
    BrkSeg<Key> *nsegs = buf.seg[cur ^ 1];
    for (int i = tid; i < nseg * 256; i += blockDim.x) buf.newid[i] = -1;
    for (int i = tid; i < BRK_MAXTASK; i += blockDim.x) { want[i] = -1; newseg_of[i] = -1; pad_cnt[i] = 0; first[i] = 0; }
    for (int i = tid; i < BRK_MAXSEG; i += blockDim.x) seg_any[i] = 0;
    for (int t = tid; t < ntask; t += blockDim.x) {
        const BrkTask tk = buf.task[t];
        s_task[t] = tk;
        s_flags[t] = SEG_FINAL;  // "nothing to decide"
        if (tk.state == TASK_ACTIVE && tk.seg >= 0) {
            const BrkSeg<Key> sg = segs[tk.seg];
            s_flags[t] = sg.flags;
            s_shift[t] = sg.shift;
            s_lo[t] = sg.lo;
        }
    }
    __syncthreads();
    // ---- digit of every active task --------------------------------------------------------------------
    for (int t = warp; t < ntask; t += nwarps) {
        BrkTask tk = s_task[t];
        if (tk.state != TASK_ACTIVE) continue;
        struct { Key lo; int shift, flags; } sg = {s_lo[t], s_shift[t], s_flags[t]};
        if (sg.flags & (SEG_FINAL | SEG_TIE)) continue;  // resolved by the final kernel
        const unsigned long long *row = buf.red_refine + (size_t)tk.seg * 256;
        unsigned long long c[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            c[j] = row[lane * 8 + j];
            sum += c[j];
        }
        unsigned long long incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        unsigned long long run = incl - sum;
        int digit = -1;
        unsigned long long rem = 0;
        if (tk.r >= run && tk.r < incl) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (tk.r >= run && tk.r < run + c[j]) {
                    digit = lane * 8 + j;
                    rem = tk.r - run;
                }
                run += c[j];
            }
        }
        const unsigned hit = __ballot_sync(0xffffffffu, digit >= 0);
        if (hit == 0) {  // the rank is not inside this segment: the counts are inconsistent
            if (lane == 0) raise_fail(buf.hdr, FAIL_INCONSISTENT);
            continue;
        }
        const int srcl = __ffs(hit) - 1;
        digit = __shfl_sync(0xffffffffu, digit, srcl);
        rem = __shfl_sync(0xffffffffu, rem, srcl);
        if (lane == 0) {
            if (sg.shift == 0) {  // the digit is the rest of the key
                tk.state = TASK_DONE;
                tk.value = (unsigned long long)(Key)(sg.lo + (Key)digit);
            } else {
                want[t] = tk.seg * 256 + digit;
                pad_cnt[t] = (buf.loc_refine[(size_t)tk.seg * 256 + digit] + 3) & ~3ull;
                tk.r = rem;
            }
            s_task[t] = tk;
        }
    }
    __syncthreads();
    // ---- new segments: one per distinct wanted (segment, digit), numbered in (segment, digit) order ------
    for (int t = tid; t < ntask; t += blockDim.x) {
        const int w = want[t];
        if (w < 0) continue;
        bool f = true;
        for (int u = 0; u < t; ++u)
            if (want[u] == w) { f = false; break; }
        first[t] = f ? 1 : 0;
    }
    if (tid == 0) s_nid = 0;
    __syncthreads();
    for (int t = tid; t < ntask; t += blockDim.x) {
        const int w = want[t];
        if (w < 0) continue;
        int id = 0;
        unsigned long long off = 0;
        for (int u = 0; u < ntask; ++u) {
            if (first[u] && want[u] >= 0 && want[u] < w) {
                ++id;
                off += pad_cnt[u];
            }
        }
        newseg_of[t] = id;
        if (first[t]) {
            const int s = w >> 8, d = w & 255;
            struct { Key lo; int shift; } sg = {s_lo[t], s_shift[t]};  // (task t still points at its old segment s)
            BrkSeg<Key> ns;
            ns.lo = sg.lo + ((Key)d << sg.shift);
            ns.hi = ns.lo + (((Key)1 << sg.shift) - 1);
            ns.shift = sg.shift > 8 ? sg.shift - 8 : 0;
            ns.flags = SEG_USED;
            ns.off = off;
            ns.cnt_local = buf.loc_refine[(size_t)w];
            ns.cnt_global = buf.red_refine[(size_t)w];
            ns.cursor = 0;
            ns.piece_cap = 0;
            ns.piece_stride = 0;
            if (ns.cnt_global <= buf.hdr[H_FINALCAP]) ns.flags |= SEG_FINAL;
            if (off + pad_cnt[t] > buf.cand_cap[cur ^ 1]) raise_fail(buf.hdr, FAIL_CAPACITY);
            nsegs[id] = ns;
            new_cnt_local[id] = ns.cnt_local;
            new_final[id] = (ns.flags & SEG_FINAL) ? 1 : 0;
            buf.newid[w] = (short)id;
            seg_any[s] = 1;
            atomicAdd(&s_nid, 1);
        }
    }
    __syncthreads();
    // ---- re-point the tasks; chunk tables of the compaction pass (old segments) and the next count pass ----
    for (int t = tid; t < ntask; t += blockDim.x) {
        BrkTask tk = s_task[t];  // (tasks this round did not touch are written back unchanged)
        if (want[t] >= 0) tk.seg = newseg_of[t];
        buf.task[t] = tk;
    }
    for (int s = tid; s < nseg; s += blockDim.x) old_cnt_local[s] = seg_any[s] ? segs[s].cnt_local : 0ull;
    __syncthreads();
    if (tid == 0) {
        const int nid = s_nid;
        unsigned long long chunks_compact = 0, chunks_count = 0;
        for (int s = 0; s < nseg; ++s) {
            buf.cprefix_compact[s] = chunks_compact;
            if (seg_any[s]) chunks_compact += (old_cnt_local[s] + BRK_RCHUNK - 1) / BRK_RCHUNK;
        }
        for (int s = nseg; s <= BRK_MAXSEG; ++s) buf.cprefix_compact[s] = chunks_compact;
        int active = 0, nfinal = 0;
        for (int s = 0; s < nid; ++s) {
            buf.cprefix_count[s] = chunks_count;
            if (!new_final[s]) {
                chunks_count += (new_cnt_local[s] + BRK_RCHUNK - 1) / BRK_RCHUNK;
                ++active;
            } else {
                ++nfinal;
            }
        }
        buf.hdr[H_NFINAL] = (unsigned long long)nfinal;
        for (int s = nid; s <= BRK_MAXSEG; ++s) buf.cprefix_count[s] = chunks_count;
        buf.hdr[H_NCHUNK_COMPACT] = chunks_compact;
        // the next count pass must see the new table only after the compaction of this round has read the
        // old one: brk_refine_advance_kernel publishes these
        buf.hdr[H_NEXT_NCHUNK] = chunks_count;
        buf.hdr[H_NEXT_NSEG] = (unsigned long long)nid;
        buf.hdr[H_OLD_NSEG] = (unsigned long long)nseg;
        buf.hdr[H_NEXT_ACTIVE] = (unsigned long long)active;  // becomes H_NACTIVE once the compaction has run
        buf.hdr[H_DID] = 1;
    }
    // the counts of this round are consumed: clear them for the next one
    __syncthreads();
    for (int i = tid; i < nseg * 256; i += blockDim.x) {
        buf.loc_refine[i] = 0;
        if (buf.red_refine != buf.loc_refine) buf.red_refine[i] = 0;
    }
}

template <typename Key>
__global__ void __launch_bounds__(BRK_THREADS) brk_refine_compact_kernel(BrkBuf<Key> buf) {
    __shared__ unsigned long long sprefix[BRK_MAXSEG + 1];
    __shared__ short sid[256];
    if (failed_any(buf.hdr) || buf.hdr[H_ROUND] == 0 || !buf.hdr[H_DID]) return;
    const unsigned long long nchunks = buf.hdr[H_NCHUNK_COMPACT];
    if (nchunks == 0) return;
    const int nseg = (int)buf.hdr[H_OLD_NSEG];
    const int cur = (int)buf.hdr[H_CUR];
    const int tid = threadIdx.x;
    for (int i = tid; i <= nseg; i += BRK_THREADS) sprefix[i] = buf.cprefix_compact[i];
    __syncthreads();
    const unsigned long long per = (nchunks + gridDim.x - 1) / gridDim.x;
    const unsigned long long c0 = (unsigned long long)blockIdx.x * per;
    const unsigned long long c1 = c0 + per < nchunks ? c0 + per : nchunks;
    const Key *src = buf.cand[cur];
    Key *dst = buf.cand[cur ^ 1];
    const BrkSeg<Key> *segs = buf.seg[cur];
    BrkSeg<Key> *nsegs = buf.seg[cur ^ 1];
    int cur_seg = -1;
    for (unsigned long long ch = c0; ch < c1; ++ch) {
        // segments without a wanted digit have equal prefix entries: take the LAST s with prefix[s] <= ch
        // among those that own chunks
        int s = seg_of_chunk(sprefix, nseg, ch);
        if (s != cur_seg) {
            __syncthreads();
            sid[tid] = buf.newid[s * 256 + tid];
            __syncthreads();
            cur_seg = s;
        }
        const BrkSeg<Key> sg = segs[s];
        const unsigned long long e0 = (ch - sprefix[s]) * BRK_RCHUNK;
        const unsigned long long e1 = e0 + BRK_RCHUNK < sg.cnt_local ? e0 + BRK_RCHUNK : sg.cnt_local;
        const Key *sp = src + sg.off;
        auto keep = [&](Key k) {
            const unsigned c = (unsigned)((Key)(k - sg.lo) >> sg.shift) & 0xffu;
            const int id = sid[c];
            if (id >= 0) {
                const unsigned long long pos = atomicAdd(&nsegs[id].cursor, 1ull);
                if (pos < nsegs[id].cnt_local) dst[nsegs[id].off + pos] = k;
            }
        };
        if (e1 - e0 == BRK_RCHUNK) {
            Key k[BRK_RCHUNK / BRK_THREADS];
            load_chunk_keys<Key>(sp + e0, tid, k);
#pragma unroll
            for (int e = 0; e < BRK_RCHUNK / BRK_THREADS; ++e) keep(k[e]);
        } else {
            for (unsigned long long i = e0 + tid; i < e1; i += BRK_THREADS) keep(sp[i]);
        }
    }
}

// publish the next round's segment table sizes (after the compaction has read the old ones)
template <typename Key> __global__ void brk_refine_advance_kernel(BrkBuf<Key> buf) {
    if (threadIdx.x != 0) return;
    if (!buf.hdr[H_FAIL] && buf.hdr[H_DID]) {
        buf.hdr[H_NCHUNK_COUNT] = buf.hdr[H_NEXT_NCHUNK];
        buf.hdr[H_NSEG] = buf.hdr[H_NEXT_NSEG];
        buf.hdr[H_NACTIVE] = buf.hdr[H_NEXT_ACTIVE];
        buf.hdr[H_CUR] = buf.hdr[H_CUR] ^ 1ull;
        buf.hdr[H_ROUND] = buf.hdr[H_ROUND] + 1;
        buf.hdr[H_DID] = 0;
    }
    const bool ok = buf.hdr[H_FAIL] == 0;
    buf.hdr[H_GATE_ROUND] = (ok && buf.hdr[H_NACTIVE] != 0) ? 1ull : 0ull;
    buf.hdr[H_GATE_SMALL] = (ok && buf.hdr[H_NFINAL] != 0 && buf.hdr[H_MULTI]) ? 1ull : 0ull;
}

// Sharded lanes: the small (FINAL) segments that still have a task are packed into one send buffer per rank --
// header of BRK_MAXSEG counts, then the keys segment after segment -- and all-gathered, so that every rank can
// sort the union.  One CTA: the volume is a few thousand keys.
template <typename Key>
__global__ void __launch_bounds__(1024) brk_pack_final_kernel(BrkBuf<Key> buf) {
    __shared__ unsigned cnt[BRK_MAXSEG], off[BRK_MAXSEG];
    __shared__ unsigned char wanted[BRK_MAXSEG];
    unsigned *hdr_out = reinterpret_cast<unsigned *>(buf.gsend);
    Key *keys_out = reinterpret_cast<Key *>(buf.gsend + (size_t)BRK_MAXSEG * 4);
    const int tid = threadIdx.x;
    if (!buf.hdr[H_GATE_SMALL]) return;  // nothing to gather (the same on every rank): the exchange step is skipped, too
    for (int i = tid; i < BRK_MAXSEG; i += blockDim.x) { cnt[i] = 0; off[i] = 0; wanted[i] = 0; hdr_out[i] = 0; }
    if (tid == 0) buf.hdr[H_GWORDS] = (unsigned long long)BRK_MAXSEG * 4 / 8;  // the header alone
    __syncthreads();
    if (failed_any(buf.hdr) || buf.hdr[H_EMPTY]) return;
    const int nseg = (int)buf.hdr[H_NSEG], ntask = (int)buf.hdr[H_NTASK];
    const int cur = (int)buf.hdr[H_CUR];
    for (int t = tid; t < ntask; t += blockDim.x) {
        const BrkTask tk = buf.task[t];
        if (tk.state == TASK_ACTIVE && tk.seg >= 0) wanted[tk.seg] = 1;
    }
    __syncthreads();
    // sizes and offsets of the segments to send: one thread per segment (independent loads of the segment table -- a
    // single thread walking it pays one memory latency per segment), then a scan by the first warps
    __shared__ unsigned long long soff[BRK_MAXSEG], spstride[BRK_MAXSEG];
    __shared__ unsigned spcap[BRK_MAXSEG];
    __shared__ unsigned wsum[BRK_MAXSEG / 32];
    __shared__ int overflow;
    if (tid == 0) overflow = 0;
    if (tid < BRK_MAXSEG) {
        unsigned c = 0;
        if (tid < nseg && wanted[tid]) {
            const BrkSeg<Key> sg = buf.seg[cur][tid];
            if ((sg.flags & SEG_FINAL) && !(sg.flags & SEG_TIE)) {
                c = (unsigned)sg.cnt_local;
                soff[tid] = sg.off;
                spcap[tid] = (unsigned)sg.piece_cap;
                spstride[tid] = sg.piece_stride;
            }
        }
        cnt[tid] = c;
        unsigned incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
            if ((tid & 31) >= d) incl += v;
        }
        off[tid] = incl - c;
        if ((tid & 31) == 31) wsum[tid >> 5] = incl;
    }
    __syncthreads();
    if (tid < BRK_MAXSEG) {
        unsigned before = 0;
        for (int w = 0; w < (tid >> 5); ++w) before += wsum[w];
        off[tid] += before;
        if (cnt[tid] && off[tid] + cnt[tid] > (unsigned)BRK_GATHER_CAP) overflow = 1;  // cannot happen with FINAL_CAP * tasks <= GATHER_CAP
    }
    __syncthreads();
    if (overflow) {  // be safe: send nothing, every rank falls back
        if (tid == 0) raise_fail(buf.hdr, FAIL_CAPACITY);
        return;
    }
    if (tid < BRK_MAXSEG) hdr_out[tid] = cnt[tid];
    if (tid == 0) {
        unsigned run = 0;
        for (int w = 0; w < BRK_MAXSEG / 32; ++w) run += wsum[w];
        buf.hdr[H_GWORDS] = ((unsigned long long)BRK_MAXSEG * 4 + (unsigned long long)run * sizeof(Key) + 7) / 8;
    }
    // the keys: one warp per segment
    const unsigned nw = (unsigned)buf.hdr[H_NW];
    const int lane = tid & 31;
    for (int sidx = tid >> 5; sidx < nseg; sidx += (int)(blockDim.x >> 5)) {
        const unsigned c = cnt[sidx];
        if (c == 0) continue;
        const Key *src = buf.cand[cur] + soff[sidx];
        Key *dst = keys_out + off[sidx];
        if (spcap[sidx] == 0) {
            for (unsigned i = lane; i < c; i += 32) dst[i] = src[i];
        } else {  // a round-0 segment: one piece per streaming CTA (fills in buf.fill), laid end to end
            const unsigned *fills = buf.fill + (size_t)sidx * nw;
            unsigned before = 0;
            for (unsigned w = 0; w < nw; ++w) {
                const unsigned f = min(fills[w], spcap[sidx]);
                const Key *piece = src + (unsigned long long)w * spstride[sidx];
                for (unsigned i = lane; i < f; i += 32)
                    if (before + i < c) dst[before + i] = piece[i];
                before += f;
            }
        }
    }
}

// segments of at most FINAL_CAP keys: one CTA sorts the keys and serves every task that points at it.
// `nranks` > 0: the keys of a segment are the union of what every rank packed (brk_pack_final_kernel + all-gather).
template <typename Key>
__global__ void __launch_bounds__(256) brk_final_kernel(BrkBuf<Key> buf, int nranks) {
    __shared__ Key s[BRK_FINAL_CAP];
    __shared__ int tseg[BRK_MAXTASK];  // segment of every still active task, -1 otherwise
    __shared__ int s_any;
    __shared__ unsigned roff[BRK_MAX_RANKS], rcnt[BRK_MAX_RANKS];
    if (failed_any(buf.hdr) || buf.hdr[H_EMPTY]) return;
    if (nranks > 0 && !buf.hdr[H_GATE_SMALL]) return;  // nothing was gathered
    const int nseg = (int)buf.hdr[H_NSEG], ntask = (int)buf.hdr[H_NTASK];
    const int cur = (int)buf.hdr[H_CUR];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) s_any = 0;
    __syncthreads();
    for (int t = tid; t < BRK_MAXTASK; t += blockDim.x) {
        int sg = -1;
        if (t < ntask) {
            const BrkTask tk = buf.task[t];
            if (tk.state == TASK_ACTIVE) sg = tk.seg;
        }
        tseg[t] = sg;
        if (sg >= 0) s_any = 1;
    }
    __syncthreads();
    if (!s_any) return;
    for (int sidx = blockIdx.x; sidx < nseg; sidx += gridDim.x) {
        const BrkSeg<Key> sg = buf.seg[cur][sidx];
        if (!(sg.flags & SEG_FINAL) || (sg.flags & (SEG_TIE | SEG_DONE))) continue;
        bool wanted = false;
        for (int t = 0; t < ntask; ++t)
            if (tseg[t] == sidx) { wanted = true; break; }
        if (!wanted) continue;
        unsigned cnt = (unsigned)sg.cnt_local;
        if (nranks > 0) cnt = (unsigned)sg.cnt_global;
        unsigned np2 = 1;
        while (np2 < cnt) np2 <<= 1;
        const Key *src = buf.cand[cur] + sg.off;
        if (nranks > 0) {
            // every rank packed the same segments in the same order: the part of rank r starts after the counts of
            // the earlier segments in r's header
            const size_t stride = (size_t)BRK_MAXSEG * 4 + (size_t)BRK_GATHER_CAP * sizeof(Key);
            for (int r = tid >> 5; r < nranks; r += blockDim.x >> 5) {
                const unsigned *h = reinterpret_cast<const unsigned *>(buf.grecv + (size_t)r * stride);
                unsigned before = 0;
                for (int u = lane; u < sidx; u += 32) before += h[u];
                before = __reduce_add_sync(0xffffffffu, before);
                if (lane == 0) { roff[r] = before; rcnt[r] = h[sidx]; }
            }
            __syncthreads();
            unsigned base = 0, total = 0;
            for (int r = 0; r < nranks; ++r) total += rcnt[r];
            if (total != cnt) {
                if (tid == 0) raise_fail(buf.hdr, FAIL_INCONSISTENT);
                __syncthreads();
                continue;
            }
            for (int r = 0; r < nranks; ++r) {
                const Key *part = reinterpret_cast<const Key *>(buf.grecv + (size_t)r * stride + (size_t)BRK_MAXSEG * 4) + roff[r];
                for (unsigned i = tid; i < rcnt[r]; i += blockDim.x) s[base + i] = part[i];
                base += rcnt[r];
            }
            for (unsigned i = cnt + tid; i < np2; i += blockDim.x) s[i] = (Key) ~(Key)0;
        } else if (sg.piece_cap == 0) {
            for (unsigned i = tid; i < np2; i += blockDim.x) s[i] = i < cnt ? src[i] : (Key) ~(Key)0;
        } else {  // a round-0 segment: one piece per streaming CTA, gathered by the warps of this CTA
            const unsigned nw = (unsigned)buf.hdr[H_NW];
            const unsigned *fills = buf.fill + (size_t)sidx * nw;
            if (tid == 0) s_any = 0;   // reused as the gather cursor (every thread has read it above)
            __syncthreads();
            for (unsigned w = tid >> 5; w < nw; w += blockDim.x >> 5) {
                const unsigned f = fills[w];
                if (f == 0) continue;
                unsigned base = 0;
                if (lane == 0) base = (unsigned)atomicAdd(&s_any, (int)f);
                base = __shfl_sync(0xffffffffu, base, 0);
                const Key *piece = src + (unsigned long long)w * sg.piece_stride;
                for (unsigned i = lane; i < f; i += 32)
                    if (base + i < BRK_FINAL_CAP) s[base + i] = piece[i];
            }
            __syncthreads();
            for (unsigned i = cnt + tid; i < np2; i += blockDim.x) s[i] = (Key) ~(Key)0;
        }
        __syncthreads();
        for (unsigned k = 2; k <= np2; k <<= 1) {
            for (unsigned j = k >> 1; j > 0; j >>= 1) {
                for (unsigned i = tid; i < np2; i += blockDim.x) {
                    unsigned ixj = i ^ j;
                    if (ixj > i) {
                        Key a = s[i], b = s[ixj];
                        const bool up = (i & k) == 0;
                        if ((a > b) == up) { s[i] = b; s[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        }
        for (int t = tid; t < ntask; t += blockDim.x) {
            BrkTask tk = buf.task[t];
            if (tk.state == TASK_ACTIVE && tk.seg == sidx) {
                if (tk.r < cnt) {
                    tk.value = (unsigned long long)s[tk.r];
                    tk.state = TASK_DONE;
                    buf.task[t] = tk;
                } else {
                    raise_fail(buf.hdr, FAIL_INCONSISTENT);
                }
            }
        }
        __syncthreads();
    }
}

// 7. Interpolate (src/quantile/interpolate.rs:64-144) and write the result
template <typename T>
__global__ void brk_output_kernel(BrkBuf<typename Traits<T>::Key> buf, QuantArgs qa, T *out, uint8_t *out_valid,
                                  long long out_axis_stride, int last) {
    using Tr = Traits<T>;
    using Key = typename Tr::Key;
    if (failed_any(buf.hdr)) {  // the caller re-runs this lane on the radix passes (immediately, or at the next synchronize)
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(qa.status, STATUS_BIT_FALLBACK);
        return;
    }
    const unsigned long long n_eff = buf.hdr[H_NEFF];
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < qa.nq; t += gridDim.x * blockDim.x) {
        if (buf.hdr[H_EMPTY]) {
            out[(long long)t * out_axis_stride] = canonical_nan<T>();
            if (out_valid) out_valid[(long long)t * out_axis_stride] = 0;
            continue;
        }
        const BrkTask tl = buf.task[(int)buf.slot_rank[2 * t]], th = buf.task[(int)buf.slot_rank[2 * t + 1]];
        if (tl.state != TASK_DONE || th.state != TASK_DONE) {
            if (last == 2 || (last && buf.hdr[H_NACTIVE] == 0)) {  // (2: no further round will be enqueued)
                raise_fail(buf.hdr, FAIL_INCONSISTENT);
                atomicOr(qa.status, STATUS_BIT_FALLBACK);
            }
            continue;
        }
        RankMath rm = slot_rank_math(qa, t, n_eff);
        out[(long long)t * out_axis_stride] =
            interpolate<T>(qa.interp, Tr::from_key((Key)tl.value), Tr::from_key((Key)th.value), rm.frac, qa.status);
        if (out_valid) out_valid[(long long)t * out_axis_stride] = 1;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
template <typename Key> constexpr size_t gather_bytes() { return (size_t)BRK_MAXSEG * 4 + (size_t)BRK_GATHER_CAP * sizeof(Key); }

struct BrkSizes {
    unsigned s1;
    unsigned long long cap0, cap1;
    size_t total;
};

// Sample size: large enough that the brackets of all requested quantiles together hold ~12 % of the lane.
// wsum = sum over quantiles of 2 c sqrt(p (1 - p)) (the +-c sigma widths in units of sqrt(sample size)).
double sample_need(double wsum) {
    const double f_target = 0.12;
    double need = (wsum / f_target) * (wsum / f_target);
    if (const char *e = getenv("NDS_BRK_SAMPLE")) {
        const double v = atof(e);
        if (v >= 1.0) need = v;
    }
    return need;
}
unsigned pick_s1(unsigned long long n, double wsum) { return pick_s1_from_need(n, sample_need(wsum)); }
double predicted_fraction(unsigned long long n, int nq, double wsum) {
    const unsigned s1 = pick_s1(n, wsum);
    return s1 ? wsum / std::sqrt((double)s1) + (double)nq / 4096.0 : 1.0;
}

template <typename Key> BrkSizes brk_sizes(unsigned long long n, int nq, double wsum, bool sharded) {
    BrkSizes z;
    z.s1 = pick_s1(n, wsum);
    if (sharded) {  // the sample size follows the TOTAL length, known only after the first exchange: take the bound
        unsigned long long cap = n + 65536;
        if (cap > (1ull << 24)) cap = 1ull << 24;
        z.s1 = (unsigned)cap;
    }
    // candidates: up to ~0.3 n, plus the head room of the per-warp pieces (<= MAXB * 148 * 16 pieces)
    const unsigned long long pieces = std::min<unsigned long long>(148, n / STR_TILE + 1);  // streaming CTAs
    z.cap0 = n / 2 + (unsigned long long)BRK_MAXB * pieces * 2 * STR_TILE + (1ull << 16);
    z.cap1 = z.cap0 / 4 + (1ull << 16);
    size_t t = 0;
    t += align256(H_WORDS * 8) + align256(2 * 8) + align256(2 * (BRK_S0 + 2) * 8) + align256((BRK_NCLS + BRK_MAXB + 2) * 8) +
         align256(BRK_MAXB * 8) + 2 * align256(((size_t)BRK_MAXSEG * 256 + 2) * 8) + align256(16) + align256((size_t)BRK_MAXB * 148 * STR_WARPS * 4);
    t += 2 * align256(BRK_S0 * sizeof(Key)) + 2 * align256(BRK_MAXB * sizeof(Key)) +
         align256(16 + BRK_S0 * sizeof(Key)) * (1 + BRK_MAX_RANKS);
    t += align256(gather_bytes<Key>()) + align256(gather_bytes<Key>() * BRK_MAX_RANKS);
    t += 2 * align256(BRK_MAXB * 8) + align256(BRK_MAXB) + align256(BRK_NC1 * 4) + align256(BRK_L2N * 2) +
         align256(sizeof(BrkLutConst<Key>));
    t += 2 * align256(BRK_MAXSEG * sizeof(BrkSeg<Key>)) + align256(BRK_MAXTASK * sizeof(BrkTask)) +
         align256(2 * (size_t)nq * 4) + align256(2 * (size_t)nq * 8) + align256((size_t)BRK_MAXSEG * 256 * 2) +
         2 * align256((BRK_MAXSEG + 1) * 8);
    t += align256(z.cap0 * sizeof(Key)) + align256(z.cap1 * sizeof(Key));
    z.total = t;
    return z;
}

template <typename Key> BrkBuf<Key> brk_carve(void *scratch, const BrkSizes &z, int nq, size_t *zero_bytes) {
    BrkBuf<Key> b;
    char *p = (char *)scratch;
    auto take = [&](size_t bytes) {
        char *r = p;
        p += align256(bytes);
        return r;
    };
    // --- zeroed before every call (one memset) ---
    b.hdr = (unsigned long long *)take(H_WORDS * 8);
    b.red_n = (unsigned long long *)take(2 * 8);
    b.s0hist = (unsigned long long *)take(2 * (BRK_S0 + 2) * 8);
    b.red_stream = (unsigned long long *)take((BRK_NCLS + BRK_MAXB + 2) * 8);
    b.cur_local = (unsigned long long *)take(BRK_MAXB * 8);
    b.red_refine_fail = (unsigned long long *)take(((size_t)BRK_MAXSEG * 256 + 2) * 8);  // [fail word | digit counts]
    b.red_refine = b.red_refine_fail + 1;
    b.red_end = (unsigned long long *)take(2 * 8);
    b.loc_refine = (unsigned long long *)take((size_t)BRK_MAXSEG * 256 * 8);
    b.fill = (unsigned *)take((size_t)BRK_MAXB * 148 * STR_WARPS * 4);
    b.nw_cap = 148 * STR_WARPS;
    *zero_bytes = (size_t)(p - (char *)scratch);
    // --- the rest is written before it is read ---
    b.s0 = (Key *)take(BRK_S0 * sizeof(Key));
    b.s0all = (Key *)take(BRK_S0 * sizeof(Key));
    b.xsend = (unsigned long long *)take(16 + BRK_S0 * sizeof(Key));
    b.xrecv = (unsigned long long *)take((16 + BRK_S0 * sizeof(Key)) * BRK_MAX_RANKS);
    b.gsend = (unsigned char *)take(gather_bytes<Key>());
    b.grecv = (unsigned char *)take(gather_bytes<Key>() * BRK_MAX_RANKS);
    b.blo = (Key *)take(BRK_MAXB * sizeof(Key));
    b.bhi = (Key *)take(BRK_MAXB * sizeof(Key));
    b.boff = (unsigned long long *)take(BRK_MAXB * 8);
    b.bcap = (unsigned long long *)take(BRK_MAXB * 8);
    b.btie = (unsigned char *)take(BRK_MAXB);
    b.lut1 = (unsigned *)take(BRK_NC1 * 4);
    b.lut2 = (unsigned short *)take(BRK_L2N * 2);
    b.lutc = (BrkLutConst<Key> *)take(sizeof(BrkLutConst<Key>));
    b.seg[0] = (BrkSeg<Key> *)take(BRK_MAXSEG * sizeof(BrkSeg<Key>));
    b.seg[1] = (BrkSeg<Key> *)take(BRK_MAXSEG * sizeof(BrkSeg<Key>));
    b.task = (BrkTask *)take(BRK_MAXTASK * sizeof(BrkTask));
    b.slot_task = (int *)take(2 * (size_t)nq * 4);
    b.slot_rank = (unsigned long long *)take(2 * (size_t)nq * 8);
    b.newid = (short *)take((size_t)BRK_MAXSEG * 256 * 2);
    b.cprefix_count = (unsigned long long *)take((BRK_MAXSEG + 1) * 8);
    b.cprefix_compact = (unsigned long long *)take((BRK_MAXSEG + 1) * 8);
    b.cand[0] = (Key *)take(z.cap0 * sizeof(Key));
    b.cand[1] = (Key *)take(z.cap1 * sizeof(Key));
    b.cand_cap[0] = z.cap0;
    b.cand_cap[1] = z.cap1;
    b.s1cap = z.s1;
    return b;
}

}  // namespace
double brk_csigma() {
    if (const char *e = getenv("NDS_BRK_SIGMA")) {
        const double v = atof(e);
        if (v > 0.5 && v < 20.0) return v;
    }
    return 5.0;
}
namespace {

// Stage timer (NDS_BRK_TIMING=1): CUDA events between the stages of one call, printed after the call has finished.
struct BrkTimer {
    bool on = false;
    cudaStream_t st = nullptr;
    std::vector<std::pair<const char *, cudaEvent_t>> marks;
    void mark(const char *name) {
        if (!on) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, st);
        marks.emplace_back(name, ev);
    }
    void report(int rank) {
        if (!on) return;
        cudaStreamSynchronize(st);
        std::string line = "[bracket-timing] rank " + std::to_string(rank) + ":";
        for (size_t i = 1; i < marks.size(); ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
            char b[96];
            std::snprintf(b, sizeof b, " %s %.1fus", marks[i].first, ms * 1000.f);
            line += b;
        }
        if (marks.size() > 1) {
            float ms = 0;
            cudaEventElapsedTime(&ms, marks.front().second, marks.back().second);
            char b[64];
            std::snprintf(b, sizeof b, " | total %.1fus", ms * 1000.f);
            line += b;
        }
        fprintf(stderr, "%s\n", line.c_str());
        for (auto &m : marks) cudaEventDestroy(m.second);
        marks.clear();
    }
};

template <typename T>
cudaError_t launch_bracket_lane(nds_ctx *ctx, const T *lane, unsigned long long n, const QuantArgs &qa, T *out,
                                uint8_t *out_valid, long long out_axis_stride, void *scratch, bool sharded,
                                bool synchronous, double wsum, int *failed, unsigned long long *n_total_out) {
    using Key = typename Traits<T>::Key;
    constexpr bool IS_FLOAT = Traits<T>::is_float;
    cudaStream_t st = ctx->stream;
    const bool multi = sharded && ctx->nccl_comm && ctx->nranks > 1;
    const bool p2p = multi && xchg_ready(ctx);   // our own exchange kernels over peer memory; otherwise NCCL collectives
    const BrkSizes z = brk_sizes<Key>(n, qa.nq, wsum, sharded);
    size_t zero_bytes = 0;
    BrkBuf<Key> buf = brk_carve<Key>(scratch, z, qa.nq, &zero_bytes);
    if (!multi) buf.red_refine = buf.loc_refine;
    cudaError_t e = cudaMemsetAsync(scratch, 0, zero_bytes, st);
    if (e != cudaSuccess) return e;
    const int sm = ctx->sm_count;
    static const bool timing_knob = getenv("NDS_BRK_TIMING") && atoi(getenv("NDS_BRK_TIMING")) != 0;
    BrkTimer timer;
    timer.on = timing_knob && synchronous;
    timer.st = st;
    timer.mark("start");
    auto count = [&]() { ctx->launches += 1; return cudaGetLastError(); };
    // The exchange steps of a lane sharded over ranks.  `dev_words` / `gate` (device words holding the same value on every
    // rank) size / skip a step on the device; NCCL cannot do either and moves `words` unconditionally (harmless: the
    // kernels that consume the result look at the same words).
    auto allreduce = [&](unsigned long long *p, size_t words, const unsigned long long *dev_words,
                         const unsigned long long *gate) -> cudaError_t {
        if (!multi) return cudaSuccess;
        if (p2p) return xchg_allreduce_u64(ctx, p, words, dev_words, gate);
        std::string err;
        ctx->xchg_steps += 1;
        if (nccl_allreduce_sum_u64(ctx, p, words, st, &err) != 0) {
            ctx->last_error = err;
            return cudaErrorUnknown;
        }
        return cudaSuccess;
    };
    auto allgather = [&](const void *send, void *recv, size_t bytes_per_rank, const unsigned long long *dev_words,
                         const unsigned long long *gate) -> cudaError_t {
        if (p2p) return xchg_allgather(ctx, send, recv, bytes_per_rank, dev_words, gate);
        std::string err;
        ctx->xchg_steps += 1;
        if (nccl_allgather_bytes(ctx, send, recv, bytes_per_rank, st, &err) != 0) {
            ctx->last_error = err;
            return cudaErrorUnknown;
        }
        return cudaSuccess;
    };
    auto read_hdr = [&](unsigned long long (&h)[H_WORDS]) -> cudaError_t {
        cudaError_t e2 = cudaMemcpyAsync(h, buf.hdr, sizeof h, cudaMemcpyDeviceToHost, st);
        if (e2 != cudaSuccess) return e2;
        if (sharded) ctx->host_syncs += 1;
        return cudaStreamSynchronize(st);
    };

    // small segments: sorted by one CTA; over ranks: packed, all-gathered, sorted by every rank
    auto finish_small = [&]() -> cudaError_t {
        cudaError_t e2;
        if (multi) {
            brk_pack_final_kernel<Key><<<1, 1024, 0, st>>>(buf);
            if ((e2 = count()) != cudaSuccess) return e2;
            if ((e2 = allgather(buf.gsend, buf.grecv, gather_bytes<Key>(), &buf.hdr[H_GWORDS], &buf.hdr[H_GATE_SMALL])) != cudaSuccess)
                return e2;
        }
        brk_final_kernel<Key><<<sm * 2, 256, 0, st>>>(buf, multi ? ctx->nranks : 0);   // up to 2 x 128 segments: one each
        return count();
    };
    // 0. threshold below which a segment is sorted by one CTA (sharded: small enough that all tasks' segments fit one
    //    rank's send buffer); sample size of an unsharded lane
    unsigned final_cap = multi ? (unsigned)std::min<int>(BRK_FINAL_CAP, BRK_GATHER_CAP / (2 * std::max(qa.nq, 1))) : (unsigned)BRK_FINAL_CAP;
    if (const char *fc = getenv("NDS_BRK_FINALCAP")) final_cap = std::min<unsigned>((unsigned)atoi(fc), BRK_FINAL_CAP);  // tests
    const double need = sample_need(wsum);
    const unsigned s1_host = sharded ? 0u : z.s1;
    brk_begin_kernel<Key><<<1, 32, 0, st>>>(buf, n, final_cap, multi ? 1 : 0, s1_host);
    if ((e = count()) != cudaSuccess) return e;
    // 1. over ranks: ONE exchange step carries every rank's element count and its share of the common splitters S0
    const Key *gathered = nullptr;
    if (multi) {
        const unsigned share = BRK_S0 / (unsigned)ctx->nranks;
        brk_pick_local_s0_kernel<T><<<32, 256, 0, st>>>(lane, n, buf, share);
        if ((e = count()) != cudaSuccess) return e;
        if ((e = allgather(buf.xsend, buf.xrecv, 16 + (size_t)share * sizeof(Key), nullptr, nullptr)) != cudaSuccess) return e;
        brk_unpack_s0_kernel<Key><<<8, 256, 0, st>>>(buf, ctx->nranks, share);
        if ((e = count()) != cudaSuccess) return e;
        gathered = buf.s0all;
        timer.mark("count+S0-exchange");
    }
    // 2. the splitters S0 (one rank: 8192 keys at hashed, evenly spread positions of the lane; over ranks: the gathered
    //    shares), sorted; then ONE kernel samples the lane (every rank at the same rate: the size follows on the device from
    //    the total length) and ranks every sampled key against S0
    {
        if (!multi) {
            brk_pick_local_s0_kernel<T><<<32, 256, 0, st>>>(lane, n, buf, (unsigned)BRK_S0);
            if ((e = count()) != cudaSuccess) return e;
            gathered = reinterpret_cast<const Key *>(buf.xsend + 2);
        }
        auto k = brk_sort_s0_kernel<Key>;
        const size_t smem = BRK_S0 * sizeof(Key);
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k<<<SORT_CTAS, SORT_THREADS, smem, st>>>(buf, gathered);
        if ((e = count()) != cudaSuccess) return e;
        timer.mark("sort-S0");
        const unsigned bound = sharded ? z.s1 : s1_host;
        auto ks = brk_sample_bucket_kernel<T>;
        const size_t smems = BRK_S0 * sizeof(Key) + (BRK_BLUT + 1) * 2 + 16;
        if ((e = cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smems)) != cudaSuccess) return e;
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ks, 512, smems);
        const unsigned grid = std::max(1u, std::min<unsigned>((bound + 511) / 512, (unsigned)(sm * std::max(1, occ))));
        ks<<<grid, 512, smems, st>>>(lane, n, s1_host, need, sharded ? 1 : 0, buf);
        if ((e = count()) != cudaSuccess) return e;
        timer.mark("sample+bucket");
    }
    if ((e = allreduce(buf.s0hist, 2 * (BRK_S0 + 2), nullptr, nullptr)) != cudaSuccess) return e;
    if (multi) timer.mark("hist-exchange");
    // 3. plan
    const unsigned long long tiles = n / STR_TILE + 1;
    const unsigned sgrid = (unsigned)std::min<unsigned long long>(tiles, (unsigned long long)std::min(sm, 148));
    const unsigned nw = sgrid;  // streaming CTAs = pieces per bracket segment
    {
        auto k = brk_plan_kernel<Key, IS_FLOAT>;
        const size_t smem = ((BRK_S0 + 1) * 4 + 15) / 16 * 16 + BRK_S0 * sizeof(Key) + BRK_NC1 * 2 + (BRK_S0 + 2) * 4;
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        static const bool lowcard_knob = !(getenv("NDS_BRK_LOWCARD") && atoi(getenv("NDS_BRK_LOWCARD")) == 0);
        k<<<1, 1024, smem, st>>>(buf, qa, brk_csigma(), n, multi ? 1 : 0, nw, lowcard_knob ? 1 : 0);
        if ((e = count()) != cudaSuccess) return e;
        timer.mark("plan");
    }
    // 4. stream (the kernel of the table mode the plan did not choose returns at once)
    {
        const size_t smem = stream_smem_bytes<Key>();
        auto k = brk_stream_kernel<T>;
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k<<<sgrid, STR_THREADS, smem, st>>>(lane, n, buf);
        if ((e = count()) != cudaSuccess) return e;
        brk_publish_fill_kernel<Key><<<1, 128, 0, st>>>(buf);
        if ((e = count()) != cudaSuccess) return e;
        timer.mark("stream");
    }
    if ((e = allreduce(buf.red_stream, BRK_NCLS + BRK_MAXB + 2, nullptr, nullptr)) != cudaSuccess) return e;
    if (multi) timer.mark("count-exchange");
    // 5. resolve
    brk_resolve_kernel<T><<<1, 256, 0, st>>>(buf, qa, qa.skipnan);
    if ((e = count()) != cudaSuccess) return e;
    timer.mark("resolve");
    // 6. refine.  Every decision is taken on the device from counts that are the same on every rank, so rounds are
    //    enqueued without looking: an idle round is a handful of kernels that return at once (over peer memory its
    //    exchange steps are skipped on the device, too).  One GPU: all possible rounds for an enqueued call, a look at
    //    the header after the second round of a synchronous one.  Over ranks: two rounds, then one look; more only when
    //    segments are still active (tie-heavy or adversarial lanes).
    unsigned long long hh[H_WORDS];
    unsigned long long agreed = 0;
    bool have_hdr = false;
    {
        auto kc = brk_refine_count_kernel<Key>;
        auto kc0 = brk_refine_round0_kernel<Key, false>;
        auto kp0 = brk_refine_round0_kernel<Key, true>;
        const size_t smem = (size_t)256 * BRK_THREADS * 2;
        if ((e = cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        const size_t smem0 = (size_t)257 * BRK_THREADS;  // 8-bit counters, one row per digit + the spare row of absent keys
        if ((e = cudaFuncSetAttribute(kc0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0)) != cudaSuccess) return e;
        const int max_rounds = (int)sizeof(Key) + 1;
        // whole waves: as many CTAs as are resident at once (3 per SM were launched where 2 fit: a second, half-empty wave)
        int occ_c0 = 2, occ_p0 = 4;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c0, kc0, BRK_THREADS, smem0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_p0, kp0, BRK_THREADS, 0);
        // (measured on C4: count 547 us at 2 CTAs per SM = one resident wave; compact 629 us at 4 per SM against 660 at 2)
        unsigned grid_c0 = (unsigned)(sm * std::max(1, occ_c0)), grid_p0 = (unsigned)(sm * 2 * std::max(1, occ_p0));
        if (const char *e = getenv("NDS_BRK_R0GRID")) {  // tuning: CTAs per SM of the count / compact pass of round 0
            int a = 0, b = 0;
            if (sscanf(e, "%d,%d", &a, &b) == 2 && a > 0 && b > 0) { grid_c0 = (unsigned)(sm * a); grid_p0 = (unsigned)(sm * b); }
        }
        if ((e = finish_small()) != cudaSuccess) return e;
        timer.mark("small-segments");
        auto one_round = [&](int round) -> cudaError_t {
            cudaError_t e2;
            if (round == 0) kc0<<<grid_c0, BRK_THREADS, smem0, st>>>(buf);
            else kc<<<sm, BRK_THREADS, smem, st>>>(buf);
            if ((e2 = count()) != cudaSuccess) return e2;
            timer.mark(round == 0 ? "r0:count" : "r:count");
            if (multi) {  // the exchange step: digit counts of every active segment, summed over ranks, behind the fail word
                brk_copy_counts_kernel<Key><<<64, 256, 0, st>>>(buf);
                if ((e2 = count()) != cudaSuccess) return e2;
                if ((e2 = allreduce(buf.red_refine_fail, 1 + (size_t)BRK_MAXSEG * 256, &buf.hdr[H_XWORDS], &buf.hdr[H_GATE_ROUND])) != cudaSuccess)
                    return e2;
                timer.mark("exch");
            }
            brk_refine_decide_kernel<Key><<<1, 1024, 0, st>>>(buf);
            if ((e2 = count()) != cudaSuccess) return e2;
            timer.mark("decide");
            if (round == 0) kp0<<<grid_p0, BRK_THREADS, 0, st>>>(buf);
            else brk_refine_compact_kernel<Key><<<sm * 2, BRK_THREADS, 0, st>>>(buf);
            if ((e2 = count()) != cudaSuccess) return e2;
            brk_refine_advance_kernel<Key><<<1, 32, 0, st>>>(buf);
            if ((e2 = count()) != cudaSuccess) return e2;
            timer.mark("compact");
            if ((e2 = finish_small()) != cudaSuccess) return e2;
            timer.mark("small");
            return cudaSuccess;
        };
        // 7. output (a slot whose ranks are not known yet is left alone while segments are still active)
        auto output = [&](int last) -> cudaError_t {
            brk_output_kernel<T><<<(qa.nq + 127) / 128, 128, 0, st>>>(buf, qa, out, out_valid, out_axis_stride, last);
            return count();
        };
        if (multi) {
            int round = 0;
            for (;;) {
                const int stop = std::min(round + 2, max_rounds);
                for (; round < stop; ++round)
                    if ((e = one_round(round)) != cudaSuccess) return e;
                if ((e = output(round >= max_rounds ? 2 : 1)) != cudaSuccess) return e;
                // agree on the outcome: either every rank keeps its result or all of them fall back
                brk_post_fail_kernel<Key><<<1, 32, 0, st>>>(buf, buf.red_end);
                if ((e = count()) != cudaSuccess) return e;
                if ((e = allreduce(buf.red_end, 2, nullptr, nullptr)) != cudaSuccess) return e;
                timer.mark("output+agree");
                if ((e = cudaMemcpyAsync(&agreed, buf.red_end, 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
                if ((e = read_hdr(hh)) != cudaSuccess) return e;
                have_hdr = true;
                // (H_FAIL, H_NACTIVE and the agreed word are the same on every rank: all of them leave together)
                if (hh[H_FAIL] || agreed || hh[H_NACTIVE] == 0 || round >= max_rounds) break;
            }
        } else {
            // an enqueued call cannot look: three rounds (24 key bits below the bracket width; the idle ones return at
            // once), and the rare lane whose segments are still active after them -- medium-sized runs of equal keys --
            // goes to the radix passes like any other give-up
            const int rounds = synchronous ? max_rounds : std::min(max_rounds, 3);
            for (int round = 0; round < rounds; ++round) {
                if ((e = one_round(round)) != cudaSuccess) return e;
                if (synchronous && round >= 1) {
                    if ((e = read_hdr(hh)) != cudaSuccess) return e;
                    if (hh[H_FAIL] || hh[H_NACTIVE] == 0) break;
                }
            }
            if ((e = output(2)) != cudaSuccess) return e;
            timer.mark("output");
        }
    }
    if (synchronous) {
        unsigned long long (&h)[H_WORDS] = hh;
        if (!have_hdr && (e = read_hdr(h)) != cudaSuccess) return e;
        *failed = (int)(h[H_FAIL] | h[H_LOCALFAIL]);
        if (n_total_out) *n_total_out = h[H_NTOTAL];
        if (multi && agreed) *failed |= FAIL_REMOTE;
        timer.report(ctx->rank);
        if (const char *dbg = getenv("NDS_BRK_DEBUG")) {
            if (atoi(dbg) >= 2) {  // brackets and exact class counts, for tools/dbg_bracket2.py
                const int B = (int)h[H_B];
                std::vector<Key> lo(BRK_MAXB), hi(BRK_MAXB);
                std::vector<unsigned long long> red(BRK_NCLS + BRK_MAXB);
                std::vector<unsigned char> tie(BRK_MAXB);
                cudaMemcpy(lo.data(), buf.blo, BRK_MAXB * sizeof(Key), cudaMemcpyDeviceToHost);
                cudaMemcpy(hi.data(), buf.bhi, BRK_MAXB * sizeof(Key), cudaMemcpyDeviceToHost);
                cudaMemcpy(tie.data(), buf.btie, BRK_MAXB, cudaMemcpyDeviceToHost);
                cudaMemcpy(red.data(), buf.red_stream, red.size() * 8, cudaMemcpyDeviceToHost);
                BrkLutConst<Key> lc;
                cudaMemcpy(&lc, buf.lutc, sizeof lc, cudaMemcpyDeviceToHost);
                fprintf(stderr, "[bracket-lut] kmin=%llu shift1=%d subshift=%d submask=%u hi32=%d\n", (unsigned long long)lc.kmin,
                        lc.shift1, lc.subshift, lc.submask, lc.hi32);
                for (int b = 0; b < B; ++b)
                    fprintf(stderr, "[bracket-b] %d lo=%llu hi=%llu tie=%d cls_count=%llu in=%llu\n", b,
                            (unsigned long long)lo[b], (unsigned long long)hi[b], (int)tie[b], red[b], red[BRK_NCLS + b]);
                fprintf(stderr, "[bracket-b] %d (above) cls_count=%llu\n", B, red[B]);
            }
            fprintf(stderr, "[bracket] rank=%d n=%llu n_total=%llu s1=%llu fail=%llu localfail=%llu B=%llu ties=%llu tasks=%llu n_eff=%llu invalid=%llu active=%llu round=%llu\n",
                    ctx->rank, n, h[H_NTOTAL], h[H_S1], h[H_FAIL], h[H_LOCALFAIL], h[H_B], h[H_NTIE], h[H_NTASK], h[H_NEFF], h[H_NINVALID],
                    h[H_NACTIVE], h[H_ROUND]);
        }
    } else {
        // asynchronous call: the FAIL word is folded into the ctx status at the next synchronize
        *failed = 0;
    }
    return cudaSuccess;
}

template <typename T>
cudaError_t launch_bracket_typed(nds_ctx *ctx, const DeviceCall &c, void *scratch, bool sharded, bool synchronous,
                                 int *failed, unsigned long long *n_total_out) {
    *failed = 0;
    for (unsigned long long lane = 0; lane < c.geom.nlanes; ++lane) {
        long long in_off, out_off;
        lane_offsets(c.geom, lane, in_off, out_off);
        int f = 0;
        cudaError_t e = launch_bracket_lane<T>(ctx, (const T *)c.data + in_off, c.geom.axis_len, c.q, (T *)c.out + out_off,
                                               c.out_valid ? c.out_valid + out_off : nullptr, c.geom.out_axis_stride,
                                               scratch, sharded, synchronous, c.brk_wsum, &f, n_total_out);
        if (e != cudaSuccess) return e;
        *failed |= f;
        if (f) break;
    }
    return cudaSuccess;
}

}  // namespace

bool bracket_supported(const nds_ctx *ctx, const DeviceCall &c, bool sharded, bool forced) {
    (void)ctx;
    // sharded over ranks: every rank contributes S0 / nranks keys of its sample to the common S0
    if (sharded && (ctx->nranks < 1 || ctx->nranks > BRK_MAX_RANKS || BRK_S0 % ctx->nranks != 0)) return false;
    switch (c.dtype) {
    case NDS_F32: case NDS_I32: case NDS_U32: case NDS_F64: case NDS_I64: case NDS_U64: break;
    default: return false;
    }
    if (c.valid || c.out_valid) return false;
    if (c.geom.axis_stride != 1 && c.geom.axis_len > 1) return false;
    if (c.q.nq < 1 || c.q.nq > BRK_MAXQ) return false;
    // (sharded: the decision must be the same on every rank, so it must not depend on the local shard length)
    if (!sharded && c.geom.axis_len < (1ull << 16)) return false;
    if (c.geom.axis_len >= (1ull << 40)) return false;
    // not worth trying when the brackets would hold more than the candidate scratch (n / 3): many quantiles of
    // a lane that is too short for a large enough sample go straight to the radix passes
    if (!forced && !sharded && predicted_fraction(c.geom.axis_len, c.q.nq, c.brk_wsum) > 0.28) return false;
    return true;
}

size_t bracket_scratch_bytes(const DeviceCall &c, bool sharded) {
    const size_t es = nds_dtype_size(c.dtype);
    if (es == 8) return brk_sizes<uint64_t>(c.geom.axis_len, c.q.nq, c.brk_wsum, sharded).total;
    return brk_sizes<uint32_t>(c.geom.axis_len, c.q.nq, c.brk_wsum, sharded).total;
}

cudaError_t launch_bracket(nds_ctx *ctx, const DeviceCall &c, void *scratch, bool sharded, bool synchronous,
                           int *failed, unsigned long long *n_total_out) {
    ctx->last_path = sharded ? "bracket_sharded" : "bracket";
    switch (c.dtype) {
    case NDS_F32: return launch_bracket_typed<float>(ctx, c, scratch, sharded, synchronous, failed, n_total_out);
    case NDS_F64: return launch_bracket_typed<double>(ctx, c, scratch, sharded, synchronous, failed, n_total_out);
    case NDS_I32: return launch_bracket_typed<int32_t>(ctx, c, scratch, sharded, synchronous, failed, n_total_out);
    case NDS_U32: return launch_bracket_typed<uint32_t>(ctx, c, scratch, sharded, synchronous, failed, n_total_out);
    case NDS_I64: return launch_bracket_typed<int64_t>(ctx, c, scratch, sharded, synchronous, failed, n_total_out);
    case NDS_U64: return launch_bracket_typed<uint64_t>(ctx, c, scratch, sharded, synchronous, failed, n_total_out);
    default: return cudaErrorNotSupported;
    }
}

}  // namespace nds
