// Lanes along a strided axis whose neighbours are contiguous ("column lanes", e.g. Axis(0) of a C-order
// matrix — BASELINE config 3): one CTA per strip of CW adjacent lanes.
//
// Reading one such lane alone touches one 4-byte element per 32-byte sector.  Here the coalescing is
// ACROSS lanes: a warp load covers 4 rows x (8 lanes x 4 B = one full sector), every thread collects 32
// consecutive rows of ONE lane in registers, transposes them into 16 bit-planes (nds_bitslice.cuh) and
// stores the plane words in shared memory (16 KiB per lane of 8192 elements, 8 lanes per CTA).  After the
// strip has been streamed once, warp w walks the key bits of lane w MSB-first exactly like the short-lane
// kernel (popc over the lane's plane words, one REDUX per bit), finishing on <= 32 candidates that are
// fetched again from L2.  Input is read from HBM once; skipnan masks come out of the exponent planes.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "nds_bitslice.cuh"
#include "nds_internal.h"

namespace nds {
namespace {

constexpr int COLS_CW_SYNC = 8;      // lanes per strip of the direct-load path: 8 x 4 B = one 32-byte sector per row (f32)
constexpr int COLS_KMAX = 8;         // plane words per thread in the walk: 8 x 32 x 32 = 8192 elements
constexpr unsigned FULLMASK = 0xffffffffu;
constexpr int COLS_CAND = 64;        // candidates ranked directly once the walk has narrowed a lane down to this many

constexpr int COLS_TILE_ROWS = 1024;                    // rows of a strip staged per cp.async group (4-byte types)
// one 32-row block of a staged tile is padded so that the row groups of a warp (4 for 8-lane strips, 8 for 4-lane
// strips) read 32 different banks
__host__ __device__ constexpr int cols_rb_words(int cw) { return 32 * cw + (cw == 8 ? 8 : 4); }
__host__ __device__ constexpr int cols_stage_words(int cw) { return 32 * cols_rb_words(cw); }  // 32 row blocks = 1024 rows
constexpr int COLS_STAGES = 2;
constexpr int COLS_CW_ASYNC = 8;   // lanes per strip of the staged path.  (4-lane strips -- 16 bytes per row, half the shared
                                   // memory, two CTAs per SM so that one streams while the other walks -- measured slower on
                                   // C3: 1.03 ms vs 0.91 ms; the two half-sector strips of a row do not pair up well enough in L2)

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// ASYNC (4-byte types, full aligned strips): the strip is staged through shared memory by cp.async (LDGSTS), two tiles
// of 1024 rows in flight, the first tiles of the NEXT strip already during the bit walk of this one.
template <typename T, bool ASYNC, int CW>
__global__ void __launch_bounds__(CW * 32, CW == 4 ? 2 : 1)
cols_bitslice_kernel(const T *__restrict__ data, LaneGeom g, QuantArgs qa, T *__restrict__ out,
                     unsigned long long n_strips, unsigned strips_per_group, unsigned pair_sync) {
    constexpr int COLS_CW = CW;                       // lanes per strip = warps per CTA
    constexpr int COLS_THREADS = CW * 32;             // CW lanes x 32 row groups
    constexpr int COLS_RB_WORDS = cols_rb_words(CW);
    constexpr int COLS_STAGE_WORDS = cols_stage_words(CW);
    using Tr = Traits<T>;
    using Key = typename Tr::Key;
    using Raw = typename RawBits<T>::U;
    constexpr int ES = (int)sizeof(T);
    constexpr int ROUNDS = ES * 8 / 16;
    constexpr int KEYBITS = ES * 8;
    constexpr bool IS_FLOAT = Tr::is_float;
    constexpr bool TOP_INVERTED = IS_FLOAT || std::is_signed<T>::value;
    constexpr int EXP_LO = (ES == 4) ? 7 : 4;  // exponent bits of the top 16-bit field are [EXP_LO, 14]
    // floats: two more plane words per 32-row block, made while the block is still in registers -- rows whose top 16 bits
    // say NaN (all-ones exponent, a high mantissa bit), and rows that are inf or a NaN with a low payload (decided by
    // looking at the element again).  The walk then needs two shared-memory reads per block for its skipnan masks.
    constexpr int NPL = IS_FLOAT ? 18 : 16;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *planes = reinterpret_cast<uint32_t *>(smem_raw);

    const unsigned n = (unsigned)g.axis_len;
    const long long pitch = g.axis_stride;
    const unsigned nrb = (n + 31u) / 32u;             // 32-row blocks per lane
    const unsigned nrbp = (nrb + 31u) / 32u * 32u;    // padded to a multiple of 32 (<= 256)
    const unsigned kw = nrbp / 32u;                   // plane words per thread in the walk
    const unsigned col_pitch = (unsigned)NPL * nrbp + 4u;  // +4: the 8 lanes of a strip hit different banks
    unsigned *clist = reinterpret_cast<unsigned *>(planes + COLS_CW * col_pitch);  // [warp][COLS_CAND]
    Key *ckeys = reinterpret_cast<Key *>(clist + COLS_CW * COLS_CAND);             // [warp][COLS_CAND] (spare: the survivors' keys stay in registers)
    // [COLS_STAGES][COLS_STAGE_WORDS], 16-byte aligned; derived from smem_raw by an OFFSET so that the compiler keeps
    // the shared address space (a pointer -> integer -> pointer round trip turns every access into a generic LD/ST)
    const size_t stage_off = ((size_t)(reinterpret_cast<unsigned char *>(ckeys + COLS_CW * COLS_CAND) - smem_raw) + 15) & ~(size_t)15;
    uint32_t *stage_base = reinterpret_cast<uint32_t *>(smem_raw + stage_off);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bc = tid % COLS_CW, brg = tid / COLS_CW;  // build role: lane (column) and row group
    const int last_outer = g.n_outer - 1;
    const unsigned long long ncols = g.oshape[last_outer];
    const long long out_col_stride = g.out_ostride[last_outer];
    auto slot_of_pos = [](int p) -> int { return 2 * (p & 15) + (p >> 4); };

    // ---- cp.async tile stream of this CTA: tile gt = (iteration, tile of the strip) ----------------------------
    const unsigned tiles = (n + COLS_TILE_ROWS - 1) / COLS_TILE_ROWS;
    // issue state: tile iss_t of strip iss_strip goes to stage iss_stage (no divisions on this path)
    unsigned iss_t = 0, iss_stage = 0;
    unsigned long long iss_strip = blockIdx.x;
    const T *iss_src = nullptr, *iss_ptr = nullptr;  // strip origin; this thread's first chunk of the tile to issue
    auto issue_tile = [&]() {
        if constexpr (ASYNC) {
            if (iss_strip < n_strips) {
                if (iss_t == 0) {  // first tile of a strip: where it starts
                    const unsigned long long grp_i = iss_strip / strips_per_group;
                    const unsigned long long col_i = (iss_strip - grp_i * strips_per_group) * COLS_CW;
                    long long io, oo;
                    lane_offsets(g, grp_i * ncols + col_i, io, oo);
                    iss_src = data + io;
                }
                uint32_t *st = stage_base + (size_t)iss_stage * COLS_STAGE_WORDS;
                // chunk q = j * 256 + tid of the tile: row q >> 1, first / second four lanes q & 1.  j moves the row by 128
                // = four 32-row blocks, so source and destination advance by constants (one 64-bit add per chunk; the
                // row * pitch products are formed once per strip)
                constexpr unsigned CPR = COLS_CW * 4 / 16;            // 16-byte chunks per row (2 or 1)
                constexpr unsigned RSTEP = COLS_THREADS / CPR;         // rows between a thread's chunks: 128
                static_assert(RSTEP == 128, "a thread's chunks are four 32-row blocks apart");
                const unsigned rl0 = (unsigned)tid / CPR, half = (unsigned)tid % CPR;
                if (iss_t == 0) iss_ptr = iss_src + (long long)rl0 * pitch + half * 4;
                const unsigned row0 = iss_t * COLS_TILE_ROWS + rl0;
                uint32_t *dst = st + (rl0 >> 5) * COLS_RB_WORDS + (rl0 & 31u) * COLS_CW + half * 4;
                const long long src_step = (long long)RSTEP * pitch;
                const T *s = iss_ptr;
                if ((iss_t + 1u) * COLS_TILE_ROWS <= n) {  // a whole tile (uniform): no row checks
#pragma unroll
                    for (int j = 0; j < COLS_TILE_ROWS / (int)RSTEP; ++j, s += src_step) cp_async16(dst + j * 4 * COLS_RB_WORDS, s);
                } else {
#pragma unroll
                    for (int j = 0; j < COLS_TILE_ROWS / (int)RSTEP; ++j, s += src_step)  // 8 chunks of 16 bytes per thread
                        if (row0 + (unsigned)j * RSTEP < n) cp_async16(dst + j * 4 * COLS_RB_WORDS, s);
                }
                iss_ptr = s;  // first chunk of the next tile
                if (++iss_t == tiles) {
                    iss_t = 0;
                    iss_strip += gridDim.x;
                }
            }
            iss_stage = (iss_stage + 1 == COLS_STAGES) ? 0 : iss_stage + 1;
            cp_async_commit();  // (an empty group past the end keeps the group count uniform)
        }
    };
    if constexpr (ASYNC) {
        issue_tile();
        issue_tile();
    }
    auto store_nan_planes = [&](const uint32_t (&w)[16], uint32_t *dst) {  // dst = this block's word of plane 0
        if constexpr (IS_FLOAT) {
            uint32_t e = 0xffffffffu, mhi = 0;
#pragma unroll
            for (int b = EXP_LO; b <= 14; ++b) e &= w[b];
#pragma unroll
            for (int b = 0; b < EXP_LO; ++b) mhi |= w[b];
            dst[16 * nrbp] = e & mhi;
            dst[17 * nrbp] = e & ~mhi;
        }
    };
    unsigned long long iter = 0;
    for (unsigned long long strip = blockIdx.x; strip < n_strips; strip += gridDim.x, ++iter) {
        // 4-lane strips, launched as clusters of two CTAs that own the two halves of the same 32-byte sectors: the pair
        // starts every strip together, so that the second request for a sector finds it in L2 (same trip count in both
        // CTAs: the host sets pair_sync only for an even grid and an even number of strips)
        if (pair_sync) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        const unsigned long long grp = strip / strips_per_group;
        const unsigned long long col0 = (strip - grp * strips_per_group) * COLS_CW;
        long long in_off, out_off;
        lane_offsets(g, grp * ncols + col0, in_off, out_off);
        const T *sp = data + in_off;  // element (row 0, first lane of the strip)
        const int cols_here = (int)min((unsigned long long)COLS_CW, ncols - col0);

        // ---- stream the strip once: 1024 rows per step, 32 rows x 1 lane per thread ---------------------
        if constexpr (ASYNC) {
            uint32_t *pdst = planes + (size_t)bc * col_pitch;
#pragma unroll 1
            for (unsigned t = 0; t < tiles; ++t) {
                const unsigned long long gt = iter * tiles + t;
                cp_async_wait<COLS_STAGES - 1>();   // tile gt has landed (for this thread's copies) ...
                __syncthreads();                    // ... and for everybody's
                const uint32_t *st = stage_base + (size_t)(gt % COLS_STAGES) * COLS_STAGE_WORDS + brg * COLS_RB_WORDS + bc;
                const unsigned rb = t * 32u + (unsigned)brg;
                uint32_t x[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) x[i] = st[i * COLS_CW];
                __syncthreads();                    // the stage is free again
                issue_tile();                       // next tile of this strip, or the first tiles of the next one
                if (rb < nrb) {                     // rows past the end of the lane are masked out in the walk
                    uint32_t w[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) w[k] = __byte_perm(x[2 * k], x[2 * k + 1], 0x7632);
                    transpose16(w);
#pragma unroll
                    for (int b = 0; b < 16; ++b) pdst[b * nrbp + rb] = w[b];
                    store_nan_planes(w, pdst + rb);
                }
            }
        } else {
            const int c = min(bc, cols_here - 1);
            uint32_t *pdst = planes + (size_t)bc * col_pitch;
#pragma unroll 1
            for (unsigned row0 = 0; row0 < nrb * 32u; row0 += 1024u) {
                const unsigned rb = row0 / 32u + (unsigned)brg;
                if (rb >= nrb) break;  // blocks past the end of the lane are masked out in the walk
                Raw x[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const unsigned row = min(rb * 32u + (unsigned)i, n - 1u);  // clamped rows are masked later
                    x[i] = to_raw<T>(__ldg(sp + (long long)row * pitch + c));
                }
                uint32_t w[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    if constexpr (ES == 4) w[k] = __byte_perm((uint32_t)x[2 * k], (uint32_t)x[2 * k + 1], 0x7632);
                    else w[k] = __byte_perm((uint32_t)(x[2 * k] >> 32), (uint32_t)(x[2 * k + 1] >> 32), 0x7632);
                }
                transpose16(w);
#pragma unroll
                for (int b = 0; b < 16; ++b) pdst[b * nrbp + rb] = w[b];
                store_nan_planes(w, pdst + rb);
            }
        }
        __syncthreads();

        // ---- warp w selects the quantiles of lane w -----------------------------------------------------
        if (warp < cols_here) {
            const T *lp = sp + warp;  // element (row, this lane) = lp[row * pitch]
            uint32_t *pl = planes + (size_t)warp * col_pitch + lane;  // word k of plane b: pl[b * nrbp + 32 * k]
            unsigned *mylist = clist + warp * COLS_CAND;

            // valid rows and (floats) NaN positions, from the planes of round 0
            uint32_t init[COLS_KMAX];
            unsigned my_nan = 0;
#pragma unroll
            for (int k = 0; k < COLS_KMAX; ++k) {
                uint32_t v = 0;
                if ((unsigned)k < kw) {
                    const unsigned rb = (unsigned)lane + 32u * k;
                    const unsigned base = rb * 32u;
                    if (base + 32u <= n) v = 0xffffffffu;
                    else if (base < n)
                        for (int p = 0; p < 32; ++p)
                            if (base + (unsigned)slot_of_pos(p) < n) v |= 1u << p;
                    if (IS_FLOAT && v) {
                        uint32_t nanm = pl[16 * nrbp + 32 * k] & v;   // NaN by the top 16 bits
                        uint32_t amb = pl[17 * nrbp + 32 * k] & v;    // inf, or a NaN whose payload sits in the low bits
                        while (amb) {
                            int p = __ffs(amb) - 1;
                            amb &= amb - 1;
                            T xv = lp[(long long)(base + slot_of_pos(p)) * pitch];
                            if (xv != xv) nanm |= 1u << p;
                        }
                        v &= ~nanm;
                        my_nan += __popc(nanm);
                    }
                }
                init[k] = v;
            }
            unsigned n_eff = n;
            if (IS_FLOAT) {
                const unsigned nan_total = __reduce_add_sync(FULLMASK, my_nan);
                if (nan_total) {
                    if (!qa.skipnan && lane == 0) atomicOr(qa.status, STATUS_BIT_NAN);
                    n_eff = n - nan_total;
                }
            }
            const long long my_out = out_off + (long long)warp * out_col_stride;

            if (n_eff == 0) {
                for (int t = lane; t < qa.nq; t += 32) out[my_out + (long long)t * g.out_axis_stride] = canonical_nan<T>();
            } else {
                bool planes_dirty = false;
                // rebuild this lane's planes for `round` from global memory (L2-resident), warp-local
                auto rebuild = [&](int round) {
#pragma unroll 1
                    for (unsigned k = 0; k < kw; ++k) {
                        const unsigned rb = (unsigned)lane + 32u * k;
                        uint32_t w[16];
#pragma unroll
                        for (int kk = 0; kk < 16; ++kk) {
                            const unsigned r0 = min(rb * 32u + 2u * kk, n - 1u), r1 = min(rb * 32u + 2u * kk + 1u, n - 1u);
                            const Raw a = to_raw<T>(__ldg(lp + (long long)r0 * pitch));
                            const Raw b = to_raw<T>(__ldg(lp + (long long)r1 * pitch));
                            const int sh = KEYBITS - 16 * (round + 1);
                            w[kk] = (uint32_t)((a >> sh) & 0xffffu) | ((uint32_t)((b >> sh) & 0xffffu) << 16);
                        }
                        transpose16(w);
#pragma unroll
                        for (int b = 0; b < 16; ++b) pl[b * nrbp + 32 * k] = w[b];
                    }
                    __syncwarp();
                };

                auto select = [&](unsigned r, bool want_next, Raw &raw_r, Raw &raw_next) {
#pragma unroll 1
                    for (int pass = 0; pass < 2; ++pass) {
                        if (planes_dirty) {
                            rebuild(0);
                            planes_dirty = false;
                        }
                        uint32_t m[COLS_KMAX];
#pragma unroll
                        for (int k = 0; k < COLS_KMAX; ++k) m[k] = init[k];
                        unsigned total = n_eff, rem = r + (unsigned)pass;
                        bool neg = false;
                        Raw prefix = 0;
                        int bits_done = 0;
                        bool small = total <= COLS_CAND;
#pragma unroll 1
                        for (int round = 0; round < ROUNDS && !small; ++round) {
                            if (round > 0) {
                                rebuild(round);
                                planes_dirty = true;
                            }
                            uint32_t inv = (round == 0) ? (TOP_INVERTED ? 0xffffffffu : 0u)
                                                        : ((IS_FLOAT && neg) ? 0xffffffffu : 0u);
                            const uint32_t *pp = pl + 15 * nrbp;
                            uint32_t x[COLS_KMAX];   // the words of plane b; those of plane b - 1 are read while b is decided
#pragma unroll
                            for (int k = 0; k < COLS_KMAX; ++k) x[k] = ((unsigned)k < kw) ? pp[32 * k] : 0u;
#pragma unroll 1
                            for (int b = 15; b >= 0; --b, pp -= nrbp) {
                                uint32_t xn[COLS_KMAX];
                                const uint32_t *pn = b > 0 ? pp - nrbp : pp;
                                unsigned ones = 0;
#pragma unroll
                                for (int k = 0; k < COLS_KMAX; ++k) {
                                    xn[k] = ((unsigned)k < kw) ? pn[32 * k] : 0u;
                                    ones += __popc(m[k] & (x[k] ^ inv));
                                }
                                ones = __reduce_add_sync(FULLMASK, ones);
                                const unsigned small_cnt = total - ones;
                                const bool take_small = rem < small_cnt;
                                total = take_small ? small_cnt : ones;
                                rem = take_small ? rem : rem - small_cnt;
                                const uint32_t keep = take_small ? inv : ~inv;
                                prefix = (Raw)(prefix << 1) | (Raw)(keep & 1u);
#pragma unroll
                                for (int k = 0; k < COLS_KMAX; ++k) {
                                    m[k] &= ~(x[k] ^ keep);
                                    x[k] = xn[k];
                                }
                                if (round == 0 && b == 15) {
                                    neg = IS_FLOAT && (keep & 1u);
                                    inv = neg ? 0xffffffffu : 0u;
                                }
                                bits_done += 1;
                                if (total <= COLS_CAND) {
                                    small = true;
                                    break;
                                }
                            }
                        }
                        if (bits_done < KEYBITS) prefix <<= (KEYBITS - bits_done);
                        Raw found, found_next = 0;
                        bool have_next = false;
                        if (total > COLS_CAND) {
                            found = prefix;
                            if (rem + 1 < total) {
                                found_next = prefix;
                                have_next = true;
                            }
                        } else {
                            unsigned mycnt = 0;
#pragma unroll
                            for (int k = 0; k < COLS_KMAX; ++k) mycnt += __popc(m[k]);
                            unsigned incl = mycnt;
#pragma unroll
                            for (int d = 1; d < 32; d <<= 1) {
                                unsigned tt = __shfl_up_sync(FULLMASK, incl, d);
                                if (lane >= d) incl += tt;
                            }
                            unsigned base = incl - mycnt;
#pragma unroll
                            for (int k = 0; k < COLS_KMAX; ++k) {
                                uint32_t bits = m[k];
                                while (bits) {
                                    int p = __ffs(bits) - 1;
                                    bits &= bits - 1;
                                    mylist[base++] = ((unsigned)lane + 32u * k) * 32u + (unsigned)slot_of_pos(p);
                                }
                            }
                            __syncwarp();
                            // up to COLS_CAND / 32 candidates per thread: fetch them again (L2)
                            constexpr int CPT = COLS_CAND / 32;
                            Raw myraw[CPT];
                            Key mykey[CPT];
#pragma unroll
                            for (int c = 0; c < CPT; ++c) {
                                const unsigned idx = (unsigned)lane + 32u * c;
                                myraw[c] = 0;
                                mykey[c] = (Key) ~(Key)0;
                                if (idx < total) {
                                    T xv = lp[(long long)mylist[idx] * pitch];
                                    myraw[c] = to_raw<T>(xv);
                                    mykey[c] = Tr::to_key(xv);
                                }
                            }
                            // The candidates share the key bits the walk has consumed; the remaining bits are walked on
                            // the registers (one ballot per candidate slot and bit) until one key is left -- a handful of
                            // bits for distinct keys.  Equal keys are the same element value, any of them will do.
                            auto fetch_rank = [&](unsigned want) -> Raw {
                                bool alive[CPT];
#pragma unroll
                                for (int c = 0; c < CPT; ++c) alive[c] = (unsigned)lane + 32u * c < total;
                                unsigned tot2 = total, rem2 = want;
#pragma unroll 1
                                for (int b = KEYBITS - 1 - bits_done; b >= 0 && tot2 > 1u; --b) {
                                    bool one[CPT];
                                    unsigned ones = 0;
#pragma unroll
                                    for (int c = 0; c < CPT; ++c) {
                                        one[c] = ((mykey[c] >> b) & (Key)1) != 0;
                                        ones += __popc(__ballot_sync(FULLMASK, alive[c] && one[c]));
                                    }
                                    const unsigned zeros = tot2 - ones;
                                    const bool take0 = rem2 < zeros;
                                    tot2 = take0 ? zeros : ones;
                                    rem2 = take0 ? rem2 : rem2 - zeros;
#pragma unroll
                                    for (int c = 0; c < CPT; ++c) alive[c] = alive[c] && (one[c] != take0);
                                }
                                Raw v = 0;
                                bool mine = false;
#pragma unroll
                                for (int c = 0; c < CPT; ++c)
                                    if (alive[c] && !mine) {
                                        v = myraw[c];
                                        mine = true;
                                    }
                                const unsigned hit = __ballot_sync(FULLMASK, mine);
                                return __shfl_sync(FULLMASK, v, __ffs(hit) - 1);
                            };
                            found = fetch_rank(rem);
                            if (want_next && pass == 0 && rem + 1 < total) {  // (a second ballot walk: only when asked for)
                                found_next = fetch_rank(rem + 1);
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
                        select(r, want_next, a, b);
                        remember(r, a);
                        if (want_next) remember(r + 1, b);
                    }
                    if (lane == 0)
                        out[my_out + (long long)t * g.out_axis_stride] =
                            interpolate<T>(qa.interp, from_raw<T>(vl), from_raw<T>(vh), rm.frac, qa.status);
                }
            }
        }
        __syncthreads();  // the planes are rewritten by the next strip
    }
}

// ---------------------------------------------------------------------------------------------------------
// TMA-staged strips (4-byte types, whole 16-byte aligned strips of a 2-D / 3-D array): the default for BASELINE config 3.
//
// What holds cols_bitslice_kernel at a third of the HBM roofline is the number of bytes it keeps in flight: two 32 KB
// cp.async tiles, one of them always being consumed, none while the lanes are walked -- 148 SMs x 32 KB / ~1.2 us of
// loaded latency is 4 TB/s for 60 % of the time.  (cols16_kernel, 16 warps on the same planes, showed that more warps
// are not the cure: with its single 64 KB stage it is slower.)  Here a dedicated producer warp streams the strips
// through a ring of 8 KB slots -- tensor-map TMA boxes of 256 rows x 8 lanes, cp.async.bulk.tensor + mbarrier, no
// registers, no address arithmetic, no LSU -- and keeps every free slot of the ring (~70 KB) in flight at all times,
// also while the previous strip is being walked.  A slot is consumed by one PAIR of warps (64 threads: 8 lanes x 8
// blocks of 32 rows), the four pairs work on different slots and never meet at a CTA barrier during the build.
// A thread reads rows r, r + 4, r + 8 ... of one lane (conflict-free on the unpadded TMA box: a warp touches 32
// consecutive words), so plane word W of a lane holds the rows (W / 4) * 128 + (W % 4) + 4 * slot: the walk does not care
// which rows a word holds, only the valid masks and the survivors' addresses do.
// ---------------------------------------------------------------------------------------------------------
constexpr int CT_CW = 8;
constexpr int CT_SLOT_ROWS = 256;
constexpr unsigned CT_SLOT_BYTES = CT_SLOT_ROWS * CT_CW * 4;   // 8192
#ifndef NDS_CT_CONSUMERS
#define NDS_CT_CONSUMERS 256  // (384 = six pairs with a ring of six slots: 0.84 ms on C3 against 0.76 ms)
#endif
constexpr int CT_CONSUMERS = NDS_CT_CONSUMERS;                // pairs of warps build; the first 8 warps walk a lane each
constexpr int CT_THREADS = CT_CONSUMERS + 32;                 // + the producer warp
constexpr int CT_MAX_SLOTS = 16;
constexpr int CT_PAIRS = CT_CONSUMERS / 64;
// lanes of the producer warp that issue boxes.  Slot q of the CTA's sequence is issued by lane q % CT_PRODUCERS, lives
// in ring entry q % nslots and is read by pair q % CT_PAIRS; nslots is a multiple of both, so a ring entry is always
// filled by the same lane and read by the same pair -- which keeps everybody within one barrier phase of everybody else
// (a parity wait cannot tell two phases apart from none).
constexpr int CT_PRODUCERS = (CT_PAIRS % 4 == 0) ? 4 : 3;
constexpr int CT_SLOT_QUANTUM = (CT_PAIRS % CT_PRODUCERS == 0) ? CT_PAIRS : CT_PAIRS * CT_PRODUCERS;

__device__ __forceinline__ uint32_t ct_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ct_mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ct_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ct_mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ct_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ct_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ct_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ct_mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "CT_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra CT_DONE;\n"
        "bra CT_WAIT;\n"
        "CT_DONE:\n"
        "}\n" ::"r"(ct_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ bool ct_mbar_test(uint64_t *bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred P1;\nmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}"
                 : "=r"(ok) : "r"(ct_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// one box of the tensor map -> shared memory, completion on the mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void ct_tma_load_3d(void *dst_smem, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            ct_smem_u32(dst_smem)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(ct_smem_u32(bar))
        : "memory");
}

// arrive on the mbarrier once all cp.async of this thread have landed (the barrier's count includes these arrivals)
__device__ __forceinline__ void ct_cp_async_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(ct_smem_u32(bar)) : "memory");
}

// USE_TMA: the producer is one thread issuing tensor-map boxes (UTMALDG); otherwise the 32 threads of the producer warp
// issue the same slots as 16-byte cp.async copies (LDGSTS) that arrive on the same mbarriers.  A box of 256 rows x 32
// bytes is 256 separate 32-byte requests either way, and that request rate is what bounds the kernel: measured on
// B200, the TMA unit retires one such row every ~3.2 cycles per SM (2.8 TB/s over the chip), the load/store unit one
// per cycle.
template <typename T, bool USE_TMA>
__global__ void __launch_bounds__(CT_THREADS, 1)
cols_tma_kernel(const __grid_constant__ CUtensorMap tmap, const T *__restrict__ data, LaneGeom g, QuantArgs qa,
                T *__restrict__ out, unsigned long long n_strips, unsigned strips_per_group, int nslots) {
    static_assert(sizeof(T) == 4, "TMA boxes of 8 lanes x 4 bytes");
    using Tr = Traits<T>;
    using Key = typename Tr::Key;
    using Raw = typename RawBits<T>::U;
    constexpr int KEYBITS = 32, ROUNDS = 2;
    constexpr bool IS_FLOAT = Tr::is_float;
    constexpr bool TOP_INVERTED = IS_FLOAT || std::is_signed<T>::value;
    constexpr int EXP_LO = 7;
    // floats: two more planes, "NaN by its top 16 bits" and "exponent all ones, top mantissa bits clear" (inf, or a NaN whose
    // payload sits in the low bits), written by the build while the 16 plane words sit in registers.  (Deriving them at the
    // start of the walk -- 15 plane reads per word, to free 16 KB for a longer ring -- cost 8 % of the kernel's time, and
    // the ring is best at eight slots anyway.)
    constexpr int NPL = IS_FLOAT ? 18 : 16;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const unsigned n = (unsigned)g.axis_len;
    const long long pitch = g.axis_stride;
    const unsigned sps = (n + CT_SLOT_ROWS - 1) / CT_SLOT_ROWS;   // slots per strip
    const unsigned nw = sps * 8u;                                  // plane words per lane (one per 32 rows)
    const unsigned kw = (nw + 31u) / 32u;                          // plane words per thread in the walk
    const unsigned nrbp = kw * 32u;                                // plane pitch in words
    const unsigned col_pitch = (unsigned)NPL * nrbp + 4u;
    uint32_t *planes = reinterpret_cast<uint32_t *>(smem_raw);
    unsigned *clist = reinterpret_cast<unsigned *>(planes + CT_CW * col_pitch);     // [lane][COLS_CAND]
    uint64_t *full = reinterpret_cast<uint64_t *>(clist + CT_CW * COLS_CAND);       // [CT_MAX_SLOTS]
    uint64_t *empty = full + CT_MAX_SLOTS;                                           // [CT_MAX_SLOTS]
    // the ring: 128-byte aligned shared addresses (derived by an offset from smem_raw, see cols_bitslice_kernel)
    const uint32_t base_addr = ct_smem_u32(smem_raw);
    const uint32_t ring_addr = (ct_smem_u32(empty + CT_MAX_SLOTS) + 127u) & ~127u;
    unsigned char *ring = smem_raw + (ring_addr - base_addr);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int last_outer = g.n_outer - 1;
    const unsigned long long ncols = g.oshape[last_outer];
    const long long out_col_stride = g.out_ostride[last_outer];
    auto slot_of_pos = [](int p) -> int { return 2 * (p & 15) + (p >> 4); };
    // row held by bit p of plane word W
    auto row_of = [&](unsigned W, int p) -> unsigned { return (W >> 2) * 128u + (W & 3u) + 4u * (unsigned)slot_of_pos(p); };

    if (tid == 0) {
        for (int s = 0; s < nslots; ++s) {
            ct_mbar_init(full + s, USE_TMA ? 1 : 32);   // one expect_tx arrival / the 32 copying threads
            ct_mbar_init(empty + s, 2);                  // the two warps that read a slot
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const unsigned long long my_strips = blockIdx.x < n_strips ? (n_strips - blockIdx.x + gridDim.x - 1) / gridDim.x : 0ull;

    if (warp == CT_CONSUMERS / 32) {
        // ================= producer: every slot of every strip of this CTA, in order =================
        if constexpr (!USE_TMA) {
            unsigned long long q = 0;
            const unsigned rl0 = (unsigned)lane >> 1, half = (unsigned)lane & 1u;  // chunk 32 i + lane: row rl0 + 16 i
            for (unsigned long long it = 0; it < my_strips; ++it) {
                const unsigned long long strip = blockIdx.x + it * gridDim.x;
                const unsigned long long grp = strip / strips_per_group;
                const unsigned long long col0 = (strip - grp * strips_per_group) * CT_CW;
                long long io, oo;
                lane_offsets(g, grp * ncols + col0, io, oo);
                const T *src = data + io + (long long)rl0 * pitch + half * 4;
                const long long step = 16ll * pitch;
                for (unsigned j = 0; j < sps; ++j, ++q) {
                    const unsigned r = (unsigned)(q % (unsigned)nslots);
                    const unsigned long long turn = q / (unsigned)nslots;
                    if (turn > 0) ct_mbar_wait(empty + r, (unsigned)((turn - 1) & 1ull));
                    unsigned char *dst = ring + (size_t)r * CT_SLOT_BYTES + rl0 * 32u + half * 16u;
                    const unsigned row0 = j * CT_SLOT_ROWS + rl0;
                    if (row0 + CT_SLOT_ROWS - rl0 <= n) {  // a whole slot (uniform)
#pragma unroll
                        for (int i = 0; i < 16; ++i, src += step) cp_async16(dst + i * 512, src);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i, src += step)
                            if (row0 + 16u * i < n) cp_async16(dst + i * 512, src);
                    }
                    ct_cp_async_arrive(full + r);
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            return;
        }
        // Several lanes issue, each its own subsequence of the slots (q = lane, lane + CT_PRODUCERS, ...): one thread that
        // waits for a free slot, arms the barrier and issues the box needs ~280 ns per box, i.e. 29 GB/s per SM -- the
        // whole chip would stop at 4.3 TB/s (measured with the data thrown away).
        if (lane < CT_PRODUCERS) {
            const unsigned long long total = my_strips * sps;
            for (unsigned long long q = (unsigned long long)lane; q < total; q += CT_PRODUCERS) {
                const unsigned long long it = q / sps;
                const unsigned j = (unsigned)(q - it * sps);
                const unsigned long long strip = blockIdx.x + it * gridDim.x;
                const unsigned long long grp = (strip >> 32) ? strip / strips_per_group : (unsigned long long)((unsigned)strip / strips_per_group);
                const int col0 = (int)((strip - grp * strips_per_group) * CT_CW);
                const unsigned r = (unsigned)(q % (unsigned)nslots);
                const unsigned long long turn = q / (unsigned)nslots;
                if (turn > 0) ct_mbar_wait(empty + r, (unsigned)((turn - 1) & 1ull));  // the readers are done with it
                ct_mbar_expect_tx(full + r, CT_SLOT_BYTES);
                ct_tma_load_3d(ring + (size_t)r * CT_SLOT_BYTES, &tmap, col0, (int)(j * CT_SLOT_ROWS), (int)grp, full + r);
            }
        }
        return;
    }

    // ================= consumers =================
    const int pair = warp >> 1, ph = warp & 1;     // build role: which slots (j % 4 == pair), which 128 rows of the slot
    const int bc = lane & 7, bsub = lane >> 3;     // lane (column) and row phase: rows bsub, bsub + 4, ... of the 128
    // this warp's position in the CTA's slot sequence (slots q with q % CT_PAIRS == pair), ring entry and phase of slot cq
    const unsigned long long total_q = my_strips * sps;
    unsigned long long cq = (unsigned long long)pair;
    unsigned cr = (unsigned)pair % (unsigned)nslots, cpar = ((unsigned)pair / (unsigned)nslots) & 1u;
    uint32_t xn[32];          // the rows of slot cq, once `have_next`
    bool have_next = false;
    // pull slot cq into xn and hand the ring entry back; non-blocking calls return false if the slot has not landed
    auto fetch = [&](bool blocking) -> bool {
        if (blocking) ct_mbar_wait(full + cr, cpar);
        else if (!__all_sync(FULLMASK, ct_mbar_test(full + cr, cpar))) return false;
        const uint32_t *st = reinterpret_cast<const uint32_t *>(ring + (size_t)cr * CT_SLOT_BYTES) + (ph * 128 + bsub) * CT_CW + bc;
#pragma unroll
        for (int i = 0; i < 32; ++i) xn[i] = st[i * 4 * CT_CW];   // a warp reads 32 consecutive words per step
        __syncwarp();
        if (lane == 0) ct_mbar_arrive(empty + cr);   // this warp has read its half of the slot
        cr += CT_PAIRS;                              // (nslots is a multiple of CT_PAIRS)
        if (cr >= (unsigned)nslots) {
            cr -= (unsigned)nslots;
            cpar ^= 1u;
        }
        return true;
    };
    if (cq < total_q) have_next = fetch(true);
    for (unsigned long long it = 0; it < my_strips; ++it) {
        const unsigned long long strip = blockIdx.x + it * gridDim.x;
        const unsigned long long grp = (strip >> 32) ? strip / strips_per_group : (unsigned long long)((unsigned)strip / strips_per_group);
        const unsigned long long col0 = (strip - grp * strips_per_group) * CT_CW;
        long long in_off, out_off;
        lane_offsets(g, grp * ncols + col0, in_off, out_off);
        const T *sp = data + in_off;

        // ---- build: this pair's slots of the strip ---------------------------------------------------------
        // Software pipeline: the rows of the pair's NEXT slot are pulled into registers (and the slot handed back to the
        // producer) before the current one is transposed, so a landed slot never waits for a transposition -- the ring's
        // entries spend their time in flight, not parked.  The first slot of the next strip is pulled before the walk.
        {
            uint32_t *pdst = planes + (size_t)bc * col_pitch;
            const unsigned long long q_end = (it + 1) * sps;   // one past this strip's last slot
#pragma unroll 1
            while (cq < q_end) {
                uint32_t w[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) w[k] = __byte_perm(xn[2 * k], xn[2 * k + 1], 0x7632);
                const unsigned j = (unsigned)(cq - it * sps);
                cq += CT_PAIRS;
                have_next = false;
                if (cq < total_q) have_next = fetch(false);   // (may belong to the next strip)
                const unsigned W = (j * 2u + (unsigned)ph) * 4u + (unsigned)bsub;
                transpose16(w);
#pragma unroll
                for (int b = 0; b < 16; ++b) pdst[b * nrbp + W] = w[b];
                if constexpr (IS_FLOAT) {
                    uint32_t e = 0xffffffffu, mhi = 0;
#pragma unroll
                    for (int b = EXP_LO; b <= 14; ++b) e &= w[b];
#pragma unroll
                    for (int b = 0; b < EXP_LO; ++b) mhi |= w[b];
                    pdst[16 * nrbp + W] = e & mhi;
                    pdst[17 * nrbp + W] = e & ~mhi;
                }
                if (!have_next && cq < q_end) have_next = fetch(true);   // still this strip's: wait for it now
            }
        }
        named_bar_sync(1, CT_CONSUMERS);   // the strip's planes are complete

        // ---- warp w selects the quantiles of lane w ---------------------------------------------------------
        if (warp < CT_CW) {
            const T *lp = sp + warp;
            uint32_t *pl = planes + (size_t)warp * col_pitch + lane;  // word k of plane b: pl[b * nrbp + 32 * k]
            unsigned *mylist = clist + warp * COLS_CAND;

            uint32_t init[COLS_KMAX];
            unsigned my_nan = 0;
#pragma unroll
            for (int k = 0; k < COLS_KMAX; ++k) {
                uint32_t v = 0;
                if ((unsigned)k < kw) {
                    const unsigned W = (unsigned)lane + 32u * k;
                    const unsigned first = (W >> 2) * 128u;   // the word's rows lie in [first, first + 128)
                    if (first + 128u <= n) v = 0xffffffffu;
                    else if (first < n)
                        for (int p = 0; p < 32; ++p)
                            if (row_of(W, p) < n) v |= 1u << p;
                    if (IS_FLOAT && v) {
                        // exponent all ones: a NaN for sure if one of the top mantissa bits is set, otherwise (inf, or a NaN
                        // whose payload sits in the low mantissa bits) the element itself is inspected
                        uint32_t nanm = pl[16 * nrbp + 32 * k] & v;
                        uint32_t amb = pl[17 * nrbp + 32 * k] & v;
                        while (amb) {
                            int p = __ffs(amb) - 1;
                            amb &= amb - 1;
                            T xv = lp[(long long)row_of(W, p) * pitch];
                            if (xv != xv) nanm |= 1u << p;
                        }
                        v &= ~nanm;
                        my_nan += __popc(nanm);
                    }
                }
                init[k] = v;
            }
            unsigned n_eff = n;
            if (IS_FLOAT) {
                const unsigned nan_total = __reduce_add_sync(FULLMASK, my_nan);
                if (nan_total) {
                    if (!qa.skipnan && lane == 0) atomicOr(qa.status, STATUS_BIT_NAN);
                    n_eff = n - nan_total;
                }
            }
            const long long my_out = out_off + (long long)warp * out_col_stride;

            if (n_eff == 0) {
                for (int t = lane; t < qa.nq; t += 32) out[my_out + (long long)t * g.out_axis_stride] = canonical_nan<T>();
            } else {
                bool planes_dirty = false;
                auto rebuild = [&](int round) {  // this lane's planes for `round` from global memory (L2-resident), warp-local
#pragma unroll 1
                    for (unsigned k = 0; k < kw; ++k) {
                        const unsigned W = (unsigned)lane + 32u * k;
                        uint32_t w[16];
#pragma unroll
                        for (int kk = 0; kk < 16; ++kk) {
                            const unsigned base = (W >> 2) * 128u + (W & 3u);
                            const unsigned r0 = min(base + 8u * kk, n - 1u), r1 = min(base + 8u * kk + 4u, n - 1u);
                            const Raw a = to_raw<T>(__ldg(lp + (long long)r0 * pitch));
                            const Raw b = to_raw<T>(__ldg(lp + (long long)r1 * pitch));
                            const int sh = KEYBITS - 16 * (round + 1);
                            w[kk] = (uint32_t)((a >> sh) & 0xffffu) | ((uint32_t)((b >> sh) & 0xffffu) << 16);
                        }
                        transpose16(w);
#pragma unroll
                        for (int b = 0; b < 16; ++b) pl[b * nrbp + 32 * k] = w[b];
                    }
                    __syncwarp();
                };

                auto select = [&](unsigned r, bool want_next, Raw &raw_r, Raw &raw_next) {
#pragma unroll 1
                    for (int pass = 0; pass < 2; ++pass) {
                        if (planes_dirty) {
                            rebuild(0);
                            planes_dirty = false;
                        }
                        uint32_t m[COLS_KMAX];
#pragma unroll
                        for (int k = 0; k < COLS_KMAX; ++k) m[k] = init[k];
                        unsigned total = n_eff, rem = r + (unsigned)pass;
                        bool neg = false;
                        Raw prefix = 0;
                        int bits_done = 0;
                        bool small = total <= COLS_CAND;
#pragma unroll 1
                        for (int round = 0; round < ROUNDS && !small; ++round) {
                            if (round > 0) {
                                rebuild(round);
                                planes_dirty = true;
                            }
                            uint32_t inv = (round == 0) ? (TOP_INVERTED ? 0xffffffffu : 0u)
                                                        : ((IS_FLOAT && neg) ? 0xffffffffu : 0u);
                            const uint32_t *pp = pl + 15 * nrbp;
                            uint32_t x[COLS_KMAX];   // the words of plane b; those of plane b - 1 are read while b is decided
#pragma unroll
                            for (int k = 0; k < COLS_KMAX; ++k) x[k] = ((unsigned)k < kw) ? pp[32 * k] : 0u;
#pragma unroll 1
                            for (int b = 15; b >= 0; --b, pp -= nrbp) {
                                uint32_t xn[COLS_KMAX];
                                const uint32_t *pn = b > 0 ? pp - nrbp : pp;
                                unsigned ones = 0;
#pragma unroll
                                for (int k = 0; k < COLS_KMAX; ++k) {
                                    xn[k] = ((unsigned)k < kw) ? pn[32 * k] : 0u;
                                    ones += __popc(m[k] & (x[k] ^ inv));
                                }
                                ones = __reduce_add_sync(FULLMASK, ones);
                                const unsigned small_cnt = total - ones;
                                const bool take_small = rem < small_cnt;
                                total = take_small ? small_cnt : ones;
                                rem = take_small ? rem : rem - small_cnt;
                                const uint32_t keep = take_small ? inv : ~inv;
                                prefix = (Raw)(prefix << 1) | (Raw)(keep & 1u);
#pragma unroll
                                for (int k = 0; k < COLS_KMAX; ++k) {
                                    m[k] &= ~(x[k] ^ keep);
                                    x[k] = xn[k];
                                }
                                if (round == 0 && b == 15) {
                                    neg = IS_FLOAT && (keep & 1u);
                                    inv = neg ? 0xffffffffu : 0u;
                                }
                                bits_done += 1;
                                if (total <= COLS_CAND) {
                                    small = true;
                                    break;
                                }
                            }
                        }
                        if (bits_done < KEYBITS) prefix <<= (KEYBITS - bits_done);
                        Raw found, found_next = 0;
                        bool have_next = false;
                        if (total > COLS_CAND) {
                            found = prefix;
                            if (rem + 1 < total) {
                                found_next = prefix;
                                have_next = true;
                            }
                        } else {
                            unsigned mycnt = 0;
#pragma unroll
                            for (int k = 0; k < COLS_KMAX; ++k) mycnt += __popc(m[k]);
                            unsigned incl = mycnt;
#pragma unroll
                            for (int d = 1; d < 32; d <<= 1) {
                                unsigned tt = __shfl_up_sync(FULLMASK, incl, d);
                                if (lane >= d) incl += tt;
                            }
                            unsigned base = incl - mycnt;
#pragma unroll
                            for (int k = 0; k < COLS_KMAX; ++k) {
                                uint32_t bits = m[k];
                                while (bits) {
                                    int p = __ffs(bits) - 1;
                                    bits &= bits - 1;
                                    mylist[base++] = row_of((unsigned)lane + 32u * k, p);
                                }
                            }
                            __syncwarp();
                            constexpr int CPT = COLS_CAND / 32;
                            Raw myraw[CPT];
                            Key mykey[CPT];
#pragma unroll
                            for (int c = 0; c < CPT; ++c) {
                                const unsigned idx = (unsigned)lane + 32u * c;
                                myraw[c] = 0;
                                mykey[c] = (Key) ~(Key)0;
                                if (idx < total) {
                                    T xv = lp[(long long)mylist[idx] * pitch];
                                    myraw[c] = to_raw<T>(xv);
                                    mykey[c] = Tr::to_key(xv);
                                }
                            }
                            __syncwarp();
                            auto fetch_rank = [&](unsigned want) -> Raw {
                                bool alive[CPT];
#pragma unroll
                                for (int c = 0; c < CPT; ++c) alive[c] = (unsigned)lane + 32u * c < total;
                                unsigned tot2 = total, rem2 = want;
#pragma unroll 1
                                for (int b = KEYBITS - 1 - bits_done; b >= 0 && tot2 > 1u; --b) {
                                    bool one[CPT];
                                    unsigned ones = 0;
#pragma unroll
                                    for (int c = 0; c < CPT; ++c) {
                                        one[c] = ((mykey[c] >> b) & (Key)1) != 0;
                                        ones += __popc(__ballot_sync(FULLMASK, alive[c] && one[c]));
                                    }
                                    const unsigned zeros = tot2 - ones;
                                    const bool take0 = rem2 < zeros;
                                    tot2 = take0 ? zeros : ones;
                                    rem2 = take0 ? rem2 : rem2 - zeros;
#pragma unroll
                                    for (int c = 0; c < CPT; ++c) alive[c] = alive[c] && (one[c] != take0);
                                }
                                Raw v = 0;
                                bool mine = false;
#pragma unroll
                                for (int c = 0; c < CPT; ++c)
                                    if (alive[c] && !mine) {
                                        v = myraw[c];
                                        mine = true;
                                    }
                                const unsigned hit = __ballot_sync(FULLMASK, mine);
                                return __shfl_sync(FULLMASK, v, __ffs(hit) - 1);
                            };
                            found = fetch_rank(rem);
                            if (want_next && pass == 0 && rem + 1 < total) {  // (a second ballot walk: only when asked for)
                                found_next = fetch_rank(rem + 1);
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
                        select(r, want_next, a, b);
                        remember(r, a);
                        if (want_next) remember(r + 1, b);
                    }
                    if (lane == 0)
                        out[my_out + (long long)t * g.out_axis_stride] =
                            interpolate<T>(qa.interp, from_raw<T>(vl), from_raw<T>(vh), rm.frac, qa.status);
                }
            }
        }
        if (!have_next && cq < total_q) have_next = fetch(true);
        named_bar_sync(1, CT_CONSUMERS);  // the planes are rewritten by the next strip
    }
}

struct ColsPlan {
    bool tma = false;     // TMA-staged strips (cols_tma_kernel): 4-byte types, aligned 8-lane strips of a 2-D / 3-D array
    size_t smem_tma = 0;
    int nslots = 0;
    int cw = 8;           // lanes per strip of the staged path (tuning knob NDS_COLS_CW=4: 4-lane strips, two CTAs per SM, paired in clusters)
    bool ok = false;
    unsigned long long n_strips = 0;
    unsigned strips_per_group = 0;
    size_t smem = 0;
    bool async = false;   // cp.async staged build (4-byte types, whole 16-byte aligned strips)
};

ColsPlan plan_cols(const nds_ctx *ctx, const DeviceCall &c) {
    ColsPlan p;
    size_t es = 0;
    switch (c.dtype) {
    case NDS_F32: case NDS_I32: case NDS_U32: es = 4; break;
    case NDS_F64: case NDS_I64: case NDS_U64: es = 8; break;
    default: return p;
    }
    const LaneGeom &g = c.geom;
    if (c.valid || c.out_valid) return p;
    if (g.n_outer < 1) return p;
    const int last = g.n_outer - 1;
    if (g.in_ostride[last] != 1 || g.oshape[last] < 32) return p;  // neighbouring lanes must be contiguous
    if (g.axis_len < 64 || g.axis_len > 8192) return p;             // 8 lanes x 16 planes x n/32 words of smem
    if (g.axis_stride < (long long)g.oshape[last]) return p;        // rows must not overlap
    const unsigned nrb = (unsigned)((g.axis_len + 31) / 32), nrbp = (nrb + 31) / 32 * 32;
    const unsigned npl = (c.dtype == NDS_F32 || c.dtype == NDS_F64) ? 18 : 16;  // floats: + the two NaN planes
    auto smem_for = [&](int cw) { return (size_t)cw * (npl * nrbp + 4) * 4 + (size_t)cw * COLS_CAND * 4 + (size_t)cw * COLS_CAND * 8 + 64; };
    auto strips_for = [&](int cw) {
        const unsigned long long spg = (g.oshape[last] + cw - 1) / cw;
        p.strips_per_group = (unsigned)spg;
        p.n_strips = (g.nlanes / g.oshape[last]) * spg;
        return spg <= 0xffffffffull;
    };
    // the staged build (4-byte types) wants every row of every strip to be aligned 16-byte chunks
    bool aligned = es == 4 && g.oshape[last] % COLS_CW_ASYNC == 0 && ((uintptr_t)c.data % 16) == 0 && (g.axis_stride * 4) % 16 == 0;
    for (int d = 0; d < last; ++d)
        if (g.oshape[d] > 1 && (g.in_ostride[d] * 4) % 16 != 0) aligned = false;
    static const int cw_knob = getenv("NDS_COLS_CW") ? atoi(getenv("NDS_COLS_CW")) : COLS_CW_ASYNC;
    const int cwa = (cw_knob == 4 && g.oshape[last] % 8 == 0) ? 4 : COLS_CW_ASYNC;
    const size_t smem_async = ((smem_for(cwa) + 15) & ~(size_t)15) + (size_t)COLS_STAGES * cols_stage_words(cwa) * 4 + 16;
    static const bool tma_knob = !(getenv("NDS_COLS_TMA") && atoi(getenv("NDS_COLS_TMA")) == 0);
    if (tma_knob && aligned && cwa == 8 && g.n_outer <= 2 && g.axis_stride > 0 && (g.n_outer == 1 || g.in_ostride[0] > 0) &&
        g.axis_len <= 0x7fffffffull && g.oshape[last] <= 0x7fffffffull && (g.n_outer == 1 || g.oshape[0] <= 0x7fffffffull)) {
        const unsigned sps = (unsigned)((g.axis_len + CT_SLOT_ROWS - 1) / CT_SLOT_ROWS);
        const unsigned kw = (sps * 8 + 31) / 32, pitch_w = npl * kw * 32 + 4;
        if (kw <= COLS_KMAX) {
            const size_t fixed = (size_t)CT_CW * pitch_w * 4 + (size_t)CT_CW * COLS_CAND * 4 + 2 * CT_MAX_SLOTS * 8 + 128;
            if ((size_t)ctx->max_smem_optin > fixed + 4 * CT_SLOT_BYTES) {
                // eight slots (64 KB in flight per SM) are the optimum on B200: measured 0.726 ms on BASELINE config 3
                // against 0.80 ms with twelve -- with more rows in flight the SMs drift apart and the 32-byte row pieces of
                // neighbouring strips stop meeting in the same DRAM pages (the dry streaming kernel shows the same)
                p.nslots = (int)std::min<size_t>(8, ((size_t)ctx->max_smem_optin - fixed) / CT_SLOT_BYTES);
                if (const char *e = getenv("NDS_COLS_SLOTS")) p.nslots = std::max(CT_SLOT_QUANTUM, std::min(p.nslots, atoi(e)));  // tuning
                // a slot must always be filled by the same producer lane: lanes that alternated on a slot could run two
                // barrier phases apart, and a parity wait cannot tell those from zero
                p.nslots = p.nslots / CT_SLOT_QUANTUM * CT_SLOT_QUANTUM;
                p.smem_tma = fixed + (size_t)p.nslots * CT_SLOT_BYTES;
                p.tma = p.nslots > 0;
            }
        }
    }
    if (aligned && smem_async <= (size_t)ctx->max_smem_optin && !getenv("NDS_COLS_NO_ASYNC")) {
        if (!strips_for(cwa)) return p;
        p.cw = cwa;
        p.async = true;
        p.smem = smem_async;
        p.ok = true;
        return p;
    }
    if (!strips_for(COLS_CW_SYNC)) return p;
    p.smem = smem_for(COLS_CW_SYNC);
    if (p.smem > (size_t)ctx->max_smem_optin) return p;
    p.ok = true;
    return p;
}

template <typename T, bool ASYNC, int CW> cudaError_t launch_cols_typed2(nds_ctx *ctx, const DeviceCall &c, const ColsPlan &p) {
    auto kernel = cols_bitslice_kernel<T, ASYNC, CW>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return e;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, CW * 32, p.smem);
    if (occ < 1) occ = 1;
    unsigned grid = (unsigned)std::min<unsigned long long>(p.n_strips, (unsigned long long)ctx->sm_count * occ);
    if (ASYNC && CW == 4 && grid >= 2 && p.n_strips % 2 == 0) {  // pairs of CTAs on the two halves of the same sectors
        grid &= ~1u;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(CW * 32);
        cfg.dynamicSmemBytes = p.smem;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kernel, (const T *)c.data, c.geom, c.q, (T *)c.out, p.n_strips, p.strips_per_group, 1u);
        ctx->launches += 1;
        return e != cudaSuccess ? e : cudaGetLastError();
    }
    kernel<<<grid, CW * 32, p.smem, ctx->stream>>>((const T *)c.data, c.geom, c.q, (T *)c.out, p.n_strips, p.strips_per_group, 0u);
    ctx->launches += 1;
    return cudaGetLastError();
}

// Diagnostic (NDS_COLS_DRY=<box width in lanes>): stream the whole array through the slot ring with TMA boxes of the given
// width and 2048 / width rows (8 KB each) and throw the data away -- the ceiling of the access pattern, nothing else.
__global__ void __launch_bounds__(64, 1)
cols_dry_kernel(const __grid_constant__ CUtensorMap tmap, unsigned long long nboxes_x, unsigned long long nboxes_y, int box_w,
                int box_h, int nslots, unsigned *sink) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + CT_MAX_SLOTS;
    const uint32_t base_addr = ct_smem_u32(smem_raw);
    const uint32_t ring_addr = (ct_smem_u32(empty + CT_MAX_SLOTS) + 127u) & ~127u;
    unsigned char *ring = smem_raw + (ring_addr - base_addr);
    if (threadIdx.x == 0) {
        for (int s = 0; s < nslots; ++s) {
            ct_mbar_init(full + s, 1);
            ct_mbar_init(empty + s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // macro-strip x of this CTA, all its boxes top to bottom, then the next macro-strip
    const unsigned long long mine = blockIdx.x < nboxes_x ? (nboxes_x - blockIdx.x + gridDim.x - 1) / gridDim.x : 0ull;
    const unsigned long long total = mine * nboxes_y;
    if (threadIdx.x >= 32 && threadIdx.x < 32 + CT_PRODUCERS) {
        for (unsigned long long q = threadIdx.x - 32; q < total; q += CT_PRODUCERS) {
            const unsigned r = (unsigned)(q % (unsigned)nslots);
            const unsigned long long turn = q / (unsigned)nslots;
            if (turn > 0) ct_mbar_wait(empty + r, (unsigned)((turn - 1) & 1ull));
            const unsigned long long bx = blockIdx.x + (q / nboxes_y) * gridDim.x, by = q % nboxes_y;
            ct_mbar_expect_tx(full + r, CT_SLOT_BYTES);
            ct_tma_load_3d(ring + (size_t)r * CT_SLOT_BYTES, &tmap, (int)(bx * box_w), (int)(by * box_h), 0, full + r);
        }
    } else if (threadIdx.x == 0) {
        unsigned acc = 0;
        for (unsigned long long q = 0; q < total; ++q) {
            const unsigned r = (unsigned)(q % (unsigned)nslots);
            ct_mbar_wait(full + r, (unsigned)((q / (unsigned)nslots) & 1ull));
            acc += *reinterpret_cast<volatile unsigned *>(ring + (size_t)r * CT_SLOT_BYTES);
            ct_mbar_arrive(empty + r);
        }
        if (acc == 0x12345678u) *sink = acc;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {  // the driver entry point, without linking libcuda
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}

template <typename T> cudaError_t launch_cols_tma(nds_ctx *ctx, const DeviceCall &c, const ColsPlan &p) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return cudaErrorNotSupported;
    const LaneGeom &g = c.geom;
    const int last = g.n_outer - 1;
    // the array as a rank-3 tensor {lanes (contiguous), rows along the quantile axis, groups}; a box is 8 lanes x 256 rows
    cuuint64_t dims[3] = {(cuuint64_t)g.oshape[last], (cuuint64_t)g.axis_len, (cuuint64_t)(g.n_outer == 2 ? g.oshape[0] : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)g.axis_stride * sizeof(T),
                             (cuuint64_t)(g.n_outer == 2 ? g.in_ostride[0] : g.axis_stride * (long long)g.axis_len) * sizeof(T)};
    cuuint32_t box[3] = {CT_CW, CT_SLOT_ROWS, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap tmap;
    if (const char *dry = getenv("NDS_COLS_DRY")) {  // diagnostic: the access pattern alone (the result is NOT computed)
        const int bw = atoi(dry);
        if (bw >= 4 && bw <= 256 && 2048 % bw == 0 && g.oshape[last] % bw == 0) {
            box[0] = (cuuint32_t)bw;
            box[1] = (cuuint32_t)(2048 / bw);
            int promo = getenv("NDS_COLS_PROMO") ? atoi(getenv("NDS_COLS_PROMO")) : 2;
            if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void *>(c.data), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return cudaErrorNotSupported;
            int ns = getenv("NDS_COLS_SLOTS") ? atoi(getenv("NDS_COLS_SLOTS")) : 10;
            ns = std::max(CT_PRODUCERS, std::min(ns, CT_MAX_SLOTS)) / CT_PRODUCERS * CT_PRODUCERS;
            const size_t smem = 2 * CT_MAX_SLOTS * 8 + 128 + (size_t)ns * CT_SLOT_BYTES;
            cudaError_t e = cudaFuncSetAttribute(cols_dry_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            const unsigned long long nbx = g.oshape[last] / bw, nby = (g.axis_len + box[1] - 1) / box[1];
            const unsigned grid = (unsigned)std::min<unsigned long long>(nbx, (unsigned long long)ctx->sm_count);
            cols_dry_kernel<<<grid, 64, smem, ctx->stream>>>(tmap, nbx, nby, bw, (int)box[1], ns, (unsigned *)c.out);
            ctx->launches += 1;
            return cudaGetLastError();
        }
    }
    const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void *>(c.data), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorNotSupported;
    static const bool ldgsts_producer = getenv("NDS_COLS_TMA") && atoi(getenv("NDS_COLS_TMA")) == 3;  // 3: cp.async producer (slower)
    const unsigned grid = (unsigned)std::min<unsigned long long>(p.n_strips, (unsigned long long)ctx->sm_count);
    auto launch = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_tma);
        if (e != cudaSuccess) return e;
        kernel<<<grid, CT_THREADS, p.smem_tma, ctx->stream>>>(tmap, (const T *)c.data, g, c.q, (T *)c.out, p.n_strips,
                                                              p.strips_per_group, p.nslots);
        ctx->launches += 1;
        return cudaGetLastError();
    };
    if (ldgsts_producer) return launch(cols_tma_kernel<T, false>);
    return launch(cols_tma_kernel<T, true>);
}

template <typename T> cudaError_t launch_cols_typed(nds_ctx *ctx, const DeviceCall &c, const ColsPlan &p) {
    if constexpr (sizeof(T) == 4) {
        if (p.tma && p.ok && p.async && p.cw == 8) {
            cudaError_t e = launch_cols_tma<T>(ctx, c, p);
            if (e != cudaErrorNotSupported) return e;   // (no tensor map could be made: the cp.async kernel takes the call)
        }
        if (p.async && p.cw == 4) return launch_cols_typed2<T, true, 4>(ctx, c, p);
        if (p.async) return launch_cols_typed2<T, true, COLS_CW_ASYNC>(ctx, c, p);
    }
    return launch_cols_typed2<T, false, COLS_CW_SYNC>(ctx, c, p);
}

}  // namespace

bool cols_bitslice_supported(const nds_ctx *ctx, const DeviceCall &c) { return plan_cols(ctx, c).ok; }

cudaError_t launch_cols_bitslice(nds_ctx *ctx, const DeviceCall &c) {
    ColsPlan p = plan_cols(ctx, c);
    if (!p.ok) return cudaErrorNotSupported;
    ctx->last_path = "cols_bitslice";
    switch (c.dtype) {
    case NDS_F32: return launch_cols_typed<float>(ctx, c, p);
    case NDS_F64: return launch_cols_typed<double>(ctx, c, p);
    case NDS_I32: return launch_cols_typed<int32_t>(ctx, c, p);
    case NDS_U32: return launch_cols_typed<uint32_t>(ctx, c, p);
    case NDS_I64: return launch_cols_typed<int64_t>(ctx, c, p);
    case NDS_U64: return launch_cols_typed<uint64_t>(ctx, c, p);
    default: return cudaErrorNotSupported;
    }
}

}  // namespace nds
