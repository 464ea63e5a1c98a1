/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU oracle for the bulk quantile path of rust-ndarray/ndarray-stats v0.7.0.  It is a plain-C
 * restatement of the reference algorithm; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / `--impl reference` legs may load it.  The product library (libndstats_b200.so)
 * never links, loads or calls anything in this directory.
 *
 * PARITY PINNING: PARTIAL.  The reference is Rust and there is no Rust toolchain in the build
 * container, so the crate itself cannot be run or linked here.  The oracle is pinned
 * against every known-answer test the reference holds for this path (tests/quantile.rs:172-278,
 * tests/sort.rs:5-44, tests/maybe_nan.rs:5-31, doc-test src/quantile/mod.rs:237-249; transcribed in
 * tests/golden/reference_kats.json) and cross-checked by an independent sort-and-index
 * implementation (select_impl=1).  Behaviour those tests do not pin (any f32 case, N64 non-skipnan,
 * Nearest at fraction exactly 0.5, +-0/+-inf, Linear to the ulp) rests on this restatement alone.
 *
 * Results-defining arithmetic that lives in un-vendored crates (Cargo.lock pins), restated here:
 *   noisy_float 0.2.0  N64*f64, floor/ceil/fract, Ord by value (-0.0 == +0.0)
 *   num-traits 0.2.19  ToPrimitive::to_usize, FromPrimitive::from_f64 (truncate toward zero,
 *                      None outside the exclusive range), from_u8(2)
 *   ndarray 0.17.1     lane enumeration, C-order result layout
 *
 * Restated (paths relative to /root/reference):
 *   src/quantile/mod.rs:417-493   quantiles_axis_mut (validation order, rank set, lane loop)
 *   src/quantile/mod.rs:495-508   quantile_axis_mut   (nq == 1; the caller drops the axis)
 *   src/quantile/mod.rs:510-544   quantile_axis_skipnan_mut
 *   src/quantile/mod.rs:612-633   Quantile1dExt (ndim == 1, axis == 0)
 *   src/quantile/interpolate.rs:5-25,64-144   index math and the five strategies
 *   src/sort.rs:133-168,184-283   partition_mut and the multi-rank quickselect
 *   src/maybe_nan/mod.rs:46-71,109-150   NaN compaction, canonical NaN out
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off -fno-fast-math; OpenMP only parallelises the
 * independent lane loop, which the reference runs sequentially, src/quantile/mod.rs:470-472).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum {
    ORACLE_OK = 0,
    ORACLE_EMPTY_INPUT = 1,      /* QuantileError::EmptyInput, src/errors.rs:121-122 */
    ORACLE_INVALID_QUANTILE = 2, /* QuantileError::InvalidQuantile(q), src/errors.rs:123-124 */
    ORACLE_NAN_IN_INPUT = 3,     /* precondition of N32/N64 violated */
    ORACLE_UNSUPPORTED = 4,
    ORACLE_PANIC = 8,            /* the reference would panic (from_f64(..).unwrap()) */
    ORACLE_BAD_ARGUMENT = 7,
};
enum { ORACLE_LOWER = 0, ORACLE_HIGHER = 1, ORACLE_NEAREST = 2, ORACLE_MIDPOINT = 3, ORACLE_LINEAR = 4 };
enum {
    ORACLE_I8 = 0, ORACLE_I16, ORACLE_I32, ORACLE_I64,
    ORACLE_U8, ORACLE_U16, ORACLE_U32, ORACLE_U64,
    ORACLE_F32, ORACLE_F64
};

static inline uint64_t oracle_rng_next(uint64_t *s) {
    uint64_t x = *s;
    x ^= x >> 12;
    x ^= x << 25;
    x ^= x >> 27;
    *s = x;
    return x * 0x2545F4914F6CDD1Dull;
}

/* float_quantile_index — src/quantile/interpolate.rs:5-7: q * ((len - 1) as f64), one multiply */
static inline double oracle_float_index(double q, size_t len) {
    volatile double x = q * (double)(len - 1);
    return x;
}
/* float_quantile_index_fraction — interpolate.rs:13-15: f64::fract = x - x.trunc() */
static inline double oracle_index_fraction(double q, size_t len) {
    double x = oracle_float_index(q, len);
    return x - trunc(x);
}
/* lower_index / higher_index — interpolate.rs:18-25 */
static inline size_t oracle_lower_index(double q, size_t len) { return (size_t)floor(oracle_float_index(q, len)); }
static inline size_t oracle_higher_index(double q, size_t len) { return (size_t)ceil(oracle_float_index(q, len)); }

/* Interpolate::needs_lower / needs_higher — interpolate.rs:65-70,78-83,91-96,111-116,130-135 */
static inline int oracle_needs_lower(int interp, double q, size_t len) {
    switch (interp) {
    case ORACLE_HIGHER: return 0;
    case ORACLE_LOWER: return 1;
    case ORACLE_NEAREST: return oracle_index_fraction(q, len) < 0.5;
    default: return 1;
    }
}
static inline int oracle_needs_higher(int interp, double q, size_t len) {
    switch (interp) {
    case ORACLE_HIGHER: return 1;
    case ORACLE_LOWER: return 0;
    case ORACLE_NEAREST: return !oracle_needs_lower(interp, q, len);
    default: return 1;
    }
}

static int cmp_size_t(const void *a, const void *b) {
    size_t x = *(const size_t *)a, y = *(const size_t *)b;
    return (x > y) - (x < y);
}

/* searched_indexes — src/quantile/mod.rs:457-467: push, sort, dedup */
static size_t oracle_searched_indexes(int interp, const double *qs, size_t nq, size_t len, size_t *out) {
    size_t n = 0;
    for (size_t t = 0; t < nq; ++t) {
        if (oracle_needs_lower(interp, qs[t], len)) out[n++] = oracle_lower_index(qs[t], len);
        if (oracle_needs_higher(interp, qs[t], len)) out[n++] = oracle_higher_index(qs[t], len);
    }
    qsort(out, n, sizeof(size_t), cmp_size_t);
    size_t m = 0;
    for (size_t i = 0; i < n; ++i)
        if (m == 0 || out[m - 1] != out[i]) out[m++] = out[i];
    return m;
}

/* num-traits 0.2.19 `impl ToPrimitive for f64` (float -> int): truncate toward zero; Some only in
 * the exclusive range (MIN-1, MAX+1) when f64 is wider than the target, [MIN, MAX+1) otherwise. */
#define DEF_FROM_F64_SIGNED_SMALL(SUFX, TY, MINV, MAXV)                   \
    static int from_f64_##SUFX(double x, TY *out) {                       \
        if (x > (double)(MINV) - 1.0 && x < (double)(MAXV) + 1.0) {       \
            *out = (TY)x;                                                 \
            return 1;                                                     \
        }                                                                 \
        return 0;                                                         \
    }
#define DEF_FROM_F64_UNSIGNED_SMALL(SUFX, TY, MAXV)      \
    static int from_f64_##SUFX(double x, TY *out) {      \
        if (x > -1.0 && x < (double)(MAXV) + 1.0) {      \
            *out = (TY)x;                                \
            return 1;                                    \
        }                                                \
        return 0;                                        \
    }
DEF_FROM_F64_SIGNED_SMALL(i8, int8_t, INT8_MIN, INT8_MAX)
DEF_FROM_F64_SIGNED_SMALL(i16, int16_t, INT16_MIN, INT16_MAX)
DEF_FROM_F64_SIGNED_SMALL(i32, int32_t, INT32_MIN, INT32_MAX)
DEF_FROM_F64_UNSIGNED_SMALL(u8, uint8_t, UINT8_MAX)
DEF_FROM_F64_UNSIGNED_SMALL(u16, uint16_t, UINT16_MAX)
DEF_FROM_F64_UNSIGNED_SMALL(u32, uint32_t, UINT32_MAX)
static int from_f64_i64(double x, int64_t *out) {
    if (x >= -9223372036854775808.0 && x < 9223372036854775808.0) {
        *out = (int64_t)x;
        return 1;
    }
    return 0;
}
static int from_f64_u64(double x, uint64_t *out) {
    if (x > -1.0 && x < 18446744073709551616.0) {
        *out = (uint64_t)x;
        return 1;
    }
    return 0;
}

/* ::std::f32::NAN / ::std::f64::NAN — src/maybe_nan/mod.rs:126-131 */
static inline float canonical_nan_f32(void) {
    uint32_t b = 0x7FC00000u;
    float f;
    memcpy(&f, &b, 4);
    return f;
}
static inline double canonical_nan_f64(void) {
    uint64_t b = 0x7FF8000000000000ull;
    double f;
    memcpy(&f, &b, 8);
    return f;
}

#define T int8_t
#define UT uint8_t
#define SUF i8
#define IS_FLOAT 0
#define IS_SIGNED 1
#include "oracle_typed.inc"
#define T int16_t
#define UT uint16_t
#define SUF i16
#define IS_FLOAT 0
#define IS_SIGNED 1
#include "oracle_typed.inc"
#define T int32_t
#define UT uint32_t
#define SUF i32
#define IS_FLOAT 0
#define IS_SIGNED 1
#include "oracle_typed.inc"
#define T int64_t
#define UT uint64_t
#define SUF i64
#define IS_FLOAT 0
#define IS_SIGNED 1
#include "oracle_typed.inc"
#define T uint8_t
#define UT uint8_t
#define SUF u8
#define IS_FLOAT 0
#define IS_SIGNED 0
#include "oracle_typed.inc"
#define T uint16_t
#define UT uint16_t
#define SUF u16
#define IS_FLOAT 0
#define IS_SIGNED 0
#include "oracle_typed.inc"
#define T uint32_t
#define UT uint32_t
#define SUF u32
#define IS_FLOAT 0
#define IS_SIGNED 0
#include "oracle_typed.inc"
#define T uint64_t
#define UT uint64_t
#define SUF u64
#define IS_FLOAT 0
#define IS_SIGNED 0
#include "oracle_typed.inc"
#define T float
#define UT uint32_t
#define SUF f32
#define IS_FLOAT 1
#define IS_SIGNED 1
#include "oracle_typed.inc"
#define T double
#define UT uint64_t
#define SUF f64
#define IS_FLOAT 1
#define IS_SIGNED 1
#include "oracle_typed.inc"

/*
 * Entry point.  Mirrors the product C ABI (include/ndstats_b200.h) so tests read the same on both
 * sides.  `data` is mutated like the reference mutates its lanes (pass a copy); strides are in
 * elements and may be negative or zero-extent.  select_impl: 0 faithful quickselect, 1 sort-and-index.
 * nthreads: lanes are split over this many OpenMP threads (1 = the reference's sequential loop).
 */
int oracle_quantiles_axis(int dtype, void *data, int ndim, const size_t *shape, const ptrdiff_t *strides,
                          int axis, const double *qs, size_t nq, int interp, int skipnan, void *out,
                          size_t *bad_q, int select_impl, int nthreads) {
    if (ndim < 1 || axis < 0 || axis >= ndim) return ORACLE_BAD_ARGUMENT;
    if (interp < ORACLE_LOWER || interp > ORACLE_LINEAR) return ORACLE_BAD_ARGUMENT;
#define DISPATCH(CODE, TY, SUFX)                                                                          \
    case CODE:                                                                                            \
        return quantiles_axis_##SUFX((TY *)data, ndim, shape, strides, axis, qs, nq, interp, skipnan,     \
                                     (TY *)out, bad_q, select_impl, nthreads);
    switch (dtype) {
        DISPATCH(ORACLE_I8, int8_t, i8)
        DISPATCH(ORACLE_I16, int16_t, i16)
        DISPATCH(ORACLE_I32, int32_t, i32)
        DISPATCH(ORACLE_I64, int64_t, i64)
        DISPATCH(ORACLE_U8, uint8_t, u8)
        DISPATCH(ORACLE_U16, uint16_t, u16)
        DISPATCH(ORACLE_U32, uint32_t, u32)
        DISPATCH(ORACLE_U64, uint64_t, u64)
        DISPATCH(ORACLE_F32, float, f32)
        DISPATCH(ORACLE_F64, double, f64)
    }
#undef DISPATCH
    return ORACLE_UNSUPPORTED;
}

/* Sort1dExt::partition_mut on a strided i64 view — used by tests/sort KATs (tests/sort.rs:5-33). */
size_t oracle_partition_mut_i64(int64_t *data, size_t len, ptrdiff_t stride, size_t pivot_index) {
    view_i64 v = {data, len, stride};
    return partition_mut_i64(v, pivot_index);
}

/* Sort1dExt::get_many_from_sorted_mut — src/sort.rs:121-131 (sort + dedup, then unchecked).
 * Writes the deduplicated ascending indexes and their values; returns how many. */
size_t oracle_get_many_from_sorted_i64(int64_t *data, size_t len, ptrdiff_t stride, const size_t *indexes,
                                       size_t k, size_t *out_idx, int64_t *out_val) {
    size_t *tmp = (size_t *)malloc((k + 1) * sizeof(size_t));
    memcpy(tmp, indexes, k * sizeof(size_t));
    qsort(tmp, k, sizeof(size_t), cmp_size_t);
    size_t m = 0;
    for (size_t i = 0; i < k; ++i)
        if (m == 0 || tmp[m - 1] != tmp[i]) tmp[m++] = tmp[i];
    memcpy(out_idx, tmp, m * sizeof(size_t));
    view_i64 v = {data, len, stride};
    uint64_t rng = 0x1234567887654321ull;
    size_t *scratch = (size_t *)malloc((m + 1) * sizeof(size_t));
    get_many_i64(v, out_idx, m, out_val, 0, &rng, scratch);
    free(scratch);
    free(tmp);
    return m;
}

/* remove_nan_mut on a strided f64 view — tests/maybe_nan.rs:5-31.  Returns the prefix length. */
size_t oracle_remove_nan_mut_f64(double *data, size_t len, ptrdiff_t stride) {
    view_f64 v = {data, len, stride};
    return remove_nan_mut_f64(v);
}

/* exposed so host-logic tests can compare the product's rank computation with the restatement */
size_t oracle_lower_index_pub(double q, size_t len) { return oracle_lower_index(q, len); }
size_t oracle_higher_index_pub(double q, size_t len) { return oracle_higher_index(q, len); }
double oracle_index_fraction_pub(double q, size_t len) { return oracle_index_fraction(q, len); }

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
