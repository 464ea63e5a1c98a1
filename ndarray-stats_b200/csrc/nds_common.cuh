// Shared device-side pieces of the quantile path: order-preserving key transforms, the reference's
// index math and the five interpolation formulas, lane geometry.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ndstats_b200.h"

namespace nds {

// deferred, data-dependent status bits accumulated in the ctx's device status word
enum : int { STATUS_BIT_NAN = 1, STATUS_BIT_CAST_OVERFLOW = 2, STATUS_BIT_FALLBACK = 4, STATUS_BIT_XCHG = 8 };

// ---------------------------------------------------------------------------------------------
// Lane geometry: every 1-D lane along `axis` of an ndarray-style view (what the reference's
// `data.lanes_mut(axis)` enumerates, src/quantile/mod.rs:470-472), lanes numbered in C order over the
// remaining dims so that the C-order result (mod.rs:469) is addressed by the same decode.
// ---------------------------------------------------------------------------------------------
struct LaneGeom {
    int n_outer;                                  // number of non-axis dims
    unsigned long long oshape[NDS_MAX_DIMS];      // their extents (C order)
    long long in_ostride[NDS_MAX_DIMS];           // input strides of those dims, in elements
    long long out_ostride[NDS_MAX_DIMS];          // strides of the C-order result for those dims
    unsigned long long axis_len;                  // lane length
    long long axis_stride;                        // input stride along the lane, in elements
    long long out_axis_stride;                    // result stride between consecutive quantiles
    unsigned long long nlanes;
};

__host__ __device__ inline void lane_offsets(const LaneGeom &g, unsigned long long lane, long long &in_off,
                                             long long &out_off) {
    in_off = 0;
    out_off = 0;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (int d = g.n_outer - 1; d >= 1; --d) {  // 64-bit division only for arrays with >= 3 dims
        const unsigned long long q = lane / g.oshape[d];
        const unsigned long long c = lane - q * g.oshape[d];
        lane = q;
        in_off += (long long)c * g.in_ostride[d];
        out_off += (long long)c * g.out_ostride[d];
    }
    if (g.n_outer >= 1) {  // lane < nlanes, so what is left is the coordinate of the outermost dim
        in_off += (long long)lane * g.in_ostride[0];
        out_off += (long long)lane * g.out_ostride[0];
    }
}

constexpr int NDS_Q_INLINE = 8;
struct QuantArgs {
    const double *qs;  // device pointer, nq entries (already validated on the host); unused when `inline_q`
    double qv[NDS_Q_INLINE];  // up to 8 quantiles travel inside the kernel arguments: no host-to-device copy per call
    int inline_q;
    const unsigned long long *ranks;  // Sort1dExt mode: nq explicit ranks instead of qs (interp = LOWER)
    int nq;
    int interp;        // nds_interp
    int skipnan;       // floats: drop NaNs per lane; with a validity mask: drop None
    int *status;       // device status word (STATUS_BIT_*)
};

// ---------------------------------------------------------------------------------------------
// Order-preserving key transform: element -> unsigned key of the same width whose unsigned order is
// the element order.  Floats: flip all bits of negatives, the sign bit of non-negatives (-0.0 sorts
// just below +0.0, a refinement of N32/N64's value order, SURVEY.md note Z).  NaN is handled by the
// callers: it never reaches a histogram as a real key.
// ---------------------------------------------------------------------------------------------
template <typename T> struct Traits;

#define NDS_INT_TRAITS(TY, KEY, SIGNBIT)                                                          \
    template <> struct Traits<TY> {                                                               \
        using Key = KEY;                                                                          \
        static constexpr bool is_float = false;                                                   \
        static constexpr int key_bits = 8 * (int)sizeof(KEY);                                     \
        __host__ __device__ static inline Key to_key(TY v) { return (Key)((Key)v ^ (Key)(SIGNBIT)); } \
        __host__ __device__ static inline TY from_key(Key k) { return (TY)(Key)(k ^ (Key)(SIGNBIT)); } \
        __host__ __device__ static inline bool is_nan(TY) { return false; }                       \
    };
NDS_INT_TRAITS(int8_t, uint8_t, 0x80u)
NDS_INT_TRAITS(int16_t, uint16_t, 0x8000u)
NDS_INT_TRAITS(int32_t, uint32_t, 0x80000000u)
NDS_INT_TRAITS(int64_t, uint64_t, 0x8000000000000000ull)
NDS_INT_TRAITS(uint8_t, uint8_t, 0u)
NDS_INT_TRAITS(uint16_t, uint16_t, 0u)
NDS_INT_TRAITS(uint32_t, uint32_t, 0u)
NDS_INT_TRAITS(uint64_t, uint64_t, 0ull)
#undef NDS_INT_TRAITS

template <> struct Traits<float> {
    using Key = uint32_t;
    static constexpr bool is_float = true;
    static constexpr int key_bits = 32;
    __device__ static inline Key to_key(float x) {
        uint32_t b = __float_as_uint(x);
        return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
    }
    __device__ static inline float from_key(Key k) {
        return __uint_as_float(k ^ ((uint32_t)((int32_t)(~k) >> 31) | 0x80000000u));
    }
    __device__ static inline bool is_nan(float x) { return x != x; }
};
template <> struct Traits<double> {
    using Key = uint64_t;
    static constexpr bool is_float = true;
    static constexpr int key_bits = 64;
    __device__ static inline Key to_key(double x) {
        uint64_t b = (uint64_t)__double_as_longlong(x);
        return b ^ ((uint64_t)((int64_t)b >> 63) | 0x8000000000000000ull);
    }
    __device__ static inline double from_key(Key k) {
        return __longlong_as_double((long long)(k ^ ((uint64_t)((int64_t)(~k) >> 63) | 0x8000000000000000ull)));
    }
    __device__ static inline bool is_nan(double x) { return x != x; }
};

// ::std::f32::NAN / ::std::f64::NAN, src/maybe_nan/mod.rs:126-131
template <typename T> __device__ inline T canonical_nan() { return T(0); }
template <> __device__ inline float canonical_nan<float>() { return __uint_as_float(0x7FC00000u); }
template <> __device__ inline double canonical_nan<double>() {
    return __longlong_as_double(0x7FF8000000000000ll);
}

// ---------------------------------------------------------------------------------------------
// Index math — src/quantile/interpolate.rs:5-25.  One IEEE multiply, then floor / ceil / fract.
// ---------------------------------------------------------------------------------------------
struct RankMath {
    unsigned long long lower, higher;
    double frac;
};
__host__ __device__ inline RankMath rank_math(double q, unsigned long long len) {
    RankMath r;
#ifdef __CUDA_ARCH__
    double x = __dmul_rn(q, __ull2double_rn(len - 1));
#else
    volatile double xv = q * (double)(len - 1);
    double x = xv;
#endif
    r.lower = (unsigned long long)floor(x);
    r.higher = (unsigned long long)ceil(x);
    r.frac = x - trunc(x);
    return r;
}
// rank slot t of a lane with n_eff valid elements: from q (quantile API) or verbatim (Sort1dExt API)
__device__ inline RankMath slot_rank_math(const QuantArgs &qa, int t, unsigned long long n_eff) {
    if (qa.ranks) {
        RankMath r;
        r.lower = r.higher = qa.ranks[t];
        r.frac = 0.0;
        return r;
    }
    return rank_math(qa.inline_q ? qa.qv[t] : qa.qs[t], n_eff);
}
// Interpolate::needs_lower / needs_higher — interpolate.rs:65-70,78-83,91-96,111-116,130-135
__host__ __device__ inline bool needs_lower(int interp, double frac) {
    return interp == NDS_LOWER || interp == NDS_MIDPOINT || interp == NDS_LINEAR ||
           (interp == NDS_NEAREST && frac < 0.5);
}
__host__ __device__ inline bool needs_higher(int interp, double frac) {
    return interp == NDS_HIGHER || interp == NDS_MIDPOINT || interp == NDS_LINEAR ||
           (interp == NDS_NEAREST && !(frac < 0.5));
}

// ---------------------------------------------------------------------------------------------
// Interpolate::interpolate — src/quantile/interpolate.rs:64-144, in the element type, one rounding
// per operation (explicit _rn intrinsics: ptxas must not contract into FMA).
// ---------------------------------------------------------------------------------------------
template <typename T> struct WrapU;
template <> struct WrapU<int8_t> { using U = uint8_t; };
template <> struct WrapU<int16_t> { using U = uint16_t; };
template <> struct WrapU<int32_t> { using U = uint32_t; };
template <> struct WrapU<int64_t> { using U = uint64_t; };
template <> struct WrapU<uint8_t> { using U = uint8_t; };
template <> struct WrapU<uint16_t> { using U = uint16_t; };
template <> struct WrapU<uint32_t> { using U = uint32_t; };
template <> struct WrapU<uint64_t> { using U = uint64_t; };

// num-traits 0.2.19 float -> int: truncate toward zero, None outside the exclusive range
template <typename T> __device__ inline bool from_f64(double x, T &out);
#define NDS_FROM_F64_SMALL_SIGNED(TY, MINV, MAXV)                                  \
    template <> __device__ inline bool from_f64<TY>(double x, TY &out) {           \
        if (x > (double)(MINV)-1.0 && x < (double)(MAXV) + 1.0) {                  \
            out = (TY)__double2ll_rz(x);                                           \
            return true;                                                           \
        }                                                                          \
        return false;                                                              \
    }
#define NDS_FROM_F64_SMALL_UNSIGNED(TY, MAXV)                                      \
    template <> __device__ inline bool from_f64<TY>(double x, TY &out) {           \
        if (x > -1.0 && x < (double)(MAXV) + 1.0) {                                \
            out = (TY)__double2ull_rz(x);                                          \
            return true;                                                           \
        }                                                                          \
        return false;                                                              \
    }
NDS_FROM_F64_SMALL_SIGNED(int8_t, -128, 127)
NDS_FROM_F64_SMALL_SIGNED(int16_t, -32768, 32767)
NDS_FROM_F64_SMALL_SIGNED(int32_t, -2147483648.0, 2147483647.0)
NDS_FROM_F64_SMALL_UNSIGNED(uint8_t, 255)
NDS_FROM_F64_SMALL_UNSIGNED(uint16_t, 65535)
NDS_FROM_F64_SMALL_UNSIGNED(uint32_t, 4294967295.0)
template <> __device__ inline bool from_f64<int64_t>(double x, int64_t &out) {
    if (x >= -9223372036854775808.0 && x < 9223372036854775808.0) {
        out = __double2ll_rz(x);
        return true;
    }
    return false;
}
template <> __device__ inline bool from_f64<uint64_t>(double x, uint64_t &out) {
    if (x > -1.0 && x < 18446744073709551616.0) {
        out = __double2ull_rz(x);
        return true;
    }
    return false;
}
#undef NDS_FROM_F64_SMALL_SIGNED
#undef NDS_FROM_F64_SMALL_UNSIGNED

template <typename T> __device__ inline double to_f64(T v) { return (double)v; }
template <> __device__ inline double to_f64<int64_t>(int64_t v) { return __ll2double_rn(v); }
template <> __device__ inline double to_f64<uint64_t>(uint64_t v) { return __ull2double_rn(v); }

template <typename T>
__device__ inline T interpolate(int interp, T lower, T higher, double frac, int *status) {
    switch (interp) {
    case NDS_LOWER: return lower;
    case NDS_HIGHER: return higher;
    case NDS_NEAREST: return frac < 0.5 ? lower : higher;
    case NDS_MIDPOINT: {
        using U = typename WrapU<T>::U;  // release-build wrapping arithmetic; `/ 2` truncates toward zero
        T diff = (T)(U)((U)higher - (U)lower);
        T half = (T)(diff / (T)2);
        return (T)(U)((U)lower + (U)half);
    }
    default: {  // NDS_LINEAR
        using U = typename WrapU<T>::U;
        double prod = __dmul_rn(frac, __dsub_rn(to_f64(higher), to_f64(lower)));
        T inc;
        if (!from_f64<T>(prod, inc)) {
            atomicOr(status, STATUS_BIT_CAST_OVERFLOW);
            return lower;
        }
        return (T)(U)((U)lower + (U)inc);
    }
    }
}
template <>
__device__ inline float interpolate<float>(int interp, float lower, float higher, double frac, int *) {
    switch (interp) {
    case NDS_LOWER: return lower;
    case NDS_HIGHER: return higher;
    case NDS_NEAREST: return frac < 0.5 ? lower : higher;
    case NDS_MIDPOINT: return __fadd_rn(lower, __fdiv_rn(__fsub_rn(higher, lower), 2.0f));
    default: {
        double prod = __dmul_rn(frac, __dsub_rn((double)higher, (double)lower));
        return __fadd_rn(lower, __double2float_rn(prod));  // N32::from_f64 == `as f32`
    }
    }
}
template <>
__device__ inline double interpolate<double>(int interp, double lower, double higher, double frac, int *) {
    switch (interp) {
    case NDS_LOWER: return lower;
    case NDS_HIGHER: return higher;
    case NDS_NEAREST: return frac < 0.5 ? lower : higher;
    case NDS_MIDPOINT: return __dadd_rn(lower, __ddiv_rn(__dsub_rn(higher, lower), 2.0));
    default: return __dadd_rn(lower, __dmul_rn(frac, __dsub_rn(higher, lower)));
    }
}

}  // namespace nds
