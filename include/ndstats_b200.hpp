// C++17 host mirror of the reference's trait surface for the bulk quantile path, header-only, over the C ABI of
// ndstats_b200.h.  Same method names, argument meaning and error behaviour as rust-ndarray/ndarray-stats v0.7:
//
//   QuantileExt::{quantiles_axis_mut, quantile_axis_mut, quantile_axis_skipnan_mut}   src/quantile/mod.rs:417-544
//   QuantileExt::{argmin, argmax, min, max} and the *_skipnan variants                 src/quantile/mod.rs:283-409
//   Quantile1dExt::{quantile_mut, quantiles_mut}                                        src/quantile/mod.rs:612-633
//   Sort1dExt::{get_from_sorted_mut, get_many_from_sorted_mut, partition_mut}           src/sort.rs:99-168
//   interpolate::{Lower, Higher, Nearest, Midpoint, Linear}                             src/quantile/interpolate.rs:51-62
//   QuantileError, MinMaxError, EmptyInput                                              src/errors.rs:5-143
//
// The Rust toolchain is absent from the build image, so this is the compiled-language mirror a maintainer can read next
// to the Rust shim in ndarray-stats_b200/rust/ (INTEGRATION.md); the Python mirror (ndarray-stats_b200/__init__.py) is what
// the parity tests drive.  Differences from the reference, all stated in ndstats_b200.h: the array is left unmodified by
// the quantile calls (the reference leaves its order unspecified), places where the reference panics throw
// `PreconditionViolated`, and there is no CPU fallback: without a CUDA device every compute call throws `DeviceError`.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <exception>
#include <limits>
#include <map>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "ndstats_b200.h"

namespace ndstats_b200 {

// ---- interpolate::{Lower, Higher, Nearest, Midpoint, Linear}: unit structs, as in the reference ---------------------
namespace interpolate {
struct Lower { static constexpr nds_interp code = NDS_LOWER; };
struct Higher { static constexpr nds_interp code = NDS_HIGHER; };
struct Nearest { static constexpr nds_interp code = NDS_NEAREST; };
struct Midpoint { static constexpr nds_interp code = NDS_MIDPOINT; };
struct Linear { static constexpr nds_interp code = NDS_LINEAR; };
}  // namespace interpolate

// ---- errors (src/errors.rs): Display strings are the reference's -------------------------------------------------------
struct EmptyInput : std::exception {
    const char *what() const noexcept override { return "Empty input."; }
};
struct QuantileError : std::exception {
    enum Kind { EmptyInput, InvalidQuantile } kind;
    double q;  // the offending quantile (InvalidQuantile)
    std::string msg;
    QuantileError(Kind k, double qq) : kind(k), q(qq) {
        if (k == EmptyInput) {
            msg = "Empty input.";
        } else {
            char b[64];
            std::snprintf(b, sizeof b, "%g", qq);
            msg = std::string(b) + " is not between 0. and 1. (inclusive).";
        }
    }
    const char *what() const noexcept override { return msg.c_str(); }
};
struct MinMaxError : std::exception {
    enum Kind { EmptyInput, UndefinedOrder } kind;
    explicit MinMaxError(Kind k) : kind(k) {}
    const char *what() const noexcept override {
        return kind == EmptyInput ? "Empty input." : "Undefined ordering between a tested pair of values.";
    }
};
// where the reference panics: axis out of bounds, index >= n, NaN in an N32 / N64 array, `from_f64(..).unwrap()`
struct PreconditionViolated : std::logic_error {
    nds_status status;
    PreconditionViolated(nds_status s, const std::string &m) : std::logic_error(m), status(s) {}
};
// CUDA / NCCL failures, unsupported calls, and -- by design -- a box without a GPU
struct DeviceError : std::runtime_error {
    nds_status status;
    DeviceError(nds_status s, const std::string &m) : std::runtime_error(m), status(s) {}
};

// ---- element types the device path covers -----------------------------------------------------------
template <class T> struct dtype_of;
template <> struct dtype_of<int8_t> { static constexpr nds_dtype value = NDS_I8; };
template <> struct dtype_of<int16_t> { static constexpr nds_dtype value = NDS_I16; };
template <> struct dtype_of<int32_t> { static constexpr nds_dtype value = NDS_I32; };
template <> struct dtype_of<int64_t> { static constexpr nds_dtype value = NDS_I64; };
template <> struct dtype_of<uint8_t> { static constexpr nds_dtype value = NDS_U8; };
template <> struct dtype_of<uint16_t> { static constexpr nds_dtype value = NDS_U16; };
template <> struct dtype_of<uint32_t> { static constexpr nds_dtype value = NDS_U32; };
template <> struct dtype_of<uint64_t> { static constexpr nds_dtype value = NDS_U64; };
template <> struct dtype_of<float> { static constexpr nds_dtype value = NDS_F32; };   // N32 (raw f32 on skipnan calls)
template <> struct dtype_of<double> { static constexpr nds_dtype value = NDS_F64; };  // N64 / f64

// ---- a view in ndarray's terms: pointer to the logical first element, shape, strides in ELEMENTS (may be negative) ---------
template <class T> struct ArrayView {
    T *ptr = nullptr;
    std::vector<size_t> shape;
    std::vector<ptrdiff_t> strides;
    static ArrayView c_order(T *p, std::vector<size_t> shp) {
        ArrayView v;
        v.ptr = p;
        v.strides.assign(shp.size(), 1);
        ptrdiff_t acc = 1;
        for (size_t d = shp.size(); d-- > 0;) {
            v.strides[d] = acc;
            acc *= (ptrdiff_t)shp[d];
        }
        v.shape = std::move(shp);
        return v;
    }
    size_t ndim() const { return shape.size(); }
    size_t len() const {
        size_t n = 1;
        for (size_t s : shape) n *= s;
        return n;
    }
};
// an owned C-order result (`Array<A, D>` of the reference, src/quantile/mod.rs:451-455, 469)
template <class T> struct Array {
    std::vector<T> data;
    std::vector<size_t> shape;
};

// ---- context: one per host thread and GPU.  Created lazily, so that argument errors surface without a device ------------
class Context {
public:
    explicit Context(int device = 0) : device_(device) {}
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    ~Context() {
        if (h_) nds_ctx_destroy(h_);
    }
    nds_ctx *handle() {
        if (!h_) {
            const nds_status st = nds_ctx_create(device_, &h_);
            if (st != NDS_OK) {
                h_ = nullptr;
                throw DeviceError(st, "nds_ctx_create failed: no usable CUDA device (ndstats_b200 has no CPU fallback)");
            }
        }
        return h_;
    }
    std::string last_path() { return h_ ? nds_ctx_last_path(h_) : "none"; }

private:
    int device_;
    nds_ctx *h_ = nullptr;
};

namespace detail {
[[noreturn]] inline void raise(nds_ctx *h, nds_status st) {
    const std::string msg = h ? nds_ctx_last_error(h) : nds_status_string(st);
    switch (st) {
    case NDS_NAN_IN_INPUT:
    case NDS_CAST_OVERFLOW:
    case NDS_INDEX_OUT_OF_BOUNDS:
    case NDS_BAD_ARGUMENT: throw PreconditionViolated(st, msg);
    default: throw DeviceError(st, msg);
    }
}
// the reference's validation order (src/quantile/mod.rs:440-455, 522-528), without a device
inline size_t validate(size_t ndim, const size_t *shape, size_t axis, const double *qs, size_t nq, nds_interp interp) {
    size_t bad = 0, result = 0;
    const nds_status st = nds_validate_quantile_args((int)ndim, shape, (int)axis, qs, nq, interp, &bad, &result);
    if (st == NDS_INVALID_QUANTILE) throw QuantileError(QuantileError::InvalidQuantile, qs[bad]);
    if (st == NDS_EMPTY_INPUT) throw QuantileError(QuantileError::EmptyInput, 0.0);
    if (st != NDS_OK) raise(nullptr, st);
    return result;
}
template <class T>
Array<T> quantiles_call(Context &ctx, const ArrayView<T> &a, size_t axis, const std::vector<double> &qs, nds_interp interp, bool skipnan) {
    const size_t result = validate(a.ndim(), a.shape.data(), axis, qs.data(), qs.size(), interp);
    Array<T> out;
    out.shape = a.shape;
    out.shape[axis] = qs.size();   // src/quantile/mod.rs:451-452
    out.data.resize(result);
    if (result == 0) return out;   // Ok(empty array), src/quantile/mod.rs:453-455
    nds_ctx *h = ctx.handle();
    size_t bad = 0;
    const nds_status st = nds_quantiles_axis(h, dtype_of<std::remove_const_t<T>>::value, a.ptr, (int)a.ndim(), a.shape.data(),
                                             a.strides.data(), (int)axis, qs.data(), qs.size(), interp, skipnan ? 1 : 0,
                                             out.data.data(), &bad);
    if (st == NDS_INVALID_QUANTILE) throw QuantileError(QuantileError::InvalidQuantile, qs[bad]);
    if (st == NDS_EMPTY_INPUT) throw QuantileError(QuantileError::EmptyInput, 0.0);
    if (st != NDS_OK) raise(h, st);
    return out;
}
template <class T> Array<T> drop_axis(Array<T> r, size_t axis) {   // index_axis_move(axis, 0), src/quantile/mod.rs:507
    r.shape.erase(r.shape.begin() + (ptrdiff_t)axis);
    return r;
}
}  // namespace detail

// ---- QuantileExt ----------------------------------------------------------------------------------------------------------
template <class T, class I>
Array<T> quantiles_axis_mut(Context &ctx, const ArrayView<T> &a, size_t axis, const std::vector<double> &qs, I) {
    return detail::quantiles_call(ctx, a, axis, qs, I::code, false);
}
template <class T, class I> Array<T> quantile_axis_mut(Context &ctx, const ArrayView<T> &a, size_t axis, double q, I) {
    return detail::drop_axis(detail::quantiles_call(ctx, a, axis, std::vector<double>{q}, I::code, false), axis);
}
template <class T, class I> Array<T> quantile_axis_skipnan_mut(Context &ctx, const ArrayView<T> &a, size_t axis, double q, I) {
    static_assert(std::is_floating_point<T>::value, "MaybeNan over raw floats; Option<T> lanes go through nds_quantiles_axis_masked");
    return detail::drop_axis(detail::quantiles_call(ctx, a, axis, std::vector<double>{q}, I::code, true), axis);
}

// ---- Quantile1dExt --------------------------------------------------------------------------------------------------------
template <class T, class I> T quantile_mut(Context &ctx, const ArrayView<T> &a, double q, I i) {
    if (a.ndim() != 1) throw PreconditionViolated(NDS_BAD_ARGUMENT, "Quantile1dExt is defined on 1-D arrays");
    return quantile_axis_mut(ctx, a, 0, q, i).data.at(0);
}
template <class T, class I> Array<T> quantiles_mut(Context &ctx, const ArrayView<T> &a, const std::vector<double> &qs, I i) {
    if (a.ndim() != 1) throw PreconditionViolated(NDS_BAD_ARGUMENT, "Quantile1dExt is defined on 1-D arrays");
    return quantiles_axis_mut(ctx, a, 0, qs, i);
}

// ---- Sort1dExt ------------------------------------------------------------------------------------------------------------
// ascending keys, as the reference's IndexMap iterates (src/sort.rs:34-35)
template <class T> std::map<size_t, T> get_many_from_sorted_mut(Context &ctx, const ArrayView<T> &a, std::vector<size_t> indexes) {
    if (a.ndim() != 1) throw PreconditionViolated(NDS_BAD_ARGUMENT, "Sort1dExt is defined on 1-D arrays");
    std::map<size_t, T> out;
    if (indexes.empty()) return out;
    std::vector<T> vals(indexes.size());
    nds_ctx *h = ctx.handle();
    const nds_status st = nds_get_many_from_sorted(h, dtype_of<std::remove_const_t<T>>::value, a.ptr, a.shape[0],
                                                   a.shape[0] ? a.strides[0] : 1, indexes.data(), indexes.size(), vals.data());
    if (st != NDS_OK) detail::raise(h, st);
    for (size_t j = 0; j < indexes.size(); ++j) out[indexes[j]] = vals[j];
    return out;
}
template <class T> T get_from_sorted_mut(Context &ctx, const ArrayView<T> &a, size_t i) {
    return get_many_from_sorted_mut(ctx, a, std::vector<size_t>{i}).at(i);
}
// in place; returns the new index of the pivot value (src/sort.rs:133-168)
template <class T> size_t partition_mut(Context &ctx, const ArrayView<T> &a, size_t pivot_index) {
    static_assert(!std::is_const<T>::value, "partition_mut works in place");
    if (a.ndim() != 1) throw PreconditionViolated(NDS_BAD_ARGUMENT, "Sort1dExt is defined on 1-D arrays");
    if (pivot_index >= a.shape[0]) throw PreconditionViolated(NDS_INDEX_OUT_OF_BOUNDS, "pivot_index >= n (the reference panics)");
    nds_ctx *h = ctx.handle();
    size_t new_index = 0;
    const nds_status st = nds_partition(h, dtype_of<T>::value, a.ptr, a.shape[0], a.strides[0], pivot_index, &new_index);
    if (st != NDS_OK) detail::raise(h, st);
    return new_index;
}

// ---- QuantileExt: the min / max family ----------------------------------------------------------------------------------------
namespace detail {
template <class T>
bool minmax_call(Context &ctx, const ArrayView<T> &a, bool want_max, bool skipnan, T *value, std::vector<size_t> *index, bool arg_call) {
    if (index) index->assign(a.ndim(), 0);
    if (a.len() == 0) {   // no device needed to know that (src/quantile/mod.rs:287, 322)
        if (arg_call) throw EmptyInput();
        if (skipnan && std::is_floating_point<T>::value) return false;   // min / max_skipnan of nothing: NaN
        throw MinMaxError(MinMaxError::EmptyInput);
    }
    nds_ctx *h = ctx.handle();
    const nds_status st = nds_minmax(h, dtype_of<std::remove_const_t<T>>::value, a.ptr, (int)a.ndim(), a.shape.data(),
                                     a.strides.data(), want_max ? 1 : 0, skipnan ? 1 : 0, value, index ? index->data() : nullptr);
    if (st == NDS_OK) return true;
    if (st == NDS_UNDEFINED_ORDER) throw MinMaxError(MinMaxError::UndefinedOrder);
    if (st == NDS_EMPTY_INPUT) {   // skipnan: nothing but NaNs
        if (arg_call) throw EmptyInput();
        return false;
    }
    raise(h, st);
}
}  // namespace detail
template <class T> std::vector<size_t> argmin(Context &ctx, const ArrayView<T> &a) {
    std::vector<size_t> idx;
    if (a.len() == 0) throw MinMaxError(MinMaxError::EmptyInput);
    detail::minmax_call<T>(ctx, a, false, false, nullptr, &idx, false);
    return idx;
}
template <class T> std::vector<size_t> argmax(Context &ctx, const ArrayView<T> &a) {
    std::vector<size_t> idx;
    if (a.len() == 0) throw MinMaxError(MinMaxError::EmptyInput);
    detail::minmax_call<T>(ctx, a, true, false, nullptr, &idx, false);
    return idx;
}
template <class T> std::vector<size_t> argmin_skipnan(Context &ctx, const ArrayView<T> &a) {
    std::vector<size_t> idx;
    detail::minmax_call<T>(ctx, a, false, true, nullptr, &idx, true);
    return idx;
}
template <class T> std::vector<size_t> argmax_skipnan(Context &ctx, const ArrayView<T> &a) {
    std::vector<size_t> idx;
    detail::minmax_call<T>(ctx, a, true, true, nullptr, &idx, true);
    return idx;
}
template <class T> T min(Context &ctx, const ArrayView<T> &a) {
    std::remove_const_t<T> v{};
    detail::minmax_call<T>(ctx, a, false, false, &v, nullptr, false);
    return v;
}
template <class T> T max(Context &ctx, const ArrayView<T> &a) {
    std::remove_const_t<T> v{};
    detail::minmax_call<T>(ctx, a, true, false, &v, nullptr, false);
    return v;
}
// NaN when no element is left (src/quantile/mod.rs:336-345, 399-409)
template <class T> T min_skipnan(Context &ctx, const ArrayView<T> &a) {
    std::remove_const_t<T> v{};
    if (!detail::minmax_call<T>(ctx, a, false, true, &v, nullptr, false)) return std::is_floating_point<T>::value ? std::numeric_limits<std::remove_const_t<T>>::quiet_NaN() : v;
    return v;
}
template <class T> T max_skipnan(Context &ctx, const ArrayView<T> &a) {
    std::remove_const_t<T> v{};
    if (!detail::minmax_call<T>(ctx, a, true, true, &v, nullptr, false)) return std::is_floating_point<T>::value ? std::numeric_limits<std::remove_const_t<T>>::quiet_NaN() : v;
    return v;
}

}  // namespace ndstats_b200
