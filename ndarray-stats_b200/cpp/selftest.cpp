// Self-test of the C++ host mirror (include/ndstats_b200.hpp).
// Without a CUDA device (the build container): the parts that need none -- the dtype table, the reference's validation
// order and error strings, and that every compute call fails loudly (no CPU fallback).
// With a device: known answers from the reference's own tests and doc-tests through the C ABI.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "ndstats_b200.hpp"

namespace nb = ndstats_b200;
using namespace nb::interpolate;

static int failures = 0;
#define CHECK(cond)                                                                \
    do {                                                                           \
        if (!(cond)) {                                                             \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            ++failures;                                                            \
        }                                                                          \
    } while (0)

template <class E, class F> static bool throws(F f, E *out = nullptr) {
    try {
        f();
    } catch (const E &e) {
        if (out) *out = e;
        return true;
    } catch (...) {
        return false;
    }
    return false;
}

static void host_only() {
    static_assert(nb::dtype_of<int8_t>::value == NDS_I8 && nb::dtype_of<uint64_t>::value == NDS_U64 &&
                      nb::dtype_of<float>::value == NDS_F32 && nb::dtype_of<double>::value == NDS_F64, "dtype table");
    static_assert(Lower::code == NDS_LOWER && Higher::code == NDS_HIGHER && Nearest::code == NDS_NEAREST &&
                      Midpoint::code == NDS_MIDPOINT && Linear::code == NDS_LINEAR, "interpolation strategies");
    CHECK(nds_abi_version() == 1);
    CHECK(nds_dtype_size(NDS_F64) == 8 && nds_dtype_size(NDS_I16) == 2);
    nb::Context ctx(0);   // lazy: no device is touched yet
    double buf[6] = {1, 2, 3, 4, 5, 6};
    auto a = nb::ArrayView<double>::c_order(buf, {2, 3});
    CHECK(a.strides[0] == 3 && a.strides[1] == 1 && a.len() == 6);
    // first bad q wins over everything else (src/quantile/mod.rs:440-444; tests/quantile.rs invalid-quantile cases)
    nb::QuantileError qe(nb::QuantileError::EmptyInput, 0);
    CHECK(throws<nb::QuantileError>([&] { nb::quantile_axis_mut(ctx, a, 1, 1.1, Linear{}); }, &qe));
    CHECK(qe.kind == nb::QuantileError::InvalidQuantile && qe.q == 1.1);
    CHECK(std::strcmp(qe.what(), "1.1 is not between 0. and 1. (inclusive).") == 0);   // src/errors.rs:131-133
    CHECK(throws<nb::QuantileError>([&] { nb::quantiles_axis_mut(ctx, a, 0, {0.5, -0.1, 2.0}, Nearest{}); }, &qe));
    CHECK(qe.kind == nb::QuantileError::InvalidQuantile && qe.q == -0.1);
    // then the empty axis (mod.rs:446-449) -- also with a bad q present: the q check comes first
    auto e = nb::ArrayView<double>::c_order(buf, {2, 0});
    CHECK(throws<nb::QuantileError>([&] { nb::quantile_axis_mut(ctx, e, 1, 0.5, Linear{}); }, &qe));
    CHECK(qe.kind == nb::QuantileError::EmptyInput && std::strcmp(qe.what(), "Empty input.") == 0);
    CHECK(throws<nb::QuantileError>([&] { nb::quantile_axis_mut(ctx, e, 1, 1.5, Linear{}); }, &qe));
    CHECK(qe.kind == nb::QuantileError::InvalidQuantile);
    CHECK(throws<nb::QuantileError>([&] { nb::quantile_axis_skipnan_mut(ctx, e, 1, 0.5, Midpoint{}); }, &qe));
    CHECK(qe.kind == nb::QuantileError::EmptyInput);
    // a zero-size result is Ok(empty array) (mod.rs:451-455): no qs, or no lanes
    auto r0 = nb::quantiles_axis_mut(ctx, a, 1, {}, Linear{});
    CHECK(r0.data.empty() && r0.shape.size() == 2 && r0.shape[0] == 2 && r0.shape[1] == 0);
    auto z = nb::ArrayView<double>::c_order(buf, {0, 3});
    auto r1 = nb::quantile_axis_mut(ctx, z, 1, 0.5, Linear{});
    CHECK(r1.data.empty() && r1.shape.size() == 1 && r1.shape[0] == 0);
    // axis out of bounds: the reference panics
    CHECK(throws<nb::PreconditionViolated>([&] { nb::quantile_axis_mut(ctx, a, 2, 0.5, Linear{}); }));
    // min / max of nothing (mod.rs:287, 322, 336-345)
    auto none = nb::ArrayView<double>::c_order(buf, {0});
    nb::MinMaxError me(nb::MinMaxError::UndefinedOrder);
    CHECK(throws<nb::MinMaxError>([&] { nb::min(ctx, none); }, &me) && me.kind == nb::MinMaxError::EmptyInput);
    CHECK(throws<nb::MinMaxError>([&] { nb::argmax(ctx, none); }, &me) && me.kind == nb::MinMaxError::EmptyInput);
    CHECK(throws<nb::EmptyInput>([&] { nb::argmin_skipnan(ctx, none); }));
    CHECK(std::isnan(nb::max_skipnan(ctx, none)));
    CHECK(std::strcmp(nb::MinMaxError(nb::MinMaxError::UndefinedOrder).what(), "Undefined ordering between a tested pair of values.") == 0);
    // Sort1dExt preconditions
    int64_t v[4] = {1, 3, 2, 10};
    auto s = nb::ArrayView<int64_t>::c_order(v, {4});
    CHECK(nb::get_many_from_sorted_mut(ctx, s, {}).empty());
    CHECK(throws<nb::PreconditionViolated>([&] { nb::partition_mut(ctx, s, 4); }));
}

static void with_device(nb::Context &ctx) {
    // tests/sort.rs:35-44
    int64_t v[4] = {1, 3, 2, 10};
    auto s = nb::ArrayView<int64_t>::c_order(v, {4});
    CHECK(nb::get_from_sorted_mut(ctx, s, 2) == 3 && nb::get_from_sorted_mut(ctx, s, 1) == 2 && nb::get_from_sorted_mut(ctx, s, 3) == 10);
    auto many = nb::get_many_from_sorted_mut(ctx, s, {3, 0});
    CHECK(many.size() == 2 && many.begin()->first == 0 && many.at(0) == 1 && many.at(3) == 10);
    CHECK(v[0] == 1 && v[1] == 3 && v[2] == 2 && v[3] == 10);   // left as it was
    CHECK(throws<nb::PreconditionViolated>([&] { nb::get_from_sorted_mut(ctx, s, 4); }));
    // index 0.5 * 3 = 1.5 between the order statistics 2 and 3 (src/quantile/interpolate.rs:64-144)
    CHECK(nb::quantile_mut(ctx, s, 0.5, Lower{}) == 2 && nb::quantile_mut(ctx, s, 0.5, Higher{}) == 3);
    CHECK(nb::quantile_mut(ctx, s, 0.5, Nearest{}) == 3);    // a fraction of exactly 0.5 goes up
    CHECK(nb::quantile_mut(ctx, s, 0.5, Midpoint{}) == 2 && nb::quantile_mut(ctx, s, 0.5, Linear{}) == 2);   // integer arithmetic truncates
    // doc-test of partition_mut (src/sort.rs:66-88)
    int32_t p[5] = {3, 1, 4, 5, 2};
    auto pv = nb::ArrayView<int32_t>::c_order(p, {5});
    const size_t at = nb::partition_mut(ctx, pv, 2);
    CHECK(at == 3 && p[3] == 4);
    for (size_t i = 0; i < at; ++i) CHECK(p[i] < 4);
    for (size_t i = at + 1; i < 5; ++i) CHECK(p[i] >= 4);
    // lanes along either axis, f64 Linear: rows {1,2,3} and {4,5,6}
    double m[6] = {1, 2, 3, 4, 5, 6};
    auto a = nb::ArrayView<double>::c_order(m, {2, 3});
    auto rows = nb::quantile_axis_mut(ctx, a, 1, 0.5, Linear{});
    CHECK(rows.shape.size() == 1 && rows.shape[0] == 2 && rows.data[0] == 2.0 && rows.data[1] == 5.0);
    auto cols = nb::quantiles_axis_mut(ctx, a, 0, {0.0, 0.5, 1.0}, Linear{});
    CHECK(cols.shape[0] == 3 && cols.shape[1] == 3 && cols.data[0] == 1.0 && cols.data[3] == 2.5 && cols.data[8] == 6.0);
    // skipnan: NaNs do not count; a lane of nothing but NaNs gives NaN (src/quantile/mod.rs:530-541)
    const double nan = std::nan("");
    double k[6] = {nan, 1.0, 3.0, nan, nan, nan};
    auto kv = nb::ArrayView<double>::c_order(k, {2, 3});
    auto sk = nb::quantile_axis_skipnan_mut(ctx, kv, 1, 0.5, Midpoint{});
    CHECK(sk.data[0] == 2.0 && std::isnan(sk.data[1]));
    CHECK(throws<nb::PreconditionViolated>([&] { nb::quantile_axis_mut(ctx, kv, 1, 0.5, Midpoint{}); }));   // N64 cannot hold a NaN
    // min / max family (tests/quantile.rs:12-169)
    CHECK(nb::min(ctx, a) == 1.0 && nb::max(ctx, a) == 6.0);
    auto am = nb::argmax(ctx, a);
    CHECK(am.size() == 2 && am[0] == 1 && am[1] == 2);
    CHECK(throws<nb::MinMaxError>([&] { nb::min(ctx, kv); }));
    CHECK(nb::min_skipnan(ctx, kv) == 1.0 && nb::max_skipnan(ctx, kv) == 3.0);
    auto ai = nb::argmin_skipnan(ctx, kv);
    CHECK(ai[0] == 0 && ai[1] == 1);
}

int main() {
    host_only();
    nb::Context ctx(0);
    bool have_device = true;
    try {
        ctx.handle();
    } catch (const nb::DeviceError &e) {
        have_device = false;
        CHECK(e.status == NDS_CUDA_ERROR);
        // the product path fails loudly: nothing is computed on the CPU
        double m[3] = {1, 2, 3};
        auto a = nb::ArrayView<double>::c_order(m, {3});
        CHECK(throws<nb::DeviceError>([&] { nb::quantile_mut(ctx, a, 0.5, Linear{}); }));
        CHECK(throws<nb::DeviceError>([&] { nb::min(ctx, a); }));
        CHECK(throws<nb::DeviceError>([&] { nb::partition_mut(ctx, a, 1); }));
    }
    if (have_device) with_device(ctx);
    std::printf("cpp mirror self-test: %s (%s)\n", failures ? "FAILED" : "ok", have_device ? "with a CUDA device" : "host-only checks: no CUDA device");
    return failures ? 1 : 0;
}
