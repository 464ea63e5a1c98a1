"""CPU tests: pin the oracle (oracle/ndstats_oracle.c) against every known-answer test the reference
holds for the quantile path (tests/golden/reference_kats.json) and cross-check its two selectors
(faithful quickselect vs sort-and-index) and an independent numpy statement on random inputs."""
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from kat_util import assert_bit_equal, kat_array, kat_expect, load_kats, np_reference, ulp_distance

KATS = load_kats()
INTERPS = ["lower", "higher", "nearest", "midpoint", "linear"]
ALL_DTYPES = [np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64, np.float32, np.float64]


@pytest.mark.parametrize("case", KATS["quantile"], ids=lambda c: c["id"])
@pytest.mark.parametrize("select_impl", [0, 1])
def test_reference_quantile_kats(oracle, case, select_impl):
    a = kat_array(case, none_as="nan")
    st_, out, _ = oracle.quantiles_axis(a, case["axis"], case["qs"], case["interp"], case["skipnan"], select_impl)
    if "error" in case:
        assert st_ == oracle.EMPTY_INPUT
        return
    assert st_ == oracle.OK
    if case["single"]:
        out = np.take(out, 0, axis=case["axis"])  # index_axis_move(axis, 0), src/quantile/mod.rs:507
    if "expect_shape" in case:
        assert list(out.shape) == case["expect_shape"]
        return
    want = kat_expect(case)
    assert out.shape == want.shape
    got = out.astype(np.float64)
    if "abs_tol" in case:
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.nanmax(np.abs(got - want)) < case["abs_tol"]
        # and in fact exact here
        assert_bit_equal(got, want)
    else:
        assert_bit_equal(got, want)


def test_get_from_sorted_kat(oracle):
    k = KATS["get_from_sorted"]
    for i, want in k["cases"]:
        idx, val = oracle.get_many_from_sorted_i64(k["data"], [i])
        assert idx == [i] and val[0] == want


def test_partition_kats(oracle):
    arrays = KATS["partition"]["arrays"] + [KATS["doc_partition"]["data"]]
    for arr in arrays:
        a = np.array(arr, dtype=np.int64)
        n = a.size
        piv = n - 1 if arr is not KATS["doc_partition"]["data"] else KATS["doc_partition"]["pivot_index"]
        pv = a[piv]
        p = oracle.partition_mut_i64(a, piv)
        assert (a[:p] < pv).all() and a[p] == pv and (a[p + 1:] >= pv).all()
        assert sorted(a.tolist()) == sorted(arr)


def test_remove_nan_kats(oracle):
    k = KATS["remove_nan"]
    for case in k["cases"]:
        base = np.array([math.nan if v == "nan" else v for v in k["data"]], dtype=np.float64)
        view = base[::case["step"]]
        got = oracle.remove_nan_mut_f64(view)
        assert sorted(got.tolist()) == case["expect_sorted"]
        # property tests of src/maybe_nan/mod.rs:410-440: idempotent, count preserved
        again = oracle.remove_nan_mut_f64(got)
        assert again.size == got.size


@given(st.lists(st.integers(-2**63, 2**63 - 1), min_size=1, max_size=60))
@settings(max_examples=60, deadline=None)
def test_get_many_is_a_sort(oracle, xs):
    """tests/sort.rs:46-73 — every index twice == full sort, keys ascending."""
    n = len(xs)
    idx, val = oracle.get_many_from_sorted_i64(xs, list(range(n)) * 2)
    assert idx == list(range(n))
    assert val.tolist() == sorted(xs)


@given(st.lists(st.integers(-2**63, 2**63 - 1), min_size=0, max_size=40))
@settings(max_examples=60, deadline=None)
def test_bulk_equals_single_i64(oracle, xs):
    """tests/quantile.rs:281-340 — bulk quantiles_mut == per-q quantile_mut, 5 strategies, unordered qs with a duplicate."""
    qs = KATS["quickcheck_qs"]["qs"]
    v = np.array(xs, dtype=np.int64)
    for interp in INTERPS:
        st_, bulk, _ = oracle.quantiles_axis(v, 0, qs, interp)
        if v.size == 0:
            assert st_ == oracle.EMPTY_INPUT
            continue
        if st_ == oracle.PANIC:  # Linear on huge ranges: the reference's from_f64().unwrap() panics too
            assert interp == "linear"
            continue
        assert st_ == oracle.OK
        for t, q in enumerate(qs):
            s1, one, _ = oracle.quantiles_axis(v, 0, [q], interp)
            assert s1 == oracle.OK and one[0] == bulk[t]


@given(st.lists(st.integers(0, 2**64 - 1), min_size=0, max_size=49))
@settings(max_examples=40, deadline=None)
def test_bulk_equals_single_axis_u64(oracle, xs):
    """tests/quantile.rs:343-419 — square u64 matrix, Axis(0)."""
    qs = KATS["quickcheck_qs"]["qs"]
    n = int(math.floor(math.sqrt(len(xs))))
    m = np.array(xs[: n * n], dtype=np.uint64).reshape(n, n)
    for interp in INTERPS:
        st_, bulk, _ = oracle.quantiles_axis(m, 0, qs, interp)
        if m.size == 0:
            assert st_ == oracle.EMPTY_INPUT
            continue
        assert st_ == oracle.OK
        for t, q in enumerate(qs):
            s1, one, _ = oracle.quantiles_axis(m, 0, [q], interp)
            assert s1 == oracle.OK
            assert (one[0] == bulk[t]).all()


def test_error_precedence(oracle):
    """src/quantile/mod.rs:440-455: bad q (first in qs order) beats empty axis beats empty result."""
    a = np.zeros((5, 0), dtype=np.float64)
    assert oracle.quantiles_axis(a, 1, [0.5, 1.5, -1], "linear")[0::2] == (oracle.INVALID_QUANTILE, 1)
    assert oracle.quantiles_axis(a, 1, [float("nan")], "linear")[0::2] == (oracle.INVALID_QUANTILE, 0)
    assert oracle.quantiles_axis(a, 1, [0.5], "linear")[0] == oracle.EMPTY_INPUT
    s, out, _ = oracle.quantiles_axis(a, 0, [0.5, 0.6], "linear")
    assert s == oracle.OK and out.shape == (2, 0)
    s, out, _ = oracle.quantiles_axis(np.zeros((3, 4)), 1, [], "linear")
    assert s == oracle.OK and out.shape == (3, 0)
    assert oracle.quantiles_axis(a, 1, [0.5], "linear", skipnan=True)[0] == oracle.EMPTY_INPUT
    assert oracle.quantiles_axis(np.array([1.0, np.nan]), 0, [0.5], "lower")[0] == oracle.NAN_IN_INPUT


@pytest.mark.parametrize("dtype", ALL_DTYPES, ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("interp", INTERPS)
def test_selectors_agree_and_match_numpy(oracle, dtype, interp):
    rng = np.random.default_rng(hash((np.dtype(dtype).name, interp)) % 2**32)
    dtype = np.dtype(dtype)
    for shape, axis in [((7, 33), 1), ((33, 7), 0), ((3, 5, 17), 1), ((64,), 0), ((1, 1), 0)]:
        if dtype.kind == "f":
            a = rng.standard_normal(shape).astype(dtype)
            a.flat[:: 7] = 0.0
        else:
            info = np.iinfo(dtype)
            # keep Linear inside the non-panicking range: half-range values
            a = rng.integers(info.min // 2, info.max // 2, size=shape, dtype=dtype, endpoint=True)
        qs = [0.0, 0.31, 0.5, 0.5, 0.999, 1.0, 0.25]
        s0, o0, _ = oracle.quantiles_axis(a, axis, qs, interp, select_impl=0)
        s1, o1, _ = oracle.quantiles_axis(a, axis, qs, interp, select_impl=1)
        assert s0 == s1 == oracle.OK
        assert_bit_equal(o0, o1)
        assert_bit_equal(o0, np_reference(a, axis, qs, interp))
        # strided + reversed view of the same data: lanes along a non-contiguous axis
        v = a[..., ::-1] if a.ndim > 1 else a[::-1]
        s2, o2, _ = oracle.quantiles_axis(v, axis, qs, interp)
        assert s2 == oracle.OK
        assert_bit_equal(o2, np_reference(v, axis, qs, interp))


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("interp", INTERPS)
def test_skipnan_matches_numpy(oracle, dtype, interp):
    rng = np.random.default_rng(7)
    a = rng.standard_normal((19, 23)).astype(dtype)
    a[rng.random(a.shape) < 0.3] = np.nan
    a[:, 3] = np.nan       # an all-NaN lane along axis 0
    a[5, :] = np.nan       # and along axis 1
    for axis in (0, 1):
        s, o, _ = oracle.quantiles_axis(a, axis, [0.37], interp, skipnan=True)
        assert s == oracle.OK
        assert_bit_equal(o, np_reference(a, axis, [0.37], interp, skipnan=True))
        nanlane = np.take(o, 0, axis=axis)[3 if axis == 0 else 5]
        assert np.isnan(nanlane)
        canonical = {4: 0x7FC00000, 8: 0x7FF8000000000000}[np.dtype(dtype).itemsize]
        assert int(np.array([nanlane]).view({4: np.uint32, 8: np.uint64}[np.dtype(dtype).itemsize])[0]) == canonical


def test_nearest_at_exact_half_and_specials(oracle):
    # fraction exactly 0.5 -> higher (src/quantile/interpolate.rs:92: `< 0.5`)
    a = np.array([10.0, 20.0])
    assert oracle.quantiles_axis(a, 0, [0.5], "nearest")[1][0] == 20.0
    # +-inf: Midpoint of (inf, inf) is inf + (inf-inf)/2 = NaN in the reference formula
    b = np.array([np.inf, np.inf, -np.inf])
    out = oracle.quantiles_axis(b, 0, [0.75], "midpoint")[1]
    assert np.isnan(out[0])
    out = oracle.quantiles_axis(b, 0, [0.0, 1.0], "lower")[1]
    assert out[0] == -np.inf and out[1] == np.inf
    # Linear on f32 goes through f64 (interpolate.rs:137-142)
    c = np.array([1.0, 1.0000001, 3.0], dtype=np.float32)
    out = oracle.quantiles_axis(c, 0, [0.3], "linear")[1]
    frac = 0.3 * 2 - 0
    want = np.float32(c[0] + np.float32(frac * (float(c[1]) - float(c[0]))))
    assert ulp_distance(out, np.array([want], dtype=np.float32)) == 0


def test_integer_linear_panics_like_reference(oracle):
    a = np.array([np.iinfo(np.int64).min, np.iinfo(np.int64).max], dtype=np.int64)
    assert oracle.quantiles_axis(a, 0, [0.75], "linear")[0] == oracle.PANIC
    assert oracle.quantiles_axis(a, 0, [0.25], "linear")[0] == oracle.OK
    # signed Midpoint wraps in release builds (interpolate.rs:122): documented, deterministic
    s, o, _ = oracle.quantiles_axis(np.array([-128, 127], dtype=np.int8), 0, [0.5], "midpoint")
    assert s == oracle.OK and o[0] == -128
