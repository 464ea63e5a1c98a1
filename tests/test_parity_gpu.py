"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Bars: bit-exact for every order statistic and for Lower/Higher/Nearest/Midpoint (-0.0 folded onto +0.0,
SURVEY.md note Z); Linear <= 1 ulp (and in fact 0 — asserted as such where noted)."""
import math
import os
import zlib

import numpy as np
import pytest

from kat_util import assert_bit_equal, kat_array, kat_expect, load_kats, np_reference, ulp_distance

pytestmark = pytest.mark.gpu

KATS = load_kats()
INTERPS = ["lower", "higher", "nearest", "midpoint", "linear"]
ALL_DTYPES = [np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64, np.float32, np.float64]
LINEAR_ULP_TOL = 1  # north_star: "Linear must match within 1 ulp"


def check(ctx, oracle, a, axis, qs, interp, skipnan=False, select_impl=0):
    st, want, _ = oracle.quantiles_axis(a, axis, qs, interp, skipnan, select_impl)
    assert st == oracle.OK, f"oracle status {st}"
    got = ctx.quantiles_axis(a, axis, qs, interp, skipnan=skipnan)
    if interp == "linear" and got.dtype.kind == "f":
        assert ulp_distance(got, want) <= LINEAR_ULP_TOL
        assert ulp_distance(got, want) == 0  # stronger than required: same operations, same roundings
    else:
        assert_bit_equal(got, want, msg=f"{a.dtype} {a.shape} axis={axis} {interp} path={ctx.last_path}")
    return got


# ---- the reference's own known-answer tests, through the CUDA path ------------------------------------
@pytest.mark.parametrize("case", KATS["quantile"], ids=lambda c: c["id"])
def test_reference_kats(nds, gpu_ctx, case):
    data = kat_array(case, none_as="mask")
    interp = case["interp"]
    if "error" in case:
        with pytest.raises(nds.EmptyInput):
            nds.quantile_axis_mut(data, case["axis"], case["qs"][0], interp, ctx=gpu_ctx)
        return
    if isinstance(data, tuple):  # Option<i32> lanes (tests/quantile.rs:216-246,260-268)
        out, valid = nds.quantile_axis_skipnan_mut(data, case["axis"], case["qs"][0], interp, ctx=gpu_ctx)
        want = kat_expect(case)
        assert out.shape == want.shape
        assert np.array_equal(valid, ~np.isnan(want))
        assert np.array_equal(out[valid].astype(np.float64), want[valid])
        return
    if case["skipnan"]:
        out = nds.quantile_axis_skipnan_mut(data, case["axis"], case["qs"][0], interp, ctx=gpu_ctx)
    elif case["single"]:
        out = nds.quantile_axis_mut(data, case["axis"], case["qs"][0], interp, ctx=gpu_ctx)
    else:
        out = nds.quantiles_axis_mut(data, case["axis"], case["qs"], interp, ctx=gpu_ctx)
    if "expect_shape" in case:
        assert list(out.shape) == case["expect_shape"]
        return
    want = kat_expect(case)
    assert out.shape == want.shape
    assert_bit_equal(np.asarray(out, dtype=np.float64), want)


def test_quantile1d_and_errors(nds, gpu_ctx):
    a = np.array([1, 3, 22, 10], dtype=np.int32)
    assert nds.quantile_mut(a, 1.0, nds.Lower, ctx=gpu_ctx) == 22
    assert nds.quantile_mut(a, 0.0, nds.Lower(), ctx=gpu_ctx) == 1
    assert np.array_equal(a, [1, 3, 22, 10])  # input untouched
    with pytest.raises(nds.EmptyInput):
        nds.quantile_mut(np.zeros(0, dtype=np.float64), 0.5, nds.Linear, ctx=gpu_ctx)
    with pytest.raises(nds.InvalidQuantile) as ei:
        nds.quantiles_mut(a, [0.5, 1.5, -2.0], nds.Linear, ctx=gpu_ctx)
    assert ei.value.q == 1.5 and "1.5 is not between 0. and 1. (inclusive)." == str(ei.value)
    with pytest.raises(nds.InvalidQuantile):  # bad q beats empty input (src/quantile/mod.rs:440-449)
        nds.quantile_mut(np.zeros(0), float("nan"), nds.Linear, ctx=gpu_ctx)
    with pytest.raises(nds.NanInInput):
        nds.quantile_mut(np.array([1.0, np.nan]), 0.5, nds.Lower, ctx=gpu_ctx)
    with pytest.raises(nds.CastOverflow):
        nds.quantile_mut(np.array([np.iinfo(np.int64).min, np.iinfo(np.int64).max]), 0.75, nds.Linear, ctx=gpu_ctx)
    with pytest.raises(IndexError):
        nds.quantile_axis_mut(a, 1, 0.5, nds.Lower, ctx=gpu_ctx)
    out = nds.quantiles_axis_mut(np.zeros((5, 0)), 0, [0.1, 0.2], nds.Lower, ctx=gpu_ctx)
    assert out.shape == (2, 0)


# ---- dtype x strategy x layout matrix, both selectors of the oracle ------------------------------------
@pytest.mark.parametrize("dtype", ALL_DTYPES, ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("interp", INTERPS)
def test_matrix(gpu_ctx, oracle, dtype, interp):
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(zlib.crc32(repr((dtype.name, interp)).encode()))
    qs = KATS["quickcheck_qs"]["qs"] + [0.31, 0.001]
    for shape, axis in [((9, 257), 1), ((257, 9), 0), ((3, 5, 130), 1), ((1000,), 0), ((1, 1), 1), ((4, 3, 2, 5), 3)]:
        if dtype.kind == "f":
            a = rng.standard_normal(shape).astype(dtype)
            a.flat[::5] = 0.0
            a.flat[1::11] = -0.0
        else:
            info = np.iinfo(dtype)
            a = rng.integers(info.min // 2, info.max // 2, size=shape, dtype=dtype, endpoint=True)
        check(gpu_ctx, oracle, a, axis, qs, interp)
        # reversed / strided views: lanes along non-contiguous axes, negative strides (tests/maybe_nan.rs:5-31 spirit)
        v = a[..., ::-2] if a.shape[-1] > 2 else a
        check(gpu_ctx, oracle, v, axis, qs, interp, select_impl=1)
        if a.ndim >= 2:
            check(gpu_ctx, oracle, np.swapaxes(a, 0, -1), a.ndim - 1 - axis if axis in (0, a.ndim - 1) else axis, qs, interp)


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("interp", INTERPS)
def test_skipnan(gpu_ctx, oracle, dtype, interp):
    rng = np.random.default_rng(11)
    a = rng.standard_normal((61, 130)).astype(dtype)
    a[rng.random(a.shape) < 0.1] = np.nan
    a[:, 7] = np.nan
    a[13, :] = np.nan
    a[2, 2] = -np.nan
    for axis in (0, 1):
        for q in (0.0, 0.5, 0.37, 1.0):
            got = check(gpu_ctx, oracle, a, axis, [q], interp, skipnan=True)
    # all-NaN lanes give the canonical quiet NaN bit pattern
    got = gpu_ctx.quantiles_axis(a, 0, [0.5], interp, skipnan=True)
    canonical = {4: 0x7FC00000, 8: 0x7FF8000000000000}[np.dtype(dtype).itemsize]
    ut = {4: np.uint32, 8: np.uint64}[np.dtype(dtype).itemsize]
    assert int(got.view(ut)[0, 7]) == canonical


def test_ties_and_extremes(gpu_ctx, oracle):
    rng = np.random.default_rng(5)
    for dtype in (np.uint32, np.int64, np.float32, np.float64):
        a = rng.integers(0, 16, size=(33, 4096)).astype(dtype)           # 16 distinct values (C5-style)
        for interp in ("lower", "higher", "midpoint"):
            check(gpu_ctx, oracle, a, 1, [0.1, 0.5, 0.9], interp, select_impl=1)
        b = np.full((4, 1000), 7, dtype=dtype)                           # all equal
        check(gpu_ctx, oracle, b, 1, [0.0, 0.5, 1.0], "nearest", select_impl=1)
    f = np.array([np.inf, -np.inf, 0.0, -0.0, 1e-45, -1e-45, 3.4e38, -3.4e38] * 5, dtype=np.float32)
    for interp in INTERPS:
        check(gpu_ctx, oracle, f, 0, [0.0, 0.13, 0.5, 0.77, 1.0], interp)
    # Nearest exactly at fraction 0.5 goes up (src/quantile/interpolate.rs:92)
    assert gpu_ctx.quantiles_axis(np.array([10.0, 20.0]), 0, [0.5], "nearest")[0] == 20.0
    # extremes of the integer ranges with the wrapping Midpoint
    for dtype in (np.int8, np.int32, np.int64, np.uint64):
        info = np.iinfo(dtype)
        e = np.array([info.min, info.max, info.min, info.max, 0], dtype=dtype)
        for interp in ("lower", "higher", "nearest", "midpoint"):
            check(gpu_ctx, oracle, e, 0, [0.0, 0.4, 0.5, 0.6, 1.0], interp)


def test_bulk_equals_single_property(nds, gpu_ctx):
    """tests/quantile.rs:281-419 — bulk == per-q for unordered qs with a duplicate, all strategies."""
    rng = np.random.default_rng(3)
    qs = KATS["quickcheck_qs"]["qs"]
    v = rng.integers(-2**62, 2**62, size=300, dtype=np.int64)
    m = rng.integers(0, 2**63, size=(17, 17), dtype=np.uint64)
    for interp in (nds.Lower, nds.Higher, nds.Nearest, nds.Midpoint, nds.Linear):
        bulk = nds.quantiles_mut(v, qs, interp, ctx=gpu_ctx)
        for t, q in enumerate(qs):
            assert nds.quantile_mut(v, q, interp, ctx=gpu_ctx) == bulk[t]
        bulk = nds.quantiles_axis_mut(m, 0, qs, interp, ctx=gpu_ctx)
        for t, q in enumerate(qs):
            assert np.array_equal(nds.quantile_axis_mut(m, 0, q, interp, ctx=gpu_ctx), bulk[t])


def test_sort1d_ext(nds, gpu_ctx):
    """tests/sort.rs:35-86 on the device: get_from_sorted == sort; get_many with every index twice == sort."""
    k = KATS["get_from_sorted"]
    a = np.array(k["data"], dtype=np.int64)
    for i, want in k["cases"]:
        assert nds.get_from_sorted_mut(a, i, ctx=gpu_ctx) == want
    rng = np.random.default_rng(9)
    xs = rng.integers(-2**63, 2**63 - 1, size=257, dtype=np.int64)
    got = nds.get_many_from_sorted_mut(xs, list(range(257)) * 2, ctx=gpu_ctx)
    assert list(got.keys()) == list(range(257))
    assert np.array_equal(np.array(list(got.values())), np.sort(xs))
    with pytest.raises(IndexError):
        nds.get_from_sorted_mut(a, 4, ctx=gpu_ctx)
    assert nds.get_many_from_sorted_mut(a, [], ctx=gpu_ctx) == {}


@pytest.mark.parametrize("path", ["generic_cta", "long_radix"])
def test_forced_paths_agree(gpu_ctx, oracle, path):
    rng = np.random.default_rng(21)
    try:
        gpu_ctx.force_path(path)
        for dtype in (np.float64, np.float32, np.int16, np.uint64):
            a = (rng.standard_normal((3, 5000)) * 1000).astype(dtype)
            for interp in INTERPS:
                check(gpu_ctx, oracle, a, 1, np.arange(1, 100) / 100.0, interp)
                assert gpu_ctx.last_path == path
            check(gpu_ctx, oracle, a.astype(np.float64) if dtype in (np.float32, np.float64) else a, 0, [0.5], "lower")
        b = rng.standard_normal(70001)
        b[rng.random(b.size) < 0.2] = np.nan
        check(gpu_ctx, oracle, b, 0, [0.01, 0.5, 0.99], "linear", skipnan=True)
        check(gpu_ctx, oracle, np.full(100, np.nan), 0, [0.5], "linear", skipnan=True)
    finally:
        gpu_ctx.force_path("auto")


def test_device_resident_and_enqueue(nds, gpu_ctx, oracle):
    import torch
    rng = np.random.default_rng(2)
    a = rng.random((128, 4096), dtype=np.float32)
    t = torch.from_numpy(a).cuda()
    out = nds.quantile_axis_mut(t, 1, 0.5, nds.Midpoint, ctx=gpu_ctx)
    st, want, _ = oracle.quantiles_axis(a, 1, [0.5], "midpoint")
    assert_bit_equal(out.cpu().numpy(), want[:, 0])
    # column lanes of the same tensor (strided), enqueue + deferred status
    out2 = gpu_ctx.quantiles_axis(t, 0, [0.25, 0.75], "linear", enqueue=True)
    gpu_ctx.synchronize()
    st, want2, _ = oracle.quantiles_axis(a, 0, [0.25, 0.75], "linear")
    assert ulp_distance(out2.cpu().numpy(), want2) == 0
    t[3, 5] = float("nan")
    gpu_ctx.quantiles_axis(t, 1, [0.5], "lower", enqueue=True)
    with pytest.raises(nds.NanInInput):
        gpu_ctx.synchronize()
    gpu_ctx.synchronize()  # status was cleared


def test_full_size_properties_c2(nds, gpu_ctx):
    """BASELINE config 2 at full size (65536 x 4096 f32, row medians, Midpoint): size-independent checks —
    the median splits each row in half (rank property), is invariant under row reversal, and the
    checksum over rows matches the checksum of per-block results."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(1002)
    t = torch.rand((65536, 4096), generator=g, device="cuda", dtype=torch.float32)
    lo = nds.quantile_axis_mut(t, 1, 0.5, nds.Lower, ctx=gpu_ctx)
    hi = nds.quantile_axis_mut(t, 1, 0.5, nds.Higher, ctx=gpu_ctx)
    mid = nds.quantile_axis_mut(t, 1, 0.5, nds.Midpoint, ctx=gpu_ctx)
    # rank property: exactly 2047 elements strictly below-or-tied-before `lo`, i.e. count(< lo) <= 2047 < count(<= lo)
    lt = (t < lo[:, None]).sum(1)
    le = (t <= lo[:, None]).sum(1)
    assert bool(((lt <= 2047) & (le > 2047)).all())
    lt = (t < hi[:, None]).sum(1)
    le = (t <= hi[:, None]).sum(1)
    assert bool(((lt <= 2048) & (le > 2048)).all())
    want_mid = lo + (hi - lo) / 2
    assert torch.equal(mid, want_mid)
    rev = nds.quantile_axis_mut(t.flip(1), 1, 0.5, nds.Midpoint, ctx=gpu_ctx)
    assert torch.equal(rev, mid)
    blocks = torch.cat([nds.quantile_axis_mut(t[i:i + 8192], 1, 0.5, nds.Midpoint, ctx=gpu_ctx) for i in range(0, 65536, 8192)])
    assert torch.equal(blocks, mid)


SHORT_DTYPES = [np.float32, np.float64, np.int32, np.uint32, np.int64, np.uint64]


@pytest.mark.parametrize("dtype", SHORT_DTYPES, ids=lambda d: np.dtype(d).name)
def test_short_bitslice_path(gpu_ctx, oracle, dtype):
    """The warp-per-lane bit-sliced kernel on every layout class it accepts, incl. data that needs the
    later 16-bit rounds (clustered / tie-heavy values) and lane lengths that are not multiples of 1024."""
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(zlib.crc32(repr(dtype.name).encode()))
    qs_sets = {"midpoint": [0.5], "linear": [0.25, 0.5, 0.75], "lower": [0.0, 0.1, 0.5, 0.9, 1.0],
               "higher": [0.999, 0.001, 0.5, 0.5], "nearest": [0.3, 0.7]}
    vec = 16 // dtype.itemsize
    try:
        gpu_ctx.force_path("short_bitslice")
        for n in (4096, 2048, 1024, 64, 100, 4096 - vec, 1028, 2052, 8192, 8192 - vec, 5000 // vec * vec):
            lanes = 150
            kinds = ["normal", "uniform", "clustered", "ties", "equal", "extremes"]
            for kind in kinds:
                if dtype.kind == "f":
                    if kind == "normal":
                        a = rng.standard_normal((lanes, n))
                    elif kind == "uniform":
                        a = rng.random((lanes, n))
                    elif kind == "clustered":
                        a = 1000.0 + rng.random((lanes, n)) * 1e-4
                    elif kind == "ties":
                        a = rng.integers(-8, 8, size=(lanes, n)).astype(np.float64)
                    elif kind == "equal":
                        a = np.full((lanes, n), -3.25)
                    else:
                        a = rng.standard_normal((lanes, n))
                        a[:, ::7] = np.inf
                        a[:, 1::13] = -np.inf
                        a[:, 2::17] = 0.0
                        a[:, 3::19] = -0.0
                        a[:, 5::23] = np.finfo(dtype).tiny / 4   # subnormal
                    a = a.astype(dtype)
                else:
                    info = np.iinfo(dtype)
                    if kind in ("normal", "uniform"):
                        a = rng.integers(info.min // 2, info.max // 2, size=(lanes, n), dtype=dtype, endpoint=True)
                    elif kind == "clustered":
                        a = (rng.integers(0, 1000, size=(lanes, n)) + 10**6).astype(dtype)
                    elif kind == "ties":
                        a = rng.integers(0, 16, size=(lanes, n)).astype(dtype)
                    elif kind == "equal":
                        a = np.full((lanes, n), 7, dtype=dtype)
                    else:
                        a = rng.choice(np.array([info.min, info.max, 0, 1], dtype=dtype), size=(lanes, n))
                for interp, qs in qs_sets.items():
                    if dtype.kind != "f" and interp == "linear" and kind == "extremes":
                        continue  # from_f64 overflow is covered elsewhere
                    check(gpu_ctx, oracle, a, 1, qs, interp, select_impl=1 if kind in ("ties", "equal", "extremes") else 0)
                    assert gpu_ctx.last_path == "short_bitslice"
        # a 3-d array whose last axis is the lane; and a column-sliced (pitched) 2-d view with aligned pitch
        a = rng.standard_normal((5, 40, 512)).astype(dtype) if dtype.kind == "f" else rng.integers(0, 1000, size=(5, 40, 512)).astype(dtype)
        check(gpu_ctx, oracle, a, 2, [0.5, 0.9], "lower")
        assert gpu_ctx.last_path == "short_bitslice"
        wide = rng.standard_normal((200, 1024)).astype(dtype) if dtype.kind == "f" else rng.integers(0, 10**6, size=(200, 1024)).astype(dtype)
        check(gpu_ctx, oracle, wide[:, :512], 1, [0.5], "higher")
        assert gpu_ctx.last_path == "short_bitslice"
    finally:
        gpu_ctx.force_path("auto")


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=lambda d: np.dtype(d).name)
def test_short_bitslice_nan(nds, gpu_ctx, oracle, dtype):
    rng = np.random.default_rng(77)
    a = rng.standard_normal((130, 2048)).astype(dtype)
    a[rng.random(a.shape) < 0.1] = np.nan
    a[5, :] = np.nan
    a[6, :-3] = np.nan
    a[7, 11] = np.inf
    a[8, 12] = -np.nan
    try:
        gpu_ctx.force_path("short_bitslice")
        for interp in INTERPS:
            for q in (0.0, 0.5, 0.77, 1.0):
                check(gpu_ctx, oracle, a, 1, [q], interp, skipnan=True)
                assert gpu_ctx.last_path == "short_bitslice"
        with pytest.raises(nds.NanInInput):
            gpu_ctx.quantiles_axis(a, 1, [0.5], "lower")
        b = np.where(np.isnan(a), 0, a)
        b[3, 3] = np.inf
        check(gpu_ctx, oracle, b, 1, [0.5, 1.0], "linear")
    finally:
        gpu_ctx.force_path("auto")


@pytest.mark.parametrize("dtype", [np.int64, np.uint64, np.int32, np.uint32, np.float64], ids=lambda d: np.dtype(d).name)
def test_short_bitslice_informative_windows(gpu_ctx, oracle, dtype):
    """The 8-byte types and the 4-byte integers pick the 16-bit window of a round from the bits that tell the lane's
    elements apart (nds_short.cu, SMART): small magnitudes in a wide type, sign-extended negatives, constant leading
    bits, values that differ only far down, a first block that under-states the lane (window re-proposed), NaNs that
    appear only in later blocks, many distinct tie values (several rounds)."""
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(zlib.crc32(repr(("smart", dtype.name)).encode()))
    lanes = 96
    cases = []
    for n in (8192, 4096, 3000 // (16 // dtype.itemsize) * (16 // dtype.itemsize), 1024, 256):
        if dtype.kind == "f":
            base = rng.random((lanes, n))
            cases += [
                ("binade", 1.0 + base * 0.5),
                ("tiny-spread", 1e-300 * (1 + base)),
                ("mixed-sign-small", (base - 0.5) * 1e-3),
                ("few-values-random-mantissa", rng.choice(rng.standard_normal(100), size=(lanes, n))),
                ("two-values-low-bits", np.where(base < 0.5, 1.0, np.nextafter(1.0, 2.0))),
                ("signed-zeros", np.where(base < 0.3, -0.0, np.where(base < 0.6, 0.0, base))),
            ]
            late = 1.0 + base
            late[:, n // 2:] *= 1e200     # the first block under-states the range
            cases.append(("late-large", late))
            late_nan = base.copy()
            late_nan[:, n - n // 4:][rng.random((lanes, n // 4)) < 0.2] = np.nan   # NaNs only at the end of the lane
            cases.append(("late-nan", late_nan))
        else:
            info = np.iinfo(dtype)
            signed = info.min < 0
            lo16 = -8 if signed else 0
            sh = 40 if dtype.itemsize == 8 else 20
            small = rng.integers(lo16, lo16 + 16, size=(lanes, n)).astype(np.int64)
            cases += [
                ("16-values-shifted", (small * (1 << sh)).astype(dtype)),
                ("small", rng.integers(-1000 if signed else 0, 1000, size=(lanes, n)).astype(dtype)),
                ("low-bits-only", rng.integers(0, 3, size=(lanes, n)).astype(dtype) + dtype.type(info.max - 5)),
                ("high-constant", rng.integers(0, 1 << 20, size=(lanes, n)).astype(dtype) + dtype.type(info.max // 2 + 12345)),
                ("full-range", rng.integers(info.min, info.max, size=(lanes, n), dtype=dtype, endpoint=True)),
                ("100-values", rng.choice(rng.integers(info.min, info.max, size=100, dtype=dtype), size=(lanes, n))),
            ]
            late = rng.integers(0, 100, size=(lanes, n)).astype(dtype)
            late[:, n // 2:] = rng.integers(info.max // 4, info.max // 2, size=(lanes, n - n // 2), dtype=dtype)
            cases.append(("late-large", late))
            if signed:
                late = rng.integers(0, 100, size=(lanes, n)).astype(dtype)
                late[:, n - 3:] = -1
                cases.append(("late-negative", late))
                cases.append(("min-max", rng.choice(np.array([info.min, info.max, -1, 0], dtype=dtype), size=(lanes, n))))
    try:
        gpu_ctx.force_path("short_bitslice")
        for name, a in cases:
            a = np.ascontiguousarray(a.astype(dtype))
            skip = bool(dtype.kind == "f" and np.isnan(a).any())
            for interp, qs in (("lower", [0.1, 0.5, 0.9]), ("higher", [0.0, 0.5, 1.0]), ("midpoint", [0.5]), ("nearest", [0.25, 0.999])):
                check(gpu_ctx, oracle, a, 1, qs, interp, skipnan=skip, select_impl=1)
                assert gpu_ctx.last_path == "short_bitslice", (name, gpu_ctx.last_path)
    finally:
        gpu_ctx.force_path("auto")


COLS_PATHS = ("cols_bitslice",)


@pytest.mark.parametrize("dtype", SHORT_DTYPES, ids=lambda d: np.dtype(d).name)
def test_cols_bitslice_path(nds, gpu_ctx, oracle, dtype):
    """Column lanes (strided axis, contiguous neighbours): the CTA-per-strip kernel, incl. ragged strip
    widths, lane lengths that are not multiples of 32/1024, tie-heavy data (later rounds) and 3-d arrays."""
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(zlib.crc32(repr(("cols", dtype.name)).encode()))
    try:
        gpu_ctx.force_path("cols_bitslice")
        for n, cols in ((8192, 64), (4096, 37), (1000, 100), (64, 33), (2049, 40)):
            for kind in ("normal", "ties", "clustered"):
                if dtype.kind == "f":
                    a = {"normal": lambda: rng.standard_normal((n, cols)),
                         "ties": lambda: rng.integers(-4, 4, size=(n, cols)).astype(np.float64),
                         "clustered": lambda: 1000.0 + rng.random((n, cols)) * 1e-4}[kind]().astype(dtype)
                else:
                    info = np.iinfo(dtype)
                    a = {"normal": lambda: rng.integers(info.min // 2, info.max // 2, size=(n, cols), dtype=dtype),
                         "ties": lambda: rng.integers(0, 16, size=(n, cols)).astype(dtype),
                         "clustered": lambda: (rng.integers(0, 1000, size=(n, cols)) + 10**6).astype(dtype)}[kind]()
                for interp, qs in (("nearest", [0.5]), ("midpoint", [0.5, 0.25]), ("linear", [0.1, 0.9, 0.5]),
                                   ("lower", [0.0, 1.0]), ("higher", [0.33])):
                    check(gpu_ctx, oracle, a, 0, qs, interp, select_impl=1 if kind == "ties" else 0)
                    assert gpu_ctx.last_path in COLS_PATHS
        b = (rng.standard_normal((3, 500, 40)) * 100).astype(dtype)
        check(gpu_ctx, oracle, b, 1, [0.5, 0.75], "lower")
        assert gpu_ctx.last_path in COLS_PATHS
    finally:
        gpu_ctx.force_path("auto")


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=lambda d: np.dtype(d).name)
def test_cols_bitslice_skipnan(nds, gpu_ctx, oracle, dtype):
    """BASELINE config 3 in miniature: skipnan down the strided axis with 10 % NaN, all strategies."""
    rng = np.random.default_rng(303)
    a = rng.random((2000, 72)).astype(dtype)
    a[rng.random(a.shape) < 0.1] = np.nan
    a[:, 5] = np.nan
    a[:-2, 6] = np.nan
    a[7, 8] = np.inf
    a[9, 8] = -np.inf
    # a NaN whose payload lives only in the low mantissa bits (not visible in the top 16 bits)
    low_nan = np.array([0x7F800001], dtype=np.uint32).view(np.float32)[0] if dtype == np.float32 else \
        np.array([0x7FF0000000000001], dtype=np.uint64).view(np.float64)[0]
    a[11, 9] = low_nan
    try:
        gpu_ctx.force_path("cols_bitslice")
        for interp in INTERPS:
            for q in (0.0, 0.5, 0.61, 1.0):
                check(gpu_ctx, oracle, a, 0, [q], interp, skipnan=True)
                assert gpu_ctx.last_path in COLS_PATHS
        with pytest.raises(nds.NanInInput):
            gpu_ctx.quantiles_axis(a, 0, [0.5], "lower")
    finally:
        gpu_ctx.force_path("auto")


# ---- long contiguous lanes: bracket select (one streaming pass), BASELINE configs 4 and 5 in miniature ----
def _long_data(rng, kind, n, dtype):
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        a = {"uniform": lambda: rng.random(n),
             "normal": lambda: rng.standard_normal(n),
             "clustered": lambda: 1000.0 + rng.random(n) * 1e-4,
             "lognormal": lambda: np.exp(3.0 * rng.standard_normal(n)),
             "ties": lambda: rng.integers(-8, 8, size=n).astype(np.float64),
             "equal": lambda: np.full(n, -3.25),
             "sorted": lambda: np.sort(rng.standard_normal(n)),
             "reversed": lambda: np.sort(rng.standard_normal(n))[::-1].copy(),
             "rounded": lambda: np.round(rng.standard_normal(n), 2),
             "extremes": lambda: rng.choice(np.array([np.inf, -np.inf, 0.0, -0.0, 1e300, -1e300, 5e-324]), size=n)}[kind]()
        with np.errstate(over="ignore"):
            return a.astype(dtype)
    info = np.iinfo(dtype)
    return {"uniform": lambda: rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True),
            "normal": lambda: (rng.standard_normal(n) * 1e6).astype(np.int64).astype(dtype),
            "clustered": lambda: (rng.integers(0, 1000, size=n) + 10**6).astype(dtype),
            "lognormal": lambda: np.minimum(np.exp(3.0 * rng.standard_normal(n)) * 100, 2e9).astype(np.int64).astype(dtype),
            "ties": lambda: (rng.integers(0, 16, size=n).astype(np.uint64) * 0x11111111).astype(dtype),
            "equal": lambda: np.full(n, 7, dtype=dtype),
            "sorted": lambda: np.sort(rng.integers(info.min // 2, info.max // 2, size=n, dtype=dtype)),
            "reversed": lambda: np.sort(rng.integers(info.min // 2, info.max // 2, size=n, dtype=dtype))[::-1].copy(),
            "rounded": lambda: (rng.integers(0, 300, size=n) * 1000).astype(dtype),
            "extremes": lambda: rng.choice(np.array([info.min, info.max, 0, 1], dtype=dtype), size=n)}[kind]()


LONG_KINDS = ["uniform", "normal", "clustered", "lognormal", "ties", "equal", "sorted", "reversed", "rounded", "extremes"]


@pytest.mark.parametrize("dtype", SHORT_DTYPES, ids=lambda d: np.dtype(d).name)
def test_bracket_path(gpu_ctx, oracle, dtype):
    """One long lane in one streaming pass: every distribution class, 99 percentiles (C4) and the tie-heavy
    quantiles of C5, lengths around the tile / alignment edges, explicit extremes q = 0 and q = 1."""
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(zlib.crc32(repr(("bracket", dtype.name)).encode()))
    dec = np.arange(1, 10) / 10.0  # 99 percentiles need a lane of tens of millions: test_bracket_c4_shape
    fallbacks = []
    try:
        gpu_ctx.force_path("bracket")
        for n in (65536, 70001, 300007, (1 << 20) + 5):
            for kind in LONG_KINDS:
                a = _long_data(rng, kind, n, dtype)
                for interp, qs in (("linear", dec), ("lower", [0.1, 0.5, 0.9]), ("higher", [0.0, 1.0, 0.5]),
                                   ("midpoint", [0.5]), ("nearest", [0.25, 0.75, 0.999, 0.001])):
                    if dtype.kind != "f" and interp == "linear" and kind in ("extremes", "uniform"):
                        continue  # T::from_f64 overflow panics in the reference: covered by test_quantile1d_and_errors
                    if n > 300007 and interp not in ("linear", "lower"):
                        continue
                    check(gpu_ctx, oracle, a, 0, qs, interp, select_impl=1)
                    assert gpu_ctx.last_path in ("bracket", "bracket->long_radix")
                    if gpu_ctx.last_path != "bracket":
                        fallbacks.append((n, kind, interp))
        # a view that starts off the 16-byte grid, and two long lanes in one call
        a = _long_data(rng, "normal", 200001, dtype)
        check(gpu_ctx, oracle, a[1:], 0, [0.5, 0.9], "lower", select_impl=1)
        b = _long_data(rng, "uniform", 2 * 131072, dtype).reshape(2, 131072)
        check(gpu_ctx, oracle, b, 1, [0.25, 0.5], "higher", select_impl=1)
        assert gpu_ctx.last_path == "bracket"
    finally:
        gpu_ctx.force_path("auto")
    # well-behaved inputs must stay on the one-pass path; the fallback is for 5-sigma events only
    assert not fallbacks, f"bracket select fell back to radix passes for {fallbacks}"


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=lambda d: np.dtype(d).name)
def test_bracket_path_nan(nds, gpu_ctx, oracle, dtype):
    rng = np.random.default_rng(4242)
    try:
        gpu_ctx.force_path("bracket")
        for n in (65536, 250001):
            a = rng.standard_normal(n).astype(dtype)
            a[rng.random(n) < 0.2] = np.nan
            a[5] = -np.nan
            a[7] = np.inf
            a[9] = -np.inf
            for interp in INTERPS:
                check(gpu_ctx, oracle, a, 0, [0.0, 0.01, 0.5, 0.61, 1.0], interp, skipnan=True, select_impl=1)
                assert gpu_ctx.last_path == "bracket"
            with pytest.raises(nds.NanInInput):
                gpu_ctx.quantiles_axis(a, 0, [0.5], "lower")
            check(gpu_ctx, oracle, np.full(n, np.nan, dtype=dtype), 0, [0.5], "linear", skipnan=True)
            mostly = np.full(n, np.nan, dtype=dtype)
            mostly[::1000] = rng.standard_normal(mostly[::1000].size)
            check(gpu_ctx, oracle, mostly, 0, [0.1, 0.5, 0.9], "linear", skipnan=True, select_impl=1)
    finally:
        gpu_ctx.force_path("auto")


def test_bracket_sort1d_and_enqueue(nds, gpu_ctx, oracle):
    """Sort1dExt on a long lane (explicit ranks) and the asynchronous entry point with its deferred fallback."""
    import torch
    rng = np.random.default_rng(99)
    a = rng.standard_normal(400000)
    idx = [0, 1, 399999, 200000, 200001, 12345, 200000]
    got = nds.get_many_from_sorted_mut(a, idx, ctx=gpu_ctx)
    srt = np.sort(a)
    assert list(got.keys()) == sorted(set(idx))
    assert all(got[i] == srt[i] for i in got)
    t = torch.from_numpy(a).cuda()
    dec = np.arange(1, 10) / 10.0
    st, want, _ = oracle.quantiles_axis(a, 0, dec, "linear", select_impl=1)
    out = gpu_ctx.quantiles_axis(t, 0, dec, "linear", enqueue=True)   # (a lane this short: a cluster of CTAs takes it)
    gpu_ctx.synchronize()
    assert ulp_distance(out.cpu().numpy(), want) == 0 and gpu_ctx.last_path in ("mid_bitslice", "bracket")
    try:
        gpu_ctx.force_path("bracket")
        out = gpu_ctx.quantiles_axis(t, 0, dec, "linear", enqueue=True)
        gpu_ctx.synchronize()
    finally:
        gpu_ctx.force_path("auto")
    assert ulp_distance(out.cpu().numpy(), want) == 0
    assert gpu_ctx.last_path == "bracket"
    # 99 percentiles of a lane this short: the brackets would hold most of the lane, automatic dispatch goes
    # straight to the radix passes
    pct = np.arange(1, 100) / 100.0
    st, want, _ = oracle.quantiles_axis(a, 0, pct, "linear", select_impl=1)
    out = gpu_ctx.quantiles_axis(t, 0, pct, "linear", enqueue=True)
    gpu_ctx.synchronize()
    assert gpu_ctx.last_path == "long_radix" and ulp_distance(out.cpu().numpy(), want) == 0
    # ranks that escape their bracket (forced here with absurdly narrow brackets): an enqueued call gives up on the
    # device and is re-run on the radix passes at synchronize(); a synchronous call falls back at once
    import os
    os.environ["NDS_BRK_SIGMA"] = "0.6"
    try:
        gpu_ctx.force_path("bracket")
        out = gpu_ctx.quantiles_axis(t, 0, pct, "linear", enqueue=True)
        out2 = gpu_ctx.quantiles_axis(t, 0, dec, "lower", enqueue=True)
        gpu_ctx.synchronize()
        assert gpu_ctx.last_path == "bracket->long_radix"
        assert ulp_distance(out.cpu().numpy(), want) == 0
        st, want2, _ = oracle.quantiles_axis(a, 0, dec, "lower", select_impl=1)
        assert_bit_equal(out2.cpu().numpy(), want2)
        got = gpu_ctx.quantiles_axis(a, 0, pct, "linear")
        assert gpu_ctx.last_path == "bracket->long_radix" and ulp_distance(got, want) == 0
    finally:
        del os.environ["NDS_BRK_SIGMA"]
        gpu_ctx.force_path("auto")


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=lambda d: np.dtype(d).name)
def test_bracket_c4_shape(nds, gpu_ctx, oracle, dtype):
    """BASELINE config 4 scaled to 2^25 elements: 99 percentiles, Linear, automatic dispatch must pick the
    one-pass path and stay on it; bit-exact against the sort-and-index oracle."""
    rng = np.random.default_rng(1004)
    a = rng.random(1 << 25).astype(dtype)
    qs = np.arange(1, 100) / 100.0
    check(gpu_ctx, oracle, a, 0, qs, "linear", select_impl=1)
    assert gpu_ctx.last_path == "bracket"
    b = (rng.standard_normal(1 << 25) * 3.0).astype(dtype)
    check(gpu_ctx, oracle, b, 0, qs, "linear", select_impl=1)
    assert gpu_ctx.last_path == "bracket"


def test_full_size_c3_against_oracle(nds, gpu_ctx, oracle):
    """BASELINE config 3 at full size (8192 x 65536 f32, Axis(0), skipnan, 10 % NaN, Nearest; the splitmix64 stream of
    SURVEY.md section 8d generated on the device): 128 column lanes spread over the array against the oracle, and -- a
    size-independent property -- every result is an element of its own lane with the right rank among the non-NaNs."""
    import torch
    rows, cols = 8192, 65536
    t = torch.empty((rows, cols), dtype=torch.float32, device="cuda")
    nds.synth.fill(gpu_ctx, t, "f32_uniform_nan10", 1003)
    out = gpu_ctx.quantiles_axis(t, 0, [0.5], "nearest", skipnan=True)
    assert gpu_ctx.last_path.startswith("cols")
    pick = np.linspace(0, cols - 1, num=128).astype(np.int64)
    idx = torch.from_numpy(pick).cuda()
    sub = t.index_select(1, idx).cpu().numpy()
    got = out.index_select(1, idx).cpu().numpy()
    # the device generator and its numpy restatement agree on these lanes
    lane_idx = np.arange(rows, dtype=np.uint64)[:, None] * np.uint64(cols) + pick.astype(np.uint64)[None, :]
    host = nds.synth.host_at("f32_uniform_nan10", lane_idx, 1003)
    assert np.array_equal(sub.view(np.uint32), host.view(np.uint32))
    st, want, _ = oracle.quantiles_axis(sub, 0, [0.5], "nearest", skipnan=True)
    assert st == oracle.OK
    assert_bit_equal(got, want, msg="c3 full size")
    # rank property on ALL 65536 lanes: Nearest selects one order statistic r of the n' non-NaN values
    valid = ~torch.isnan(t)
    n_eff = valid.sum(dim=0)
    x = 0.5 * (n_eff.double() - 1.0)
    r = torch.where((x - torch.trunc(x)) < 0.5, torch.floor(x), torch.ceil(x)).long()
    v = out[0]
    lt = ((t < v[None, :]) & valid).sum(dim=0)
    le = ((t <= v[None, :]) & valid).sum(dim=0)
    assert bool(((lt <= r) & (r < le)).all())
    del t, valid
    torch.cuda.empty_cache()


@pytest.mark.parametrize("dtype", [np.uint32, np.int64, np.float32, np.float64], ids=lambda d: np.dtype(d).name)
def test_bracket_low_cardinality(nds, gpu_ctx, oracle, dtype):
    """Tie-heavy long lanes (BASELINE config 5): at most 63 distinct keys in the sample -> every distinct key is counted in
    a class of its own and nothing is compacted.  All five strategies, ranks all over the lane; a key the sample cannot
    have seen (one occurrence, requested as the maximum) must still come out right (the call falls back)."""
    rng = np.random.default_rng(55)
    n = (1 << 21) + 12345
    for nvals, adversarial in ((16, False), (16, True), (60, False), (2, False), (1, False)):
        if np.dtype(dtype).kind == "f":
            vals = np.sort(rng.standard_normal(nvals)).astype(dtype)
            if nvals >= 2:
                vals[0], vals[1] = dtype(-0.0), dtype(0.0)
            if adversarial:
                vals = (np.float64(1.0) + np.arange(nvals) * np.finfo(dtype).eps).astype(dtype)
        elif adversarial:
            vals = (np.int64(0x80000000 if dtype == np.uint32 else -3) + np.arange(nvals)).astype(dtype)
        else:
            vals = (np.arange(nvals, dtype=np.int64) * (0x11111111 if dtype == np.uint32 else (1 << 40))).astype(dtype)
            if dtype == np.int64:
                vals = vals - np.int64(8 << 40)
        a = vals[rng.integers(0, nvals, n)]
        qs = [0.0, 0.1, 0.25, 0.5, 0.75, 0.9, 0.999, 1.0]
        for interp in INTERPS:
            check(gpu_ctx, oracle, a, 0, qs, interp, select_impl=1)
            assert gpu_ctx.last_path == "bracket", gpu_ctx.last_path
    # a key the sample has not seen
    a = np.asarray([3, 5, 7], dtype=dtype)[rng.integers(0, 3, n)]
    a[n // 3] = 9
    check(gpu_ctx, oracle, a, 0, [0.5, 1.0], "lower", select_impl=1)
    assert gpu_ctx.last_path == "bracket->long_radix", gpu_ctx.last_path
    # ... and the same lane asked for ranks far from the unseen key: it is counted in the gap above 7 (the table misses
    # it), the requested ranks resolve from the counts
    check(gpu_ctx, oracle, a, 0, [0.2, 0.5], "higher", select_impl=1)
    assert gpu_ctx.last_path == "bracket", gpu_ctx.last_path
    if np.dtype(dtype).kind == "f":  # NaNs beside few values, skipnan and not
        b = a.copy()
        b[rng.random(n) < 0.05] = np.nan
        check(gpu_ctx, oracle, b, 0, [0.2, 0.5, 0.9], "midpoint", skipnan=True, select_impl=1)
        with pytest.raises(nds.NanInInput):
            gpu_ctx.quantiles_axis(b, 0, [0.5], "lower")


def test_c5_i64_lane_shape(nds, gpu_ctx, oracle):
    """BASELINE config 5's lane half: 8192-element i64 lanes with 16 distinct values ((z & 15) - 8) * 2^40 (the splitmix64
    stream of SURVEY.md section 8d), Lower and Higher -- the warp-per-lane kernel with its all-candidates-equal check."""
    import torch
    t = torch.empty((512, 8192), dtype=torch.int64, device="cuda")
    nds.synth.fill(gpu_ctx, t, "i64_16values", 1005)
    a = t.cpu().numpy()
    for interp in ("lower", "higher", "midpoint"):
        check(gpu_ctx, oracle, a, 1, [0.1, 0.5, 0.9], interp, select_impl=1)
        assert gpu_ctx.last_path == "short_bitslice", gpu_ctx.last_path


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int16, np.uint8], ids=lambda d: np.dtype(d).name)
def test_mid_length_lanes_dispatch(nds, gpu_ctx, oracle, dtype):
    """Lanes too long for the lane kernels and too few for one CTA each to pay: a handful of lanes, or many quantiles of a few
    lanes, run one after the other over the whole grid (every requested rank per pass); the call must not fall into the
    rank-at-a-time CTA kernel.  The 65536-element x 99-percentile call of the smoke test is one of these."""
    rng = np.random.default_rng(31)
    pct = np.arange(1, 100) / 100.0
    if np.dtype(dtype).kind == "f":
        a = rng.standard_normal((3, 65536)).astype(dtype)
    else:
        info = np.iinfo(dtype)
        a = rng.integers(info.min, info.max, size=(3, 65536), dtype=dtype, endpoint=True)
    grid_wide = ("long_radix", "bracket", "bracket->long_radix")
    # (one-byte keys need a single digit pass per rank: there the cost model may rightly keep the CTA kernel)
    ok_paths = grid_wide if np.dtype(dtype).itemsize >= 2 else grid_wide + ("generic_cta",)
    if np.dtype(dtype).itemsize >= 4:
        ok_paths = ok_paths + ("mid_bitslice",)   # contiguous lanes of up to 98304 elements: the CTA-per-lane bit-sliced kernel
    check(gpu_ctx, oracle, a, 1, pct, "linear" if np.dtype(dtype).kind == "f" else "lower", select_impl=1)
    assert gpu_ctx.last_path in ok_paths, gpu_ctx.last_path
    check(gpu_ctx, oracle, a[0], 0, pct, "nearest", select_impl=1)          # the smoke shape
    assert gpu_ctx.last_path in ok_paths, gpu_ctx.last_path
    b = a[:, :20001].T.copy()                                                 # 20001 x 3: strided lanes along axis 0
    check(gpu_ctx, oracle, b, 0, [0.25, 0.5], "midpoint", select_impl=1)     # (either path may win the cost model)
    check(gpu_ctx, oracle, b, 0, pct, "higher", select_impl=1)               # 99 ranks of strided lanes: grid-wide passes
    if np.dtype(dtype).itemsize >= 4:   # (or the layout stage + the CTA-per-lane kernel: three lanes run side by side)
        assert gpu_ctx.last_path in ("long_radix", "bracket", "bracket->long_radix", "gather->mid_bitslice"), gpu_ctx.last_path
    wide = np.tile(a[:, :9000], (20, 1))                                      # 60 lanes x 9000, 2 quantiles: one CTA per lane
    check(gpu_ctx, oracle, wide, 1, [0.5, 0.9], "higher", select_impl=1)
    assert gpu_ctx.last_path == ("mid_bitslice" if np.dtype(dtype).itemsize >= 4 else "generic_cta"), gpu_ctx.last_path


@pytest.mark.parametrize("dtype", SHORT_DTYPES, ids=lambda d: np.dtype(d).name)
def test_mid_bitslice_path(nds, gpu_ctx, oracle, dtype):
    """Contiguous lanes between the warp-per-lane kernel and the grid-wide passes (8193 ... 98304 elements), and calls with
    too few lanes for warp-per-lane: one CTA per lane, bit-sliced with informative-bit windows (nds_mid.cu).  Lengths up
    to the limit, ragged last blocks, fewer blocks than warps, every data kind of the short-lane tests plus the window
    cases (small magnitudes, late large values, late NaNs, many tie values)."""
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(zlib.crc32(repr(("mid", dtype.name)).encode()))
    vec = 16 // dtype.itemsize
    qs_sets = {"midpoint": [0.5], "linear": [0.25, 0.5, 0.75], "lower": [0.0, 0.1, 0.9, 1.0], "higher": [0.999, 0.5, 0.5],
               "nearest": [0.3, 0.7]}
    try:
        gpu_ctx.force_path("mid_bitslice")
        for n, lanes in ((98304, 3), (98304 - vec, 2), (65536, 5), (20000 // vec * vec, 160), (9000 // vec * vec, 7), (1024, 9), (64, 3), (300 // vec * vec, 4)):
            kinds = ["normal", "uniform", "ties", "equal", "small", "late-large", "100-values"]
            if dtype.kind == "f":
                kinds += ["extremes", "late-nan"]
            for kind in kinds:
                skip = False
                if dtype.kind == "f":
                    if kind == "normal":
                        a = rng.standard_normal((lanes, n))
                    elif kind == "uniform":
                        a = rng.random((lanes, n))
                    elif kind == "ties":
                        a = rng.integers(-8, 8, size=(lanes, n)).astype(np.float64)
                    elif kind == "equal":
                        a = np.full((lanes, n), -3.25)
                    elif kind == "small":
                        a = (rng.random((lanes, n)) - 0.5) * 1e-30
                    elif kind == "late-large":
                        a = 1.0 + rng.random((lanes, n))
                        a[:, n - n // 3:] *= 1e30
                    elif kind == "100-values":
                        a = rng.choice(rng.standard_normal(100), size=(lanes, n))
                    elif kind == "extremes":
                        a = rng.standard_normal((lanes, n))
                        a[:, ::7] = np.inf
                        a[:, 1::13] = -np.inf
                        a[:, 2::17] = 0.0
                        a[:, 3::19] = -0.0
                        a[:, 5::23] = np.finfo(dtype).tiny / 4
                    else:
                        a = rng.random((lanes, n))
                        a[:, n - n // 4:][rng.random((lanes, n // 4)) < 0.2] = np.nan
                        a[0, :] = np.nan
                        skip = True
                    a = a.astype(dtype)
                else:
                    info = np.iinfo(dtype)
                    if kind in ("normal", "uniform"):
                        a = rng.integers(info.min // 2, info.max // 2, size=(lanes, n), dtype=dtype, endpoint=True)
                    elif kind == "ties":
                        a = rng.integers(0, 16, size=(lanes, n)).astype(dtype)
                    elif kind == "equal":
                        a = np.full((lanes, n), 7, dtype=dtype)
                    elif kind == "small":
                        a = rng.integers(-1000 if info.min < 0 else 0, 1000, size=(lanes, n)).astype(dtype)
                    elif kind == "late-large":
                        a = rng.integers(0, 100, size=(lanes, n)).astype(dtype)
                        a[:, n - n // 3:] = rng.integers(info.max // 4, info.max // 2, size=(lanes, n // 3), dtype=dtype)
                    else:
                        a = rng.choice(rng.integers(info.min, info.max, size=100, dtype=dtype), size=(lanes, n))
                for interp, qs in qs_sets.items():
                    if n > 30000 and interp in ("nearest", "higher") and kind not in ("uniform", "ties"):
                        continue  # (keeps the oracle's share of the run time down)
                    check(gpu_ctx, oracle, a, 1, qs, interp, skipnan=skip, select_impl=1)
                    assert gpu_ctx.last_path == "mid_bitslice", gpu_ctx.last_path
        # a 3-d array whose last axis is the lane, and a pitched 2-d view
        a = rng.standard_normal((3, 4, 12000)).astype(dtype) if dtype.kind == "f" else rng.integers(0, 10**6, size=(3, 4, 12000)).astype(dtype)
        check(gpu_ctx, oracle, a, 2, [0.5, 0.9], "lower")
        assert gpu_ctx.last_path == "mid_bitslice"
        check(gpu_ctx, oracle, a[:, :, :10000], 2, [0.5], "midpoint")
        assert gpu_ctx.last_path == "mid_bitslice"
        with pytest.raises(nds.NdsError):
            gpu_ctx.quantiles_axis(a[:, :, ::2], 2, [0.5], "lower")   # strided lanes: not this kernel
    finally:
        gpu_ctx.force_path("auto")
    # lanes shared by a thread-block cluster (2, 4, 8 CTAs): beyond 98304 elements, and forced on shorter lanes
    try:
        gpu_ctx.force_path("mid_bitslice")
        for n, lanes in ((131072, 3), (200000 // vec * vec, 2), (786432, 2), (786432 - vec, 1), (300000 // vec * vec, 40)):
            for kind in ("uniform", "ties", "late-large") + (("late-nan",) if dtype.kind == "f" else ()):
                skip = False
                if kind == "uniform":
                    a = rng.random((lanes, n)) if dtype.kind == "f" else rng.integers(-10**6, 10**6, size=(lanes, n))
                    if dtype.kind == "u":
                        a = np.abs(a)
                elif kind == "ties":
                    a = rng.integers(0, 16, size=(lanes, n))
                elif kind == "late-large":
                    a = rng.integers(0, 100, size=(lanes, n)).astype(np.float64)
                    a[:, n - n // 5:] += 2.0 ** 30
                else:
                    a = rng.random((lanes, n))
                    a[:, n - n // 4:][rng.random((lanes, n // 4)) < 0.2] = np.nan
                    skip = True
                a = a.astype(dtype)
                for interp, qs in (("linear", [0.25, 0.5, 0.75]), ("lower", [0.0, 0.001, 0.5, 1.0])):
                    if lanes == 40 and (interp != "linear" or kind != "uniform"):
                        continue
                    if dtype.kind != "f" and interp == "linear" and kind == "late-large":
                        pass
                    check(gpu_ctx, oracle, a, 1, qs, interp, skipnan=skip, select_impl=1)
                    assert gpu_ctx.last_path == "mid_bitslice", gpu_ctx.last_path
        for force in ("2", "4", "8"):
            os.environ["NDS_MID_CLUSTER"] = force
            for n in (20000 // vec * vec, 9000 // vec * vec, 65536):
                a = (rng.standard_normal((5, n)) * 1000).astype(dtype) if dtype.kind != "u" else rng.integers(0, 10**6, size=(5, n)).astype(dtype)
                check(gpu_ctx, oracle, a, 1, [0.1, 0.5, 0.9, 0.99], "midpoint", select_impl=1)
                b = rng.integers(0, 3, size=(4, n)).astype(dtype)
                check(gpu_ctx, oracle, b, 1, [0.0, 0.4, 1.0], "higher", select_impl=1)
    finally:
        os.environ.pop("NDS_MID_CLUSTER", None)
        gpu_ctx.force_path("auto")
    # default dispatch: a few short lanes (fewer than the warp-per-lane kernel wants) and many mid-length lanes
    a = rng.standard_normal((5, 2048)).astype(dtype) if dtype.kind == "f" else rng.integers(0, 10**6, size=(5, 2048)).astype(dtype)
    check(gpu_ctx, oracle, a, 1, [0.25, 0.5], "higher")
    assert gpu_ctx.last_path == "mid_bitslice", gpu_ctx.last_path
    if dtype.kind == "f":
        b = a.copy()
        b[1, 5] = np.nan
        with pytest.raises(nds.NanInInput):
            gpu_ctx.quantiles_axis(b, 1, [0.5], "lower")


@pytest.mark.parametrize("dtype", SHORT_DTYPES, ids=lambda d: np.dtype(d).name)
def test_gather_layout_stage(nds, gpu_ctx, oracle, dtype):
    """Lanes the bit-sliced kernels cannot read in place (strided beyond the column kernel's reach, negative strides,
    contiguous but not 16-byte aligned, odd lengths) are gathered into aligned, padded scratch (nds_layout.cu) and
    selected by the CTA-per-lane kernel: transpose mode (lanes closer than elements) and direct mode."""
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(zlib.crc32(repr(("gather", dtype.name)).encode()))

    def data(shape):
        if dtype.kind == "f":
            return rng.standard_normal(shape).astype(dtype)
        return rng.integers(-10**6 if dtype.kind == "i" else 0, 10**6, size=shape).astype(dtype)

    cases = [
        (data((10001, 70)), 0),                 # tall matrix, columns: transpose mode, odd lane length
        (data((300000, 5)), 0),                 # very tall, few columns: clusters behind the layout stage
        (data((40, 10001)), 1),                 # contiguous lanes whose rows are not 16-byte aligned: direct mode
        (data((33, 9000))[:, ::3], 1),          # strided elements inside a lane
        (data((50, 4000))[:, ::-1], 1),         # negative stride
        (data((12000, 6, 7)), 0),               # 3-d, lanes along the slowest axis
        (data((5, 12001, 3)), 1),               # 3-d, lanes along the middle axis
        (np.asfortranarray(data((31, 700))), 1),  # Fortran order: lanes along the strided axis
    ]
    try:
        gpu_ctx.force_path("gather")
        for a, axis in cases:
            for interp, qs in (("linear", [0.25, 0.5, 0.75]), ("lower", [0.0, 0.37, 1.0]), ("midpoint", [0.5])):
                check(gpu_ctx, oracle, a, axis, qs, interp, select_impl=1)
                assert gpu_ctx.last_path == "gather->mid_bitslice", gpu_ctx.last_path
        if dtype.kind == "f":
            a = data((9001, 40))
            a[rng.random(a.shape) < 0.1] = np.nan
            a[:, 3] = np.nan
            check(gpu_ctx, oracle, a, 0, [0.1, 0.5], "nearest", skipnan=True, select_impl=1)
            with pytest.raises(nds.NanInInput):
                gpu_ctx.quantiles_axis(a, 0, [0.5], "lower")
    finally:
        gpu_ctx.force_path("auto")
    # default dispatch: the tall matrix goes through the layout stage by itself
    a = data((20000, 48))
    check(gpu_ctx, oracle, a, 0, [0.5, 0.9], "higher", select_impl=1)
    assert gpu_ctx.last_path == "gather->mid_bitslice", gpu_ctx.last_path
