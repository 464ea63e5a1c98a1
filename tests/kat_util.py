"""Helpers shared by the CPU (oracle) and GPU (product) parity tests."""
import json
import math
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_kats():
    with open(os.path.join(GOLDEN, "reference_kats.json")) as f:
        return json.load(f)


def kat_array(case, none_as="nan"):
    """Build the numpy input of a quantile KAT.

    Option<i32> inputs (null entries) become float64 with NaN when none_as == "nan", or a pair
    (int values, validity mask) when none_as == "mask".
    """
    dtype = np.dtype(case["dtype"])
    if "shape" in case:
        return np.zeros(case["shape"], dtype=dtype)
    raw = case["data"]

    def has_null(x):
        if isinstance(x, list):
            return any(has_null(v) for v in x)
        return x is None

    def conv(x, fill):
        if isinstance(x, list):
            return [conv(v, fill) for v in x]
        if x is None:
            return fill
        if x == "nan":
            return math.nan
        return x

    if has_null(raw):
        if none_as == "nan":
            return np.array(conv(raw, math.nan), dtype=np.float64)
        vals = np.array(conv(raw, 0), dtype=dtype)
        mask = np.array(conv(_map(raw, lambda v: v is not None), False), dtype=bool)
        return vals, mask
    return np.array(conv(raw, None), dtype=dtype)


def _map(x, f):
    if isinstance(x, list):
        return [_map(v, f) for v in x]
    return f(x)


def kat_expect(case):
    exp = case["expect"]

    def conv(x):
        if isinstance(x, list):
            return [conv(v) for v in x]
        if x is None or x == "nan":
            return math.nan
        return x

    return np.array(conv(exp), dtype=np.float64)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({4: np.uint32, 8: np.uint64, 2: np.uint16, 1: np.uint8}[a.dtype.itemsize])


def fold_zero(a):
    """Map -0.0 to +0.0 (the reference's choice between them depends on its random pivot, SURVEY note Z)."""
    a = np.array(a, copy=True)
    if a.dtype.kind == "f":
        a[a == 0] = 0.0
    return a


def assert_bit_equal(got, want, fold=True, msg=""):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, f"{msg} shape {got.shape} != {want.shape}"
    assert got.dtype == want.dtype, f"{msg} dtype {got.dtype} != {want.dtype}"
    if fold:
        got, want = fold_zero(got), fold_zero(want)
    if got.dtype.kind == "f":
        # NaN results of IEEE arithmetic (inf - inf inside Midpoint/Linear) carry a platform-defined sign and
        # payload (x86 SSE: 0xFFC00000, CUDA: 0x7FFFFFFF, and the Rust reference inherits its host's); only
        # their positions are compared.  The canonical NaN of all-NaN skipnan lanes (f32::NAN) IS specified
        # and is checked bit-for-bit by the tests that produce it.
        gn, wn = np.isnan(got), np.isnan(want)
        assert np.array_equal(gn, wn), f"{msg} NaN positions differ"
        gb, wb = bits(np.where(gn, 0, got).astype(got.dtype)), bits(np.where(wn, 0, want).astype(want.dtype))
        bad = np.nonzero(gb != wb)
        if bad[0].size:
            i = tuple(b[0] for b in bad)
            raise AssertionError(f"{msg} {bad[0].size} mismatches; first at {i}: got {got[i]!r} ({gb[i]:#x}) want {want[i]!r} ({wb[i]:#x})")
    else:
        bad = np.nonzero(got != want)
        if bad[0].size:
            i = tuple(b[0] for b in bad)
            raise AssertionError(f"{msg} {bad[0].size} mismatches; first at {i}: got {got[i]!r} want {want[i]!r}")


def ulp_distance(got, want):
    """Max distance in units in the last place between two float arrays (NaN == NaN, -0 == +0)."""
    got = fold_zero(np.asarray(got))
    want = fold_zero(np.asarray(want))
    assert got.shape == want.shape and got.dtype == want.dtype
    if got.size == 0:
        return 0
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert (nan_g == nan_w).all(), "NaN pattern differs"
    it = {4: np.int32, 8: np.int64}[got.dtype.itemsize]
    g = np.ascontiguousarray(got).view(it).astype(np.int64 if it == np.int32 else object)
    w = np.ascontiguousarray(want).view(it).astype(np.int64 if it == np.int32 else object)
    # map sign-magnitude to a monotone integer line
    minint = -(1 << (8 * got.dtype.itemsize - 1))
    g = np.where(g < 0, minint - g, g)
    w = np.where(w < 0, minint - w, w)
    d = np.abs(g - w)
    d = np.where(nan_g, 0, d)
    return int(d.max())


# --- an independent numpy statement of the semantics (third opinion next to the two C selectors) ---

def np_reference(a, axis, qs, interp, skipnan=False):
    """Sort-based, pure numpy/python evaluation of the reference formulas, lane by lane.  Small inputs only."""
    a = np.asarray(a)
    qs = [float(q) for q in np.asarray(qs, dtype=np.float64).reshape(-1)]
    moved = np.moveaxis(a, axis, -1)
    lanes_shape = moved.shape[:-1]
    out = np.zeros(lanes_shape + (len(qs),), dtype=a.dtype)
    for idx in np.ndindex(*lanes_shape):
        lane = np.array(moved[idx])
        if skipnan and lane.dtype.kind == "f":
            lane = lane[~np.isnan(lane)]
        n = lane.size
        if n == 0:
            out[idx] = np.nan
            continue
        s = np.sort(lane, kind="stable")
        for t, q in enumerate(qs):
            x = q * float(n - 1)
            lo, hi = int(math.floor(x)), int(math.ceil(x))
            frac = x - math.trunc(x)
            out[idx + (t,)] = _interp(s[lo], s[hi], frac, interp, a.dtype)
    return np.moveaxis(out, -1, axis)


def _wrap(v, dtype):
    nbits = dtype.itemsize * 8
    v &= (1 << nbits) - 1
    if dtype.kind == "i" and v >= 1 << (nbits - 1):
        v -= 1 << nbits
    return v


def _interp(lo, hi, frac, interp, dtype):
    with np.errstate(all="ignore"):
        if interp == "lower":
            return lo
        if interp == "higher":
            return hi
        if interp == "nearest":
            return lo if frac < 0.5 else hi
        if interp == "midpoint":
            if dtype.kind == "f":
                return dtype.type(lo + dtype.type(dtype.type(hi - lo) / dtype.type(2)))
            d = _wrap(int(hi) - int(lo), dtype)          # release-mode wrapping subtraction
            half = (abs(d) // 2) * (1 if d >= 0 else -1)  # integer division truncates toward zero
            return dtype.type(_wrap(int(lo) + half, dtype))
        if interp == "linear":
            prod = frac * (float(hi) - float(lo))
            if dtype.kind == "f":
                return dtype.type(lo + dtype.type(prod))
            return dtype.type(_wrap(int(lo) + int(math.trunc(prod)), dtype))
    raise ValueError(interp)
