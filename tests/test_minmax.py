"""QuantileExt::{min, max, argmin, argmax}(+_skipnan) — SURVEY.md §8(f) row N2.
CPU: the oracle restatement against the reference's own known-answer tests (tests/quantile.rs:12-169 and the trait's
doc-tests).  GPU: the CUDA reduction through the C ABI against the oracle."""
import numpy as np
import pytest

from kat_util import load_kats

KATS = load_kats()["minmax"]
FN = {"argmin": (False, False), "argmax": (True, False), "min": (False, False), "max": (True, False),
      "argmin_skipnan": (False, True), "argmax_skipnan": (True, True), "min_skipnan": (False, True),
      "max_skipnan": (True, True)}


def _arr(case):
    data = case["data"]
    a = np.array([[np.nan if x == "nan" else x for x in row] for row in data], dtype=case["dtype"])
    return a.reshape(len(data), -1)


@pytest.mark.parametrize("case", KATS, ids=lambda c: c["id"])
def test_oracle_minmax_kats(oracle, case):
    want_max, skipnan = FN[case["fn"]]
    st, v, idx = oracle.minmax(_arr(case), want_max, skipnan)
    if "error" in case:
        assert st == case["error"]
    elif "expect_index" in case:
        assert st == "ok" and list(idx) == case["expect_index"]
    elif case["expect_value"] == "nan":
        assert st == "EmptyInput" and np.isnan(v)   # min/max_skipnan return NaN (tests/quantile.rs:84-88)
    else:
        assert st == "ok" and v == case["expect_value"]


def test_oracle_argmin_matches_min_property(oracle):
    """tests/quantile.rs:27-31, 105-109 (quickcheck): a[argmin] == min, a[argmax] == max."""
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = rng.standard_normal(int(rng.integers(1, 40))).astype(np.float32)
        for want_max in (False, True):
            st, v, idx = oracle.minmax(a, want_max)
            assert st == "ok" and a[idx] == v == (a.max() if want_max else a.min())


@pytest.mark.gpu
@pytest.mark.parametrize("case", KATS, ids=lambda c: c["id"])
def test_gpu_minmax_kats(nds, gpu_ctx, case):
    a = _arr(case)
    fn = getattr(nds, case["fn"])
    if case.get("error") == "EmptyInput":
        with pytest.raises(nds.EmptyInput):   # MinMaxError::EmptyInput is an EmptyInput too
            fn(a, ctx=gpu_ctx)
        if not case["fn"].endswith("skipnan"):
            with pytest.raises(nds.MinMaxError):
                fn(a, ctx=gpu_ctx)
    elif case.get("error") == "UndefinedOrder":
        with pytest.raises(nds.UndefinedOrder) as ei:
            fn(a, ctx=gpu_ctx)
        assert str(ei.value) == "Undefined ordering between a tested pair of values."
    elif "expect_index" in case:
        assert list(fn(a, ctx=gpu_ctx)) == case["expect_index"]
    elif case["expect_value"] == "nan":
        assert np.isnan(fn(a, ctx=gpu_ctx))
    else:
        assert fn(a, ctx=gpu_ctx) == case["expect_value"]


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64,
                                   np.float32, np.float64], ids=lambda d: np.dtype(d).name)
def test_gpu_minmax_matrix(nds, gpu_ctx, oracle, dtype):
    """Every dtype, contiguous / strided / reversed / 3-d layouts, ties (first in logical order wins), signed zeros,
    infinities, NaN positions; host arrays and device tensors."""
    import torch
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(np.dtype(dtype).num)
    for shape in ((1,), (7,), (1000003,), (37, 129), (5, 6, 7), (2, 3, 4, 5)):
        if dtype.kind == "f":
            a = rng.standard_normal(shape).astype(dtype)
            flat = a.reshape(-1)
            flat[:: max(1, flat.size // 7)] = flat.min()     # repeated minima: the first must win
            flat[1:: max(1, flat.size // 5)] = flat.max()
        else:
            info = np.iinfo(dtype)
            a = rng.integers(max(info.min, -50), min(info.max, 50), size=shape).astype(dtype)   # heavy ties
        views = [a, a[..., ::-1], a[..., ::2]]
        if a.ndim >= 2:
            views += [np.swapaxes(a, 0, -1), a[::-1]]
        for v in views:
            for want_max in (False, True):
                st, val, idx = oracle.minmax(v, want_max, False)
                assert st == "ok"
                got_idx = (nds.argmax if want_max else nds.argmin)(v, ctx=gpu_ctx)
                got_val = (nds.max if want_max else nds.min)(v, ctx=gpu_ctx)
                assert (got_idx if isinstance(got_idx, tuple) else (got_idx,)) == idx
                assert got_val == val and np.array([got_val]).tobytes() == np.array([val], dtype=dtype).tobytes()
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda() if dtype not in (np.uint16, np.uint32, np.uint64) else None
        if t is not None:
            st, val, idx = oracle.minmax(a, True, False)
            got = nds.argmax(t, ctx=gpu_ctx)
            assert (got if isinstance(got, tuple) else (got,)) == idx
    if dtype.kind == "f":
        z = np.array([3.0, -0.0, 5.0, 0.0, -0.0], dtype=dtype)           # -0.0 == +0.0: the first zero wins
        assert nds.argmin(z, ctx=gpu_ctx) == 1 and np.signbit(nds.min(z, ctx=gpu_ctx))
        z2 = z[::-1].copy() * -1                                          # [0, -0, -5, 0, -3]
        assert nds.argmax(z2, ctx=gpu_ctx) == 0 and not np.signbit(nds.max(z2, ctx=gpu_ctx))
        b = rng.standard_normal((300, 70)).astype(dtype)
        b[rng.random(b.shape) < 0.3] = np.nan
        b[0, 0] = np.nan
        b[5, 5] = np.inf
        b[6, 6] = -np.inf
        for want_max in (False, True):
            st, val, idx = oracle.minmax(b, want_max, True)
            assert (nds.argmax_skipnan if want_max else nds.argmin_skipnan)(b, ctx=gpu_ctx) == idx
            assert (nds.max_skipnan if want_max else nds.min_skipnan)(b, ctx=gpu_ctx) == val
            with pytest.raises(nds.UndefinedOrder):
                (nds.max if want_max else nds.min)(b, ctx=gpu_ctx)


def test_oracle_max_skipnan_returns_the_last_of_equal_maxima(oracle):
    """src/quantile/mod.rs:399-409: `acc.max(elem)`; Ord::max returns its second operand on a tie."""
    a = np.array([0.0, -1.0, np.nan, -0.0], dtype=np.float64)
    st, v, idx = oracle.minmax(a, True, True)
    assert st == "ok" and v == 0.0 and np.signbit(v) and idx == (0,)
    st, v, idx = oracle.minmax(a[::-1].copy(), True, True)
    assert st == "ok" and not np.signbit(v)
    st, v, idx = oracle.minmax(np.array([-0.0, 0.0]), False, True)    # min_skipnan: Ord::min keeps the first
    assert st == "ok" and np.signbit(v)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gpu_max_skipnan_tie_is_the_last(nds, gpu_ctx, oracle, dtype):
    rng = np.random.default_rng(3)
    a = -np.abs(rng.standard_normal(100000)).astype(dtype)
    a[[5, 70000]] = 0.0
    a[[900, 99999]] = -0.0
    a[::97] = np.nan
    a[[5, 900, 70000, 99999]] = [0.0, -0.0, 0.0, -0.0]
    v = nds.max_skipnan(a, ctx=gpu_ctx)
    assert v == 0.0 and np.signbit(v)                      # the last zero is -0.0
    assert nds.argmax_skipnan(a, ctx=gpu_ctx) == 5         # the first one
    st, want, _ = oracle.minmax(a, True, True)
    assert np.asarray(v).tobytes() == np.asarray(want).tobytes()
    b = a.reshape(100, 1000)[:, ::-1]                       # a strided view: logical order is what counts
    v = nds.max_skipnan(b, ctx=gpu_ctx)
    st, want, _ = oracle.minmax(b, True, True)
    assert np.asarray(v).tobytes() == np.asarray(want).tobytes()
