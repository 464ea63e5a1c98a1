"""Sort1dExt::partition_mut (src/sort.rs:133-168) -- SURVEY.md §8(f) row N1.
CPU: the closed form the device kernels implement (nds_partition.cu: index = count of smaller elements, Hoare's swaps =
the k-th misplaced element from the left with the k-th from the right) against the oracle's faithful sequential
restatement, index AND arrangement.  GPU: `nds_partition` through the C ABI against the same oracle, the reference's
known-answer arrays (tests/sort.rs:5-33) and doc-test (src/sort.rs:66-88), every dtype, strided views, device tensors."""
import numpy as np
import pytest

from kat_util import load_kats

KATS = load_kats()


def closed_form(a, pivot_index):
    """numpy statement of what nds_partition.cu computes."""
    a = a.copy()
    n = len(a)
    v = a[pivot_index]
    a[pivot_index], a[0] = a[0], a[pivot_index]
    less = a[1:] < v
    m = int(less.sum())
    pos = np.arange(1, n)
    left = pos[(~less) & (pos <= m)]
    right = pos[less & (pos > m)][::-1]
    assert len(left) == len(right)
    tmp = a[left].copy()
    a[left] = a[right]
    a[right] = tmp
    a[0], a[m] = a[m], a[0]
    return m, a


def check_post_conditions(a, new_index, pivot_value):
    """tests/sort.rs:24-31"""
    assert (a[:new_index] < pivot_value).all()
    assert a[new_index] == pivot_value
    assert (pivot_value <= a[new_index + 1:]).all()


def kat_cases():
    arrays = [np.array(x, dtype=np.int64) for x in KATS["partition"]["arrays"]]
    cases = [(a, len(a) - 1) for a in arrays]
    cases.append((np.array(KATS["doc_partition"]["data"], dtype=np.int64), KATS["doc_partition"]["pivot_index"]))
    return cases


def test_closed_form_is_the_hoare_walk(oracle):
    rng = np.random.default_rng(11)
    cases = kat_cases()
    for _ in range(3000):
        n = int(rng.integers(2, 60))
        a = rng.integers(0, int(rng.integers(1, 15)), n).astype(np.int64)
        cases.append((a, int(rng.integers(0, n))))
    for a, piv in cases:
        want = a.copy()
        want_index = oracle.partition_mut_i64(want, piv)
        got_index, got = closed_form(a, piv)
        assert got_index == want_index and np.array_equal(got, want)
        check_post_conditions(got, got_index, a[piv])


def test_symbol_exported(nds):
    assert hasattr(nds._cabi.load(), "nds_partition")


# ---- GPU ---------------------------------------------------------------------------------------------------
def _oracle_partition(oracle, a, piv):
    """the oracle works on i64: order-preserving for every value the tests use"""
    w = a.astype(np.int64)
    idx = oracle.partition_mut_i64(w, piv)
    return idx, w


@pytest.mark.gpu
def test_reference_kats(nds, gpu_ctx, oracle):
    for a, piv in kat_cases():
        pivot_value = a[piv]
        want_index, want = _oracle_partition(oracle, a, piv)
        got = a.copy()
        new_index = nds.partition_mut(got, piv, ctx=gpu_ctx)
        check_post_conditions(got, new_index, pivot_value)
        assert new_index == want_index and np.array_equal(got, want)
    assert gpu_ctx.last_path == "partition"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["int8", "int16", "int32", "int64", "uint8", "uint16", "uint32", "uint64", "float32", "float64"])
@pytest.mark.parametrize("n", [2, 3, 255, 2049, 70001, 1 << 20])
def test_matches_oracle(nds, gpu_ctx, oracle, dtype, n):
    rng = np.random.default_rng(n * 31 + len(dtype))
    hi = min(n // 3 + 2, {1: 100, 2: 30000}.get(np.dtype(dtype).itemsize, 1 << 30))
    lo = -hi if dtype.startswith(("int", "float")) else 0
    for piv in sorted({0, n // 2, n - 1, int(rng.integers(0, n))}):
        a = rng.integers(lo, hi, n).astype(dtype)      # integer-valued floats: exact in the i64 oracle
        want_index, want = _oracle_partition(oracle, a, piv)
        got = a.copy()
        new_index = nds.partition_mut(got, piv, ctx=gpu_ctx)
        assert new_index == want_index
        assert np.array_equal(got.astype(np.int64), want)


@pytest.mark.gpu
def test_shapes_of_input(nds, gpu_ctx, oracle):
    n = 50000
    for name, a in {"equal": np.full(n, 7, dtype=np.int32), "sorted": np.arange(n, dtype=np.int32),
                    "reversed": np.arange(n, dtype=np.int32)[::-1].copy(),
                    "two values": (np.arange(n) % 2).astype(np.int32)}.items():
        for piv in (0, n // 2, n - 1):
            want_index, want = _oracle_partition(oracle, a, piv)
            got = a.copy()
            assert nds.partition_mut(got, piv, ctx=gpu_ctx) == want_index, name
            assert np.array_equal(got.astype(np.int64), want), name


@pytest.mark.gpu
def test_strided_views(nds, gpu_ctx, oracle):
    rng = np.random.default_rng(5)
    base = rng.integers(-1000, 1000, 30000).astype(np.int64)
    for view in (lambda x: x[::3], lambda x: x[::-1], lambda x: x[5:20000:7], lambda x: x[::-2]):
        buf = base.copy()
        v = view(buf)
        piv = len(v) // 3
        want_index, want = _oracle_partition(oracle, np.ascontiguousarray(v), piv)
        untouched = buf.copy()
        new_index = nds.partition_mut(v, piv, ctx=gpu_ctx)
        assert new_index == want_index and np.array_equal(v, want)
        # elements outside the view are left alone
        mask = np.ones(len(buf), dtype=bool)
        mask[view(np.arange(len(buf)))] = False
        assert np.array_equal(buf[mask], untouched[mask])


@pytest.mark.gpu
def test_floats_follow_ieee_order(nds, gpu_ctx):
    """N64 compares -0.0 == +0.0 (noisy_float wraps the IEEE comparison): neither is "smaller" than the other."""
    rng = np.random.default_rng(2)
    a = rng.standard_normal(10001)
    a[::7] = 0.0
    a[3::7] = -0.0
    piv = 7      # a +0.0
    got = a.copy()
    new_index = nds.partition_mut(got, piv, ctx=gpu_ctx)
    check_post_conditions(got, new_index, 0.0)
    assert new_index == int((a < 0.0).sum())
    assert np.array_equal(np.sort(got), np.sort(a))
    want_index, want = closed_form(a, piv)
    assert new_index == want_index and np.array_equal(got.view(np.uint64), want.view(np.uint64))


@pytest.mark.gpu
def test_device_tensor_in_place(nds, gpu_ctx, oracle):
    import torch
    rng = np.random.default_rng(9)
    a = rng.integers(-50, 50, 123457).astype(np.int32)
    t = torch.from_numpy(a).cuda()
    want_index, want = _oracle_partition(oracle, a, 1000)
    assert nds.partition_mut(t, 1000, ctx=gpu_ctx) == want_index
    assert np.array_equal(t.cpu().numpy().astype(np.int64), want)


@pytest.mark.gpu
def test_edge_cases(nds, gpu_ctx):
    a = np.array([5], dtype=np.int64)
    assert nds.partition_mut(a, 0, ctx=gpu_ctx) == 0 and a[0] == 5
    with pytest.raises(IndexError):          # "Panics if pivot_index is greater than or equal to n", src/sort.rs:77-78
        nds.partition_mut(np.array([1, 2, 3]), 3, ctx=gpu_ctx)
    with pytest.raises(IndexError):
        nds.partition_mut(np.array([], dtype=np.int64), 0, ctx=gpu_ctx)
    b = np.array([2, 1], dtype=np.uint8)
    assert nds.partition_mut(b, 0, ctx=gpu_ctx) == 1 and list(b) == [1, 2]
