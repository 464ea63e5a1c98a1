"""FreedmanDiaconis::from_array, the in-tree caller of quantile_mut."""
import numpy as np
import pytest


def test_oracle_fd_reference_tests(oracle):
    """src/histogram/strategies.rs:593-619 (fd_tests)."""
    assert oracle.freedman_diaconis(np.array([1, 1, 1, 1, 1, 1, 1]))[0] == "Strategy"          # constant_array_are_bad
    assert oracle.freedman_diaconis(np.array([-20] + [1] * 12 + [20]))[0] == "Strategy"        # zero_iqr_is_bad
    assert oracle.freedman_diaconis(np.array([], dtype=np.uint64))[0] == "EmptyInput"          # empty_arrays_are_bad
    st, bw, lo, hi, nb = oracle.freedman_diaconis(np.arange(100, dtype=np.int64))
    assert st == "ok" and (bw, lo, hi) == (2 * (74 - 25) // 4, 0, 99) and nb == 99 // bw + 1


@pytest.mark.gpu
def test_gpu_fd_reference_tests(nds, gpu_ctx):
    for bad in (np.array([1, 1, 1, 1, 1, 1, 1]), np.array([-20] + [1] * 12 + [20])):
        with pytest.raises(nds.BinsBuildError) as ei:
            nds.FreedmanDiaconis.from_array(bad, ctx=gpu_ctx)
        assert ei.value.is_strategy()
    with pytest.raises(nds.BinsBuildError) as ei:
        nds.FreedmanDiaconis.from_array(np.array([], dtype=np.uint64), ctx=gpu_ctx)
    assert ei.value.is_empty_input()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.int32, np.int64, np.uint32, np.float32, np.float64], ids=lambda d: np.dtype(d).name)
def test_gpu_fd_matches_oracle(nds, gpu_ctx, oracle, dtype):
    import torch
    rng = np.random.default_rng(17)
    for n in (15, 1000, 300001, 1 << 22):
        a = (rng.standard_normal(n) * 1000).astype(dtype) if np.dtype(dtype).kind != "u" else \
            rng.integers(0, 100000, size=n).astype(dtype)
        want = oracle.freedman_diaconis(a)
        assert want[0] == "ok"
        fd = nds.FreedmanDiaconis.from_array(a, ctx=gpu_ctx)
        assert (fd.bin_width(), fd.builder.min, fd.builder.max, fd.n_bins()) == want[1:]
        edges = fd.build()
        assert edges.size == want[4] + 1 and edges[0] == want[2] and edges[-1] > want[3]
        if np.dtype(dtype) in (np.int32, np.int64, np.float32, np.float64) and n >= 1000:
            fd2 = nds.FreedmanDiaconis.from_array(torch.from_numpy(a).cuda(), ctx=gpu_ctx)   # device-resident input
            assert fd2.bin_width() == want[1] and fd2.n_bins() == want[4]


def test_oracle_count_rules_reference_tests(oracle):
    """src/histogram/strategies.rs:527-590, 622-645 (sqrt_tests, rice_tests, sturges_tests, auto_tests)."""
    const = np.array([1, 1, 1, 1, 1, 1, 1])
    for rule in ("sqrt", "rice", "sturges"):
        assert oracle.count_rule(const, rule)[0] == "Strategy"                                   # constant_array_are_bad
        assert oracle.count_rule(np.array([], dtype=np.uint64), rule)[0] == "EmptyInput"          # empty_arrays_are_bad
    assert oracle.auto_rule(const)[0] == "Strategy"
    assert oracle.auto_rule(np.array([], dtype=np.uint64))[0] == "EmptyInput"
    assert oracle.auto_rule(np.array([-20] + [1] * 12 + [20]))[0] == "sturges"                    # zero_iqr_is_handled_by_sturged
    st, bw, lo, hi, nb = oracle.count_rule(np.arange(100, dtype=np.int64), "sqrt")
    assert (st, bw, lo, hi, nb) == ("ok", 9, 0, 99, 12)


@pytest.mark.gpu
def test_gpu_count_rules_reference_tests(nds, gpu_ctx):
    const = np.array([1, 1, 1, 1, 1, 1, 1])
    for cls in (nds.Sqrt, nds.Rice, nds.Sturges, nds.Auto):
        with pytest.raises(nds.BinsBuildError) as ei:
            cls.from_array(const, ctx=gpu_ctx)
        assert ei.value.is_strategy()
        with pytest.raises(nds.BinsBuildError) as ei:
            cls.from_array(np.array([], dtype=np.uint64), ctx=gpu_ctx)
        assert ei.value.is_empty_input()
    auto = nds.Auto.from_array(np.array([-20] + [1] * 12 + [20]), ctx=gpu_ctx)    # zero IQR: Sturges takes over
    assert isinstance(auto.builder, nds.Sturges)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.int32, np.int64, np.uint32, np.float32, np.float64], ids=lambda d: np.dtype(d).name)
def test_gpu_count_rules_match_oracle(nds, gpu_ctx, oracle, dtype):
    rng = np.random.default_rng(23)
    for n in (2, 15, 1000, 300001):
        a = (rng.standard_normal(n) * 1000).astype(dtype) if np.dtype(dtype).kind != "u" else \
            rng.integers(0, 100000, size=n).astype(dtype)
        if a.min() == a.max():
            continue
        for cls, rule in ((nds.Sqrt, "sqrt"), (nds.Rice, "rice"), (nds.Sturges, "sturges")):
            want = oracle.count_rule(a, rule)
            if want[0] != "ok":
                with pytest.raises(nds.BinsBuildError):
                    cls.from_array(a, ctx=gpu_ctx)
                continue
            got = cls.from_array(a, ctx=gpu_ctx)
            assert (got.bin_width(), got.builder.min, got.builder.max, got.n_bins()) == want[1:]
        want = oracle.auto_rule(a)
        if want[0] in ("fd", "sturges"):
            got = nds.Auto.from_array(a, ctx=gpu_ctx)
            assert isinstance(got.builder, nds.FreedmanDiaconis if want[0] == "fd" else nds.Sturges)
            assert (got.bin_width(), got.n_bins()) == (want[1], want[4])
