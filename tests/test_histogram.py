"""histogram::{Edges, Bins, Grid, GridBuilder, HistogramExt} over the device binning kernel.
Reference doc-tests: src/histogram/bins.rs:7-32, 199-229, 343-430; grid.rs:17-90, 160-218; histograms.rs:88-129."""
import numpy as np
import pytest


def test_oracle_histogram_reference_doc_example(oracle):
    obs = np.array([[1., 0.5], [-0.5, 1.], [-1., -0.5], [0.5, -1.]])
    e = [-1., 0., 1., 2.]
    assert np.array_equal(oracle.histogram_counts(obs, [e, e]), [[1, 0, 1], [1, 0, 0], [0, 1, 0]])   # histograms.rs:119-127


def test_edges_bins_grid_doc_tests(nds):
    edges = nds.Edges([1, 15, 10, 10, 20])                       # bins.rs:44-56: sorted and de-duplicated
    assert list(edges.as_array_view()) == [1, 10, 15, 20]
    e = nds.Edges([0, 2, 3])                                     # bins.rs:199-216
    assert e.indices_of(1) == (0, 1) and e.indices_of(5) is None and e.indices_of(3) is None and e.indices_of(0) == (0, 1)
    bins = nds.Bins(nds.Edges([0, 2, 4, 6]))                     # bins.rs:343-430
    assert len(bins) == 3 and bins.index_of(1) == 0 and bins.index_of(6) is None and bins.index_of(-1) is None
    assert bins.range_of(1) == (0, 2) and bins.range_of(10) is None and bins.index(1) == (2, 4)
    assert len(nds.Bins(nds.Edges([]))) == 0 and len(nds.Bins(nds.Edges([0]))) == 0
    grid = nds.Grid([nds.Bins(nds.Edges([0, 1])), nds.Bins(nds.Edges([-1, 0, 1]))])   # grid.rs:160-218
    assert grid.ndim() == 2 and grid.shape() == [1, 2]
    assert grid.index_of([0, 0]) == [0, 1] and grid.index_of([0, -1]) == [0, 0] and grid.index_of([1, -1]) is None
    with pytest.raises(ValueError):
        grid.index_of([0, 0, 0])
    assert grid.index([0, 1]) == [(0, 1), (0, 1)]


@pytest.mark.gpu
def test_gpu_histogram_reference_doc_example(nds, gpu_ctx):
    """histograms.rs:88-129: GridBuilder::<Sqrt<N64>>::from_array(...).build(), then observations.histogram(grid)."""
    obs = np.array([[1., 0.5], [-0.5, 1.], [-1., -0.5], [0.5, -1.]])
    grid = nds.GridBuilder.from_array(nds.Sqrt, obs, ctx=gpu_ctx).build()
    e = nds.Bins(nds.Edges([-1., 0., 1., 2.]))
    assert grid == nds.Grid([e, e])
    h = nds.histogram(obs, grid, ctx=gpu_ctx)
    assert h.ndim() == 2 and np.array_equal(h.counts(), [[1, 0, 1], [1, 0, 0], [0, 1, 0]])
    h2 = nds.Histogram(grid)                                     # histograms.rs:28-56: add_observation / BinNotFound
    h2.add_observation([0.5, 0.6])
    assert h2.counts()[grid.index_of([0.5, 0.6])[0], grid.index_of([0.5, 0.6])[1]] == 1
    with pytest.raises(nds.BinNotFound):
        h2.add_observation([100.0, 0.0])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32, np.int64, np.uint8, np.int16], ids=lambda d: np.dtype(d).name)
def test_gpu_histogram_matches_oracle(nds, gpu_ctx, oracle, dtype):
    import torch
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(5 + dtype.itemsize)
    for n_obs, n_dims, nbins in ((1, 1, 3), (1000, 1, 7), (100003, 2, 50), (300000, 3, 40), (200000, 2, 300), (50000, 4, 6)):
        if dtype.kind == "f":
            a = rng.standard_normal((n_obs, n_dims)).astype(dtype)
            edges = [np.sort(rng.standard_normal(nbins + 1)).astype(dtype) * 1.5 for _ in range(n_dims)]
        else:
            info = np.iinfo(dtype)
            lo, hi = max(info.min, -1000), min(info.max, 1000)
            a = rng.integers(lo, hi, size=(n_obs, n_dims), endpoint=True).astype(dtype)
            edges = [np.unique(rng.integers(lo, hi, size=nbins + 1, endpoint=True)).astype(dtype) for _ in range(n_dims)]
        want = oracle.histogram_counts(a, edges)
        grid = nds.Grid([nds.Bins(nds.Edges(e)) for e in edges])
        got = nds.histogram(a, grid, ctx=gpu_ctx).counts()
        assert got.shape == want.shape and np.array_equal(got, want), (dtype, n_obs, n_dims, nbins)
        assert int(got.sum()) <= n_obs
        # strided views (every second observation, reversed coordinates) and a device-resident matrix
        v = a[::2, ::-1]
        got = nds.histogram(v, nds.Grid([nds.Bins(nds.Edges(e)) for e in edges[::-1]]), ctx=gpu_ctx).counts()
        assert np.array_equal(got, oracle.histogram_counts(v, edges[::-1]))
        if dtype in (np.float64, np.float32, np.int32, np.int64) and n_obs >= 1000:
            got = nds.histogram(torch.from_numpy(a).cuda(), grid, ctx=gpu_ctx).counts()
            assert np.array_equal(got, want)
    # an empty matrix, and a grid with an empty projection
    assert nds.histogram(np.zeros((0, 2), dtype=dtype), nds.Grid([[0, 1], [0, 1, 2]]), ctx=gpu_ctx).counts().sum() == 0
    assert nds.histogram(np.zeros((5, 2), dtype=dtype), nds.Grid([[0], [0, 1, 2]]), ctx=gpu_ctx).counts().shape == (0, 2)
    if dtype.kind == "f":   # NaN coordinates fall in no bin
        a = np.array([[0.5, np.nan], [0.5, 0.5]], dtype=dtype)
        assert nds.histogram(a, nds.Grid([[0, 1], [0, 1]]), ctx=gpu_ctx).counts()[0, 0] == 1


@pytest.mark.gpu
def test_gpu_grid_builder_strategies(nds, gpu_ctx, oracle):
    """Every strategy as the GridBuilder's type parameter; the histogram of the data over its own grid holds (nearly)
    every row and equals the restated reference loop."""
    rng = np.random.default_rng(8)
    obs = rng.standard_normal((20000, 3)) * [1.0, 10.0, 0.1]
    for strategy in (nds.Sqrt, nds.Rice, nds.Sturges, nds.FreedmanDiaconis, nds.Auto):
        grid = nds.GridBuilder.from_array(strategy, obs, ctx=gpu_ctx).build()
        assert grid.ndim() == 3
        h = nds.histogram(obs, grid, ctx=gpu_ctx)
        # (the last edge is min + n * width, the bin count comes from repeated addition: rounding may leave a column's
        # maximum on or beyond the last edge, as in the reference)
        assert obs.shape[0] - 3 <= int(h.counts().sum()) <= obs.shape[0]
        assert np.array_equal(h.counts(), oracle.histogram_counts(obs, [p.edges.as_array_view() for p in grid.projections()]))
    with pytest.raises(nds.BinsBuildError):
        nds.GridBuilder.from_array(nds.Sqrt, np.ones((10, 2)), ctx=gpu_ctx)       # a constant column: Strategy error
