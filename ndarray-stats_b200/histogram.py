"""histogram::{Edges, Bins, Grid, GridBuilder, Histogram, HistogramExt} (src/histogram/{bins,grid,histograms}.rs) — the
consumer of the bin-building strategies.  The small container types are host objects; the
binning of the observations, the only part that touches the data, runs on the device (``nds_histogram``).
"""
import numpy as np


class BinNotFound(Exception):
    """src/histogram/errors.rs:6-18"""

    def __str__(self):
        return "No bin has been found."


class Edges:
    """A sorted, de-duplicated collection of bin edges (bins.rs:33-235)."""

    def __init__(self, edges):
        a = np.asarray(edges)
        self._e = np.unique(a.reshape(-1))          # From<Vec<A>>: sort_unstable + dedup (bins.rs:37-73)

    def __len__(self):
        return int(self._e.size)

    def is_empty(self):
        return self._e.size == 0

    def __getitem__(self, i):
        return self._e[i]

    def as_array_view(self):
        return self._e

    def __iter__(self):
        return iter(self._e)

    def __eq__(self, other):
        return isinstance(other, Edges) and np.array_equal(self._e, other._e)

    def indices_of(self, value):
        """(left, right) edge indexes of the bin that holds ``value``, or None (bins.rs:217-229)."""
        n = self._e.size
        i = int(np.searchsorted(self._e, value, side="left"))
        if i < n and self._e[i] == value:           # Ok(i)
            return None if i == n - 1 else (i, i + 1)
        if i == 0 or i == n:                        # Err(0) / Err(n)
            return None
        return (i - 1, i)


class Bins:
    """Left-closed, right-open intervals between consecutive edges (bins.rs:261-430)."""

    def __init__(self, edges):
        self.edges = edges if isinstance(edges, Edges) else Edges(edges)

    def __len__(self):
        return max(len(self.edges) - 1, 0)          # bins.rs:291-299

    def is_empty(self):
        return len(self) == 0

    def index_of(self, value):
        t = self.edges.indices_of(value)
        return None if t is None else t[0]

    def range_of(self, value):
        t = self.edges.indices_of(value)
        return None if t is None else (self.edges[t[0]], self.edges[t[1]])

    def index(self, i):
        if not 0 <= i < len(self):
            raise IndexError("bin index out of bounds")   # the reference panics (bins.rs:417-430)
        return (self.edges[i], self.edges[i + 1])

    def __eq__(self, other):
        return isinstance(other, Bins) and self.edges == other.edges


class Grid:
    """An n-dimensional grid of bins, one ``Bins`` per axis (grid.rs:91-300)."""

    def __init__(self, projections):
        self._p = [p if isinstance(p, Bins) else Bins(p) for p in projections]

    def ndim(self):
        return len(self._p)

    def shape(self):
        return [len(p) for p in self._p]

    def projections(self):
        return self._p

    def index_of(self, point):
        point = np.asarray(point).reshape(-1)
        if point.size != self.ndim():
            raise ValueError(f"Dimension mismatch: the point has {point.size} dimensions, the grid expected "
                             f"{self.ndim()} dimensions.")              # grid.rs:204-211 (assert_eq!)
        idx = [p.index_of(v) for v, p in zip(point, self._p)]
        return None if any(i is None for i in idx) else idx

    def index(self, index):
        if len(index) != self.ndim():
            raise ValueError("Dimension mismatch")
        return [p.index(i) for i, p in zip(index, self._p)]

    def __eq__(self, other):
        return isinstance(other, Grid) and self._p == other._p


class GridBuilder:
    """``GridBuilder<B>``: one bin-building strategy per column of the observation matrix (grid.rs:312-365)."""

    def __init__(self, builders):
        self.bin_builders = builders

    @classmethod
    def from_array(cls, strategy, array, ctx=None):
        """``strategy``: one of ``Sqrt, Rice, Sturges, FreedmanDiaconis, Auto`` (the reference's type parameter B).
        Raises ``BinsBuildError`` as the strategy does (grid.rs:337-344)."""
        if len(array.shape) != 2:
            raise ValueError("GridBuilder::from_array is defined on 2-D arrays")
        return cls([strategy.from_array(array[:, d], ctx=ctx) for d in range(array.shape[1])])

    def build(self):
        return Grid([Bins(Edges(b.build())) for b in self.bin_builders])   # grid.rs:356-359


class Histogram:
    """Counts over a grid (histograms.rs:6-72)."""

    def __init__(self, grid, counts=None):
        self._grid = grid
        self._counts = np.zeros(tuple(grid.shape()), dtype=np.uint64) if counts is None else counts

    def add_observation(self, observation):
        idx = self._grid.index_of(observation)
        if idx is None:
            raise BinNotFound()
        self._counts[tuple(idx)] += 1

    def ndim(self):
        return self._grid.ndim()

    def counts(self):
        return self._counts

    def grid(self):
        return self._grid


def histogram(a, grid, ctx=None):
    """``HistogramExt::histogram`` (histograms.rs:137-147): the rows of the 2-D array ``a`` (numpy or CUDA tensor) binned
    on the device; rows outside the grid are skipped."""
    from . import _ctx_for
    if len(a.shape) != 2:
        raise ValueError("histogram is defined on 2-D arrays (observations x dimensions)")
    counts = _ctx_for(a, ctx).histogram_counts(a, [p.edges.as_array_view() for p in grid.projections()])
    return Histogram(grid, counts)
