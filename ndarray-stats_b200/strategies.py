"""histogram::strategies — the in-tree CALLERS of the quantile / min / max path:
FreedmanDiaconis (quartiles + extremes), Sqrt / Rice / Sturges (extremes), Auto (the narrower of FreedmanDiaconis and
Sturges).

Mirror of ``FreedmanDiaconis::from_array`` (src/histogram/strategies.rs:394-433) and of the ``EquiSpaced`` builder it
wraps (strategies.rs:214-254).  The two quartiles come from ONE bulk device call (``quantiles_mut([0.25, 0.75],
&Nearest)`` — the reference issues two single-q calls on its copy, the values are the same), minimum and maximum from
the device reductions; the few scalar operations that follow are done on the host in the element type, with the
reference's conversions (``T::from_usize``, ``T::from_f64``: truncation for integers).
"""
import numpy as np


class BinsBuildError(Exception):
    """src/histogram/errors.rs:23-48"""

    def __init__(self, kind):
        super().__init__(kind)
        self.kind = kind

    def is_empty_input(self):
        return self.kind == "EmptyInput"

    def is_strategy(self):
        return self.kind == "Strategy"


def _from_f64(dtype, x):
    """num-traits FromPrimitive::from_f64 into the element type: floats round, integers truncate toward zero."""
    dtype = np.dtype(dtype)
    return dtype.type(x) if dtype.kind == "f" else dtype.type(int(x))


class EquiSpaced:
    """strategies.rs:214-254"""

    def __init__(self, bin_width, lo, hi):
        if bin_width <= 0 or lo >= hi:          # strategies.rs:221-222
            raise BinsBuildError("Strategy")
        self.bin_width, self.min, self.max = bin_width, lo, hi

    def n_bins(self):                           # strategies.rs:243-251: repeated addition in the element type
        max_edge, n = self.min, 0
        while max_edge <= self.max:
            max_edge = max_edge + self.bin_width
            n += 1
        return n

    def build(self):                            # strategies.rs:232-241: min + from_usize(i) * bin_width
        t = type(self.bin_width)
        return np.array([self.min + t(i) * self.bin_width for i in range(self.n_bins() + 1)], dtype=np.dtype(t))


class FreedmanDiaconis:
    """bin_width = 2 * IQR * n^(-1/3) (strategies.rs:156-180, 394-433)."""

    def __init__(self, builder):
        self.builder = builder

    @staticmethod
    def compute_bin_width(n_points, iqr):       # strategies.rs:421-427
        dtype = np.asarray(iqr).dtype
        denominator = float(n_points) ** (1.0 / 3.0)
        with np.errstate(over="ignore"):
            return dtype.type(2) * iqr // _from_f64(dtype, denominator) if dtype.kind in "iu" \
                else dtype.type(2) * iqr / _from_f64(dtype, denominator)

    @classmethod
    def from_array(cls, a, ctx=None):
        """``a``: 1-D numpy array or CUDA tensor of a NaN-free ordered type."""
        from . import Nearest, _ctx_for, _is_torch, NDS_EMPTY_INPUT, NDS_UNDEFINED_ORDER
        if len(a.shape) != 1:
            raise ValueError("from_array is defined on 1-D arrays")
        n_points = int(a.shape[0])
        if n_points == 0:
            raise BinsBuildError("EmptyInput")  # strategies.rs:397-399
        c = _ctx_for(a, ctx)
        q = c.quantiles_axis(a, 0, [0.25, 0.75], Nearest)        # strategies.rs:402-403
        q = q.cpu().numpy() if _is_torch(q) else q
        with np.errstate(over="ignore"):
            iqr = q[1] - q[0]
        bin_width = cls.compute_bin_width(n_points, iqr)
        st_lo, lo, _ = c.minmax(a, False, False)                 # a.min()?, a.max()?  strategies.rs:407-408
        st_hi, hi, _ = c.minmax(a, True, False)
        if NDS_UNDEFINED_ORDER in (st_lo, st_hi):
            raise BinsBuildError("Strategy")                     # From<MinMaxError>, src/histogram/errors.rs
        if NDS_EMPTY_INPUT in (st_lo, st_hi):
            raise BinsBuildError("EmptyInput")
        return cls(EquiSpaced(bin_width, lo, hi))

    def build(self):
        return self.builder.build()

    def n_bins(self):
        return self.builder.n_bins()

    def bin_width(self):
        return self.builder.bin_width


def _round_half_away(x):
    """f64::round (half away from zero) followed by `as usize` (saturating at 0)."""
    import math
    r = math.floor(abs(x) + 0.5)
    return int(r) if x >= 0 else 0


def _min_max(a, ctx):
    """a.min()? / a.max()? with From<MinMaxError> for BinsBuildError (src/histogram/errors.rs)."""
    from . import _ctx_for, NDS_EMPTY_INPUT, NDS_UNDEFINED_ORDER
    c = _ctx_for(a, ctx)
    st_lo, lo, _ = c.minmax(a, False, False)
    st_hi, hi, _ = c.minmax(a, True, False)
    if NDS_EMPTY_INPUT in (st_lo, st_hi):
        raise BinsBuildError("EmptyInput")
    if NDS_UNDEFINED_ORDER in (st_lo, st_hi):
        raise BinsBuildError("Strategy")
    return lo, hi


def compute_bin_width(lo, hi, n_bins):
    """(max - min) / T::from_usize(n_bins) in the element type (strategies.rs:505-511)."""
    dtype = np.asarray(lo).dtype
    with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
        rng = hi - lo
        if dtype.kind in "iu":
            if n_bins == 0:
                raise ZeroDivisionError("attempt to divide by zero")   # the reference panics
            q = abs(int(rng)) // n_bins                                 # Rust integer division truncates toward zero
            return dtype.type(q if rng >= 0 else -q)
        return rng / dtype.type(n_bins)


class _CountRule:
    """Sqrt / Rice / Sturges: a number of bins from the number of elements, equi-spaced over [min, max]."""

    def __init__(self, builder):
        self.builder = builder

    @staticmethod
    def n_bins_rule(n_elems):
        raise NotImplementedError

    @classmethod
    def from_array(cls, a, ctx=None):
        if len(a.shape) != 1:
            raise ValueError("from_array is defined on 1-D arrays")
        n_elems = int(a.shape[0])
        n_bins = cls.n_bins_rule(n_elems) if n_elems else 0
        lo, hi = _min_max(a, ctx)           # EmptyInput surfaces here, as in the reference (a.min()?)
        return cls(EquiSpaced(compute_bin_width(lo, hi, n_bins), lo, hi))

    def build(self):
        return self.builder.build()

    def n_bins(self):
        return self.builder.n_bins()

    def bin_width(self):
        return self.builder.bin_width


class Sqrt(_CountRule):
    """n_bins = round(sqrt(n)) (strategies.rs:256-285)."""

    @staticmethod
    def n_bins_rule(n_elems):
        return _round_half_away(float(n_elems) ** 0.5)


class Rice(_CountRule):
    """n_bins = round(2 n^(1/3)) (strategies.rs:299-328)."""

    @staticmethod
    def n_bins_rule(n_elems):
        return _round_half_away(2.0 * float(n_elems) ** (1.0 / 3.0))


class Sturges(_CountRule):
    """n_bins = round(log2 n) + 1 (strategies.rs:342-371)."""

    @staticmethod
    def n_bins_rule(n_elems):
        import math
        return _round_half_away(math.log2(float(n_elems))) + 1


class Auto:
    """The strategy with the smaller bin width among FreedmanDiaconis and Sturges; whichever exists if the other fails
    (strategies.rs:439-503)."""

    def __init__(self, builder):
        self.builder = builder

    @classmethod
    def from_array(cls, a, ctx=None):
        fd = st = fd_err = None
        try:
            fd = FreedmanDiaconis.from_array(a, ctx=ctx)
        except BinsBuildError as e:
            fd_err = e
        try:
            st = Sturges.from_array(a, ctx=ctx)
        except BinsBuildError:
            st = None
        if fd is None and st is None:
            raise fd_err
        if fd is None:
            return cls(st)
        if st is None:
            return cls(fd)
        return cls(st if fd.bin_width() > st.bin_width() else fd)

    def build(self):
        return self.builder.build()

    def n_bins(self):
        return self.builder.n_bins()

    def bin_width(self):
        return self.builder.bin_width()
