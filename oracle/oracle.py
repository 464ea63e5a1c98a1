"""TEST INFRASTRUCTURE — numpy/ctypes front end of the CPU oracle (oracle/ndstats_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  It never touches the product library.  Parity pinning: partial — see the
header of ndstats_oracle.c.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libndstats_oracle.so")

OK, EMPTY_INPUT, INVALID_QUANTILE, NAN_IN_INPUT, UNSUPPORTED = 0, 1, 2, 3, 4
BAD_ARGUMENT, PANIC = 7, 8
LOWER, HIGHER, NEAREST, MIDPOINT, LINEAR = range(5)
INTERP = {"lower": LOWER, "higher": HIGHER, "nearest": NEAREST, "midpoint": MIDPOINT, "linear": LINEAR}

DTYPES = {
    np.dtype(np.int8): 0, np.dtype(np.int16): 1, np.dtype(np.int32): 2, np.dtype(np.int64): 3,
    np.dtype(np.uint8): 4, np.dtype(np.uint16): 5, np.dtype(np.uint32): 6, np.dtype(np.uint64): 7,
    np.dtype(np.float32): 8, np.dtype(np.float64): 9,
}


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc; seconds)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("ndstats_oracle.c", "oracle_typed.inc", "Makefile")
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.oracle_quantiles_axis.restype = ctypes.c_int
        L.oracle_quantiles_axis.argtypes = [
            ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t),
            ctypes.POINTER(ctypes.c_ssize_t), ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_size_t,
            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t), ctypes.c_int,
            ctypes.c_int,
        ]
        L.oracle_partition_mut_i64.restype = ctypes.c_size_t
        L.oracle_partition_mut_i64.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_ssize_t, ctypes.c_size_t]
        L.oracle_get_many_from_sorted_i64.restype = ctypes.c_size_t
        L.oracle_get_many_from_sorted_i64.argtypes = [
            ctypes.c_void_p, ctypes.c_size_t, ctypes.c_ssize_t, ctypes.POINTER(ctypes.c_size_t), ctypes.c_size_t,
            ctypes.POINTER(ctypes.c_size_t), ctypes.c_void_p,
        ]
        L.oracle_remove_nan_mut_f64.restype = ctypes.c_size_t
        L.oracle_remove_nan_mut_f64.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_ssize_t]
        for name in ("oracle_lower_index_pub", "oracle_higher_index_pub"):
            getattr(L, name).restype = ctypes.c_size_t
            getattr(L, name).argtypes = [ctypes.c_double, ctypes.c_size_t]
        L.oracle_index_fraction_pub.restype = ctypes.c_double
        L.oracle_index_fraction_pub.argtypes = [ctypes.c_double, ctypes.c_size_t]
        L.oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, status, bad_q=None):
        super().__init__(f"oracle status {status} (bad_q={bad_q})")
        self.status = status
        self.bad_q = bad_q


def _lowest_address_view(a):
    """Element offset (from the lowest touched address) of a's logical origin."""
    off = 0
    for n, s in zip(a.shape, a.strides):
        if s < 0 and n > 0:
            off += (n - 1) * (-s)
    return off


def quantiles_axis(a, axis, qs, interp, skipnan=False, select_impl=0, nthreads=1, mutate=False):
    """Reference semantics of quantiles_axis_mut / quantile_axis_skipnan_mut on a numpy array.

    Returns (status, out, bad_q).  ``out`` is C-order with shape[axis] = len(qs) (None on error).
    The input is copied unless mutate=True (the reference permutes lanes in place).
    """
    a = np.asarray(a)
    if a.dtype not in DTYPES:
        return UNSUPPORTED, None, None
    if isinstance(interp, str):
        interp = INTERP[interp]
    work = a if mutate else a.copy(order="K")
    if not mutate and work.strides != a.strides and a.size:
        # copy(order="K") may normalise negative strides; rebuild the same logical view
        work = np.array(a, order="C", copy=True)
    qs = np.ascontiguousarray(qs, dtype=np.float64).reshape(-1)
    ndim = work.ndim
    itemsize = work.dtype.itemsize
    shape = (ctypes.c_size_t * ndim)(*work.shape)
    strides = (ctypes.c_ssize_t * ndim)(*[s // itemsize for s in work.strides])
    out_shape = list(work.shape)
    out_shape[axis] = qs.size
    out = np.zeros(out_shape, dtype=work.dtype)
    bad = ctypes.c_size_t(0)
    st = lib().oracle_quantiles_axis(
        DTYPES[work.dtype], ctypes.c_void_p(work.__array_interface__["data"][0]), ndim, shape, strides, axis,
        qs.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), qs.size, interp, int(bool(skipnan)),
        ctypes.c_void_p(out.ctypes.data), ctypes.byref(bad), select_impl, nthreads,
    )
    if st == INVALID_QUANTILE:
        return st, None, bad.value
    if st != OK:
        return st, None, None
    return st, out, None


def partition_mut_i64(arr, pivot_index):
    arr = np.asarray(arr)
    assert arr.dtype == np.int64 and arr.ndim == 1
    return lib().oracle_partition_mut_i64(
        ctypes.c_void_p(arr.__array_interface__["data"][0]), arr.size, arr.strides[0] // 8, pivot_index)


def get_many_from_sorted_i64(arr, indexes):
    arr = np.array(arr, dtype=np.int64)
    idx = np.ascontiguousarray(indexes, dtype=np.uintp)
    out_idx = (ctypes.c_size_t * max(1, idx.size))()
    out_val = np.zeros(max(1, idx.size), dtype=np.int64)
    m = lib().oracle_get_many_from_sorted_i64(
        ctypes.c_void_p(arr.ctypes.data), arr.size, 1, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_size_t)),
        idx.size, out_idx, ctypes.c_void_p(out_val.ctypes.data))
    return list(out_idx[:m]), out_val[:m].copy()


def remove_nan_mut_f64(view):
    """In place on a (possibly strided / reversed) 1-D float64 view; returns the NaN-free prefix view."""
    assert view.dtype == np.float64 and view.ndim == 1
    n = lib().oracle_remove_nan_mut_f64(
        ctypes.c_void_p(view.__array_interface__["data"][0]), view.size, view.strides[0] // 8 if view.size else 1)
    return view[:n]


def lower_index(q, n):
    return lib().oracle_lower_index_pub(q, n)


def higher_index(q, n):
    return lib().oracle_higher_index_pub(q, n)


def index_fraction(q, n):
    return lib().oracle_index_fraction_pub(q, n)


def max_threads():
    return lib().oracle_max_threads()


# ---- QuantileExt::{min, max, argmin, argmax} and the *_skipnan variants -----------------
# TEST INFRASTRUCTURE like the rest of this module.  A plain restatement of src/quantile/mod.rs:283-409: the array is
# walked in logical (C) order, the current extremum is replaced only on a strict improvement, so among equal extrema
# the first one wins and -0.0 == +0.0; a NaN makes the plain calls fail (every element is compared), the skipnan
# calls ignore it.  Returns (status, value, index tuple): status "ok" | "EmptyInput" | "UndefinedOrder".
def minmax(a, want_max=False, skipnan=False):
    a = np.asarray(a)
    if a.size == 0:
        return "EmptyInput", (np.nan if skipnan and a.dtype.kind == "f" else None), None   # mod.rs:287, 336-345
    flat = a.reshape(-1)  # logical C order (copy for strided views)
    isnan = np.isnan(flat) if a.dtype.kind == "f" else np.zeros(flat.size, dtype=bool)
    if not skipnan and isnan.any():
        return "UndefinedOrder", None, None                                               # mod.rs:292, 321
    if isnan.all():
        return "EmptyInput", np.nan, None                                                 # mod.rs:314-318 / NaN for min_skipnan
    vals = np.where(isnan, (-np.inf if want_max else np.inf), flat) if a.dtype.kind == "f" else flat
    i = int(np.argmax(vals) if want_max else np.argmin(vals))   # numpy: first occurrence, -0.0 == +0.0
    value = flat[i]
    if want_max and skipnan:
        # max_skipnan folds with `acc.max(elem)` (mod.rs:399-409) and Ord::max returns its SECOND operand on a tie: the
        # value is the last of the equal maxima (only +-0.0 can tell); argmax_skipnan keeps the first (mod.rs:365-382)
        value = flat[flat.size - 1 - int(np.argmax(vals[::-1]))]
    return "ok", value, tuple(int(c) for c in np.unravel_index(i, a.shape))


# ---- histogram::strategies::FreedmanDiaconis::from_array (src/histogram/strategies.rs:394-433) ----
# TEST INFRASTRUCTURE.  Quartiles through the restated quantile path above (two single-q Nearest calls, like the
# reference), the rest in the element type.  Returns ("ok", bin_width, min, max, n_bins) or ("EmptyInput"|"Strategy",).
def freedman_diaconis(a):
    a = np.asarray(a)
    n = a.size
    if n == 0:
        return ("EmptyInput",)
    st, q1, _ = quantiles_axis(a, 0, [0.25], NEAREST)
    st2, q3, _ = quantiles_axis(a, 0, [0.75], NEAREST)
    assert st == OK and st2 == OK
    with np.errstate(over="ignore"):
        iqr = q3[0] - q1[0]
        denom = float(n) ** (1.0 / 3.0)
        if a.dtype.kind in "iu":
            d = a.dtype.type(int(denom))                      # T::from_f64 truncates
            bw = a.dtype.type(2) * iqr // d
        else:
            bw = a.dtype.type(2) * iqr / a.dtype.type(denom)
    lo, hi = a.min(), a.max()
    if bw <= 0 or lo >= hi:                                   # EquiSpaced::new, strategies.rs:221-222
        return ("Strategy",)
    edge, nb = lo, 0
    while edge <= hi:                                         # strategies.rs:243-251
        edge = edge + bw
        nb += 1
    return ("ok", bw, lo, hi, nb)


# ---- Sqrt / Rice / Sturges / Auto (src/histogram/strategies.rs:256-371, 439-503).  TEST INFRASTRUCTURE. ----
def _rust_round_usize(x):
    import math
    return int(math.floor(x + 0.5)) if x >= 0 else 0


def count_rule(a, rule):
    """("ok", bin_width, min, max, n_bins) or ("EmptyInput"|"Strategy",) for rule in {"sqrt", "rice", "sturges"}."""
    import math
    a = np.asarray(a)
    n = a.size
    if n == 0:
        return ("EmptyInput",)                                    # a.min()? on an empty array
    nb = {"sqrt": lambda: _rust_round_usize(math.sqrt(n)), "rice": lambda: _rust_round_usize(2.0 * n ** (1.0 / 3.0)),
          "sturges": lambda: _rust_round_usize(math.log2(n)) + 1}[rule]()
    lo, hi = a.min(), a.max()
    with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
        rng = hi - lo
        bw = a.dtype.type(int(rng) // nb) if a.dtype.kind in "iu" else rng / a.dtype.type(nb)
    if bw <= 0 or lo >= hi:
        return ("Strategy",)
    edge, k = lo, 0
    while edge <= hi:
        edge = edge + bw
        k += 1
    return ("ok", bw, lo, hi, k)


def auto_rule(a):
    fd, st = freedman_diaconis(a), count_rule(a, "sturges")
    if fd[0] != "ok" and st[0] != "ok":
        return fd
    if fd[0] != "ok":
        return ("sturges",) + st[1:]
    if st[0] != "ok":
        return ("fd",) + fd[1:]
    return (("sturges",) + st[1:]) if fd[1] > st[1] else (("fd",) + fd[1:])


# ---- HistogramExt::histogram (src/histogram/histograms.rs:137-147).  TEST INFRASTRUCTURE: the reference's loop restated
# with numpy -- per dimension the bin of Edges::indices_of (left-closed, right-open, bins.rs:217-229), rows outside skipped.
def histogram_counts(a, edges):
    a = np.asarray(a)
    edges = [np.unique(np.asarray(e, dtype=a.dtype)) for e in edges]
    shape = tuple(max(e.size - 1, 0) for e in edges)
    counts = np.zeros(shape, dtype=np.uint64)
    if counts.size == 0 or a.shape[0] == 0:
        return counts
    inside = np.ones(a.shape[0], dtype=bool)
    idx = []
    for d, e in enumerate(edges):
        v = a[:, d]
        b = np.searchsorted(e, v, side="right").astype(np.int64) - 1      # number of edges <= v, minus one
        with np.errstate(invalid="ignore"):
            ok = (v >= e[0]) & (v < e[-1])
        inside &= ok
        idx.append(np.where(ok, b, 0))
    np.add.at(counts, tuple(i[inside] for i in idx), 1)
    return counts
