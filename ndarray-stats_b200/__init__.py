"""ndarray_stats_b200 — B200-native bulk quantile path of rust-ndarray/ndarray-stats.

Host-side mirror of the reference's extension-trait surface over the C ABI of
``libndstats_b200.so`` (``include/ndstats_b200.h``).  Same names, argument meaning and error behaviour
as the reference (paths relative to the reference checkout):

=================================  ==========================================
``quantile_axis_mut``              src/quantile/mod.rs:495-508
``quantiles_axis_mut``             src/quantile/mod.rs:417-493
``quantile_axis_skipnan_mut``      src/quantile/mod.rs:510-544
``quantile_mut`` / ``quantiles_mut``  src/quantile/mod.rs:612-633 (Quantile1dExt)
``Lower`` … ``Linear``             src/quantile/interpolate.rs:51-62
``QuantileError``                  src/errors.rs:119-143
``get_from_sorted_mut`` / ``get_many_from_sorted_mut``  src/sort.rs:99-131 (Sort1dExt)
``partition_mut``                                          src/sort.rs:133-168 (Sort1dExt)
=================================  ==========================================

Arrays are numpy arrays (host; staged through the context's device arena) or torch CUDA tensors
(device-resident; nothing is copied).  There is no CPU fallback: importing works anywhere, but every
compute call needs the CUDA library and a GPU and raises otherwise.
"""
import builtins
import ctypes
import os

import numpy as np

from . import _cabi
from ._cabi import (  # noqa: F401  (re-exported status codes)
    NDS_BAD_ARGUMENT, NDS_CAST_OVERFLOW, NDS_CUDA_ERROR, NDS_EMPTY_INPUT, NDS_INDEX_OUT_OF_BOUNDS,
    NDS_INVALID_QUANTILE, NDS_NAN_IN_INPUT, NDS_NCCL_ERROR, NDS_OK, NDS_UNDEFINED_ORDER, NDS_UNSUPPORTED, library_path,
)

__all__ = [
    "Axis", "Lower", "Higher", "Nearest", "Midpoint", "Linear", "QuantileError", "EmptyInput", "InvalidQuantile",
    "NanInInput", "CastOverflow", "Context", "default_context", "quantile_axis_mut", "quantiles_axis_mut",
    "quantile_axis_skipnan_mut", "quantile_mut", "quantiles_mut", "get_from_sorted_mut", "get_many_from_sorted_mut", "partition_mut",
    "shard_lanes", "MinMaxError", "MinMaxEmptyInput", "UndefinedOrder", "argmin", "argmax", "min", "max",
    "argmin_skipnan", "argmax_skipnan", "min_skipnan", "max_skipnan", "FreedmanDiaconis", "Sqrt", "Rice", "Sturges", "Auto", "EquiSpaced", "BinsBuildError",
    "Edges", "Bins", "Grid", "GridBuilder", "Histogram", "BinNotFound", "histogram",
]


# ---- the reference's small vocabulary ---------------------------------------------------------------
class Axis(int):
    """ndarray::Axis"""


class _Interp:
    code = -1

    def __repr__(self):
        return type(self).__name__


class Lower(_Interp):
    """Select the lower value (interpolate.rs:77-88)."""
    code = 0


class Higher(_Interp):
    """Select the higher value (interpolate.rs:64-75)."""
    code = 1


class Nearest(_Interp):
    """Select the nearest value; a fraction of exactly 0.5 goes up (interpolate.rs:90-105)."""
    code = 2


class Midpoint(_Interp):
    """lower + (higher - lower) / 2 in the element type (interpolate.rs:107-124)."""
    code = 3


class Linear(_Interp):
    """lower + T::from_f64(fraction * (higher - lower)) (interpolate.rs:126-144)."""
    code = 4


def _interp_code(interpolate):
    if isinstance(interpolate, type) and issubclass(interpolate, _Interp):
        return interpolate.code
    if isinstance(interpolate, _Interp):
        return interpolate.code
    if isinstance(interpolate, str):
        return {"lower": 0, "higher": 1, "nearest": 2, "midpoint": 3, "linear": 4}[interpolate.lower()]
    raise TypeError(f"not an interpolation strategy: {interpolate!r}")


class QuantileError(Exception):
    """src/errors.rs:119-125"""


class EmptyInput(QuantileError):
    def __str__(self):
        return "Empty input."  # errors.rs:128


class InvalidQuantile(QuantileError):
    def __init__(self, q):
        super().__init__(q)
        self.q = q

    def __str__(self):
        return f"{self.q} is not between 0. and 1. (inclusive)."  # errors.rs:129-131


class MinMaxError(Exception):
    """src/errors.rs:18-26"""


class MinMaxEmptyInput(MinMaxError, EmptyInput):
    """MinMaxError::EmptyInput (also an EmptyInput: `impl From<EmptyInput> for MinMaxError`, errors.rs:40-44)"""

    def __str__(self):
        return "Empty input."


class UndefinedOrder(MinMaxError):
    def __str__(self):
        return "Undefined ordering between a tested pair of values."  # errors.rs:31-33


class NanInInput(ValueError):
    """A NaN reached a non-skipnan call: N32/N64 cannot hold one (the reference's types forbid it)."""


class CastOverflow(OverflowError):
    """Integer Linear interpolation left the element range (the reference panics, interpolate.rs:142)."""


class NdsError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"ndstats_b200 status {status}: {message}")
        self.status = status


_NP_DTYPES = {
    np.dtype(np.int8): 0, np.dtype(np.int16): 1, np.dtype(np.int32): 2, np.dtype(np.int64): 3,
    np.dtype(np.uint8): 4, np.dtype(np.uint16): 5, np.dtype(np.uint32): 6, np.dtype(np.uint64): 7,
    np.dtype(np.float32): 8, np.dtype(np.float64): 9,
}


def _torch_dtype_code(t):
    import torch
    table = {
        torch.int8: 0, torch.int16: 1, torch.int32: 2, torch.int64: 3, torch.uint8: 4,
        torch.float32: 8, torch.float64: 9,
    }
    for name, code in (("uint16", 5), ("uint32", 6), ("uint64", 7)):
        if hasattr(torch, name):
            table[getattr(torch, name)] = code
    if t.dtype not in table:
        raise TypeError(f"unsupported tensor dtype {t.dtype}")
    return table[t.dtype]


def _is_torch(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


class Context:
    """One GPU, one stream, one scratch arena (``nds_ctx``).  Not thread-safe, like the C object."""

    def __init__(self, device=0):
        self._lib = _cabi.load()  # raises if the CUDA library is not built: no fallback
        handle = ctypes.c_void_p()
        st = self._lib.nds_ctx_create(int(device), ctypes.byref(handle))
        if st != NDS_OK:
            raise NdsError(st, "nds_ctx_create failed (no CUDA device? this package has no CPU path)")
        self._h = handle
        self.device = int(device)

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.nds_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- plumbing -------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        self._lib.nds_ctx_set_stream(self._h, ctypes.c_void_p(cuda_stream_ptr or 0))

    def synchronize(self):
        self._raise(self._lib.nds_ctx_synchronize(self._h))

    @property
    def kernel_launches(self):
        return int(self._lib.nds_ctx_kernel_launches(self._h))

    @property
    def last_path(self):
        return self._lib.nds_ctx_last_path(self._h).decode()

    def force_path(self, name):
        self._raise(self._lib.nds_ctx_force_path(self._h, (name or "auto").encode()))

    def _raise(self, st, qs=None, bad=None):
        if st == NDS_OK:
            return
        if st == NDS_EMPTY_INPUT:
            raise EmptyInput()
        if st == NDS_INVALID_QUANTILE:
            raise InvalidQuantile(float(qs[bad]) if qs is not None else float("nan"))
        msg = self._lib.nds_ctx_last_error(self._h).decode()
        if st == NDS_NAN_IN_INPUT:
            raise NanInInput(msg)
        if st == NDS_CAST_OVERFLOW:
            raise CastOverflow(msg)
        if st == NDS_INDEX_OUT_OF_BOUNDS:
            raise IndexError(msg)
        raise NdsError(st, msg)

    # -- the C entry points, array-typed --------------------------------------------------------
    def _describe(self, a):
        """-> (dtype code, data pointer, shape, element strides, keepalive, is_torch)"""
        if _is_torch(a):
            if not a.is_cuda:
                raise TypeError("torch inputs must be CUDA tensors (pass numpy arrays for host data)")
            import torch
            # torch's default stream has handle 0 = the legacy default stream: name it explicitly (cudaStreamLegacy),
            # NULL would mean "the ctx's own stream" and lose the ordering with the kernels that produced `a`
            self.set_stream(torch.cuda.current_stream(a.device).cuda_stream or 1)
            return _torch_dtype_code(a), a.data_ptr(), tuple(a.shape), tuple(a.stride()), a, True
        a = np.asarray(a)
        if a.dtype not in _NP_DTYPES:
            raise TypeError(f"unsupported dtype {a.dtype}")
        item = a.dtype.itemsize
        if any(s % item for s in a.strides):
            raise TypeError("strides must be multiples of the element size")
        ptr = a.__array_interface__["data"][0]
        return _NP_DTYPES[a.dtype], ptr, a.shape, tuple(s // item for s in a.strides), a, False

    def quantiles_axis(self, a, axis, qs, interpolate, skipnan=False, enqueue=False, out=None):
        """``nds_quantiles_axis`` / ``nds_quantiles_axis_enqueue`` on a numpy array or CUDA tensor.

        Returns the C-order result with shape[axis] = len(qs) (same container kind as the input).

        ``enqueue=True`` (CUDA tensors only) returns without a host synchronisation.  For short and column lanes the
        result is complete in stream order.  For a long lane on the one-pass bracket path it is PROVISIONAL until
        ``synchronize()``: in the rare case that a rank escaped its sampled bracket the call is re-run there on the
        radix passes, which re-reads ``a`` -- keep ``a`` unmodified and read ``out`` only after ``synchronize()``
        (or ``force_path("long_radix")`` for strict stream order)."""
        code, ptr, shape, strides, keep, is_torch = self._describe(a)
        ndim = len(shape)
        axis = int(axis)
        if not 0 <= axis < ndim:
            raise IndexError(f"axis {axis} out of bounds for a {ndim}-d array")  # the reference panics
        qs_arr = np.ascontiguousarray(qs, dtype=np.float64).reshape(-1)
        out_shape = list(shape)
        out_shape[axis] = qs_arr.size
        if out is None:
            if is_torch:
                import torch
                out = torch.empty(out_shape, dtype=a.dtype, device=a.device)
            else:
                out = np.empty(out_shape, dtype=keep.dtype)
        out_ptr = out.data_ptr() if is_torch else out.ctypes.data
        c_shape = (ctypes.c_size_t * ndim)(*shape)
        c_strides = (ctypes.c_ssize_t * ndim)(*strides)
        bad = ctypes.c_size_t(0)
        fn = self._lib.nds_quantiles_axis_enqueue if enqueue else self._lib.nds_quantiles_axis
        st = fn(self._h, code, ctypes.c_void_p(ptr), ndim, c_shape, c_strides, axis,
                qs_arr.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), qs_arr.size, _interp_code(interpolate),
                int(bool(skipnan)), ctypes.c_void_p(out_ptr), ctypes.byref(bad))
        self._raise(st, qs_arr, bad.value)
        return out

    def quantiles_axis_masked(self, a, valid, axis, qs, interpolate):
        """``nds_quantiles_axis_masked``: Option<T> lanes as (values, validity mask).  numpy only.
        Returns (out, out_valid)."""
        a = np.asarray(a)
        valid = np.asarray(valid)
        if valid.shape != a.shape:
            raise ValueError("mask shape differs from data shape")
        # the C ABI wants the mask on the data's element strides: lay it out over the same span
        item = a.dtype.itemsize
        lo, hi = self._span_elems(a)
        mask = np.zeros(hi - lo + 1, dtype=np.uint8)
        mask_view = np.lib.stride_tricks.as_strided(
            mask[-lo:], shape=a.shape, strides=tuple(s // item for s in a.strides), writeable=True)
        if a.size:
            mask_view[...] = valid.astype(bool)
        code, ptr, shape, strides, keep, _ = self._describe(a)
        ndim = len(shape)
        axis = int(axis)
        if not 0 <= axis < ndim:
            raise IndexError(f"axis {axis} out of bounds for a {ndim}-d array")
        qs_arr = np.ascontiguousarray(qs, dtype=np.float64).reshape(-1)
        out_shape = list(shape)
        out_shape[axis] = qs_arr.size
        out = np.zeros(out_shape, dtype=a.dtype)
        out_valid = np.zeros(out_shape, dtype=np.uint8)
        bad = ctypes.c_size_t(0)
        mask_ptr = mask.ctypes.data - lo
        st = self._lib.nds_quantiles_axis_masked(
            self._h, code, ctypes.c_void_p(ptr), ctypes.c_void_p(mask_ptr), ndim, (ctypes.c_size_t * ndim)(*shape),
            (ctypes.c_ssize_t * ndim)(*strides), axis, qs_arr.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
            qs_arr.size, _interp_code(interpolate), ctypes.c_void_p(out.ctypes.data),
            ctypes.c_void_p(out_valid.ctypes.data), ctypes.byref(bad))
        self._raise(st, qs_arr, bad.value)
        return out, out_valid.astype(bool)

    @staticmethod
    def _span_elems(a):
        item = a.dtype.itemsize
        lo = hi = 0
        for n, s in zip(a.shape, a.strides):
            if n == 0:
                continue
            ext = (n - 1) * (s // item)
            if ext < 0:
                lo += ext
            else:
                hi += ext
        return lo, hi

    def get_many_from_sorted(self, a, indexes):
        """``nds_get_many_from_sorted``: values at the given sorted positions, in the order given."""
        code, ptr, shape, strides, keep, is_torch = self._describe(a)
        if len(shape) != 1:
            raise ValueError("Sort1dExt is defined on 1-D arrays")
        idx = np.ascontiguousarray(indexes, dtype=np.uintp).reshape(-1)
        if is_torch:
            import torch
            out = torch.empty(idx.size, dtype=a.dtype, device=a.device)
            out_ptr = out.data_ptr()
        else:
            out = np.empty(idx.size, dtype=keep.dtype)
            out_ptr = out.ctypes.data
        st = self._lib.nds_get_many_from_sorted(
            self._h, code, ctypes.c_void_p(ptr), shape[0], strides[0] if shape[0] else 1,
            idx.ctypes.data_as(ctypes.POINTER(ctypes.c_size_t)), idx.size, ctypes.c_void_p(out_ptr))
        self._raise(st)
        return out

    def partition(self, a, pivot_index):
        """``nds_partition``: partitions the 1-D array ``a`` (writable numpy array or CUDA tensor) in place around the
        value at ``pivot_index``; returns the pivot's new position."""
        code, ptr, shape, strides, keep, is_torch = self._describe(a)
        if len(shape) != 1:
            raise ValueError("Sort1dExt is defined on 1-D arrays")
        if not is_torch and not keep.flags.writeable:
            raise ValueError("partition_mut needs a writable array")
        new_index = ctypes.c_size_t(0)
        st = self._lib.nds_partition(self._h, code, ctypes.c_void_p(ptr), shape[0], strides[0] if shape[0] else 1,
                                     int(pivot_index), ctypes.byref(new_index))
        self._raise(st)
        return int(new_index.value)

    def histogram_counts(self, a, edges):
        """``nds_histogram``: counts of the rows of the 2-D array ``a`` (numpy or CUDA tensor) in the grid whose
        dimension d has the sorted edges ``edges[d]``.  Returns a numpy uint64 array of shape (len(e) - 1 for e in edges)."""
        code, ptr, shape, strides, keep, is_torch = self._describe(a)
        if len(shape) != 2:
            raise ValueError("histogram is defined on 2-D arrays (observations x dimensions)")
        n_obs, n_dims = shape
        if len(edges) != n_dims:
            raise ValueError("the grid's dimensionality differs from the observations'")   # the reference panics (grid.rs:204-211)
        np_dtype = keep.dtype if not is_torch else {v: k for k, v in _NP_DTYPES.items()}[code]
        ed = [np.ascontiguousarray(e, dtype=np_dtype).reshape(-1) for e in edges]
        out_shape = tuple(e.size - 1 if e.size else 0 for e in ed)   # (this module's `max` is QuantileExt::max)
        counts = np.zeros(out_shape, dtype=np.uint64)
        if counts.size == 0:
            return counts
        cat = np.concatenate(ed) if ed else np.zeros(0, dtype=np_dtype)
        n_edges = (ctypes.c_size_t * n_dims)(*[e.size for e in ed])
        st = self._lib.nds_histogram(self._h, code, ctypes.c_void_p(ptr if n_obs else 0), n_obs, n_dims,
                                     strides[0] if n_obs else 1, strides[1] if n_obs else 1,
                                     ctypes.c_void_p(cat.ctypes.data), n_edges, ctypes.c_void_p(counts.ctypes.data))
        self._raise(st)
        return counts

    def minmax(self, a, want_max, skipnan, want_index=True):
        """``nds_minmax`` -> (status, value or None, index tuple or None); the callers map the status.
        ``want_index=False`` asks for the value alone (max_skipnan then returns the LAST of equal maxima, as the
        reference's ``acc.max(elem)`` fold does)."""
        code, ptr, shape, strides, keep, is_torch = self._describe(a)
        ndim = len(shape)
        np_dtype = keep.dtype if not is_torch else {v: k for k, v in _NP_DTYPES.items()}[code]
        val = np.zeros(1, dtype=np_dtype)
        idx = (ctypes.c_size_t * (ndim or 1))()
        st = self._lib.nds_minmax(self._h, code, ctypes.c_void_p(ptr if a_size(shape) else 0), ndim,
                                  (ctypes.c_size_t * (ndim or 1))(*shape), (ctypes.c_ssize_t * (ndim or 1))(*strides),
                                  int(bool(want_max)), int(bool(skipnan)), ctypes.c_void_p(val.ctypes.data),
                                  idx if want_index else None)
        if st not in (NDS_OK, NDS_EMPTY_INPUT, NDS_UNDEFINED_ORDER):
            self._raise(st)
        return st, val[0], tuple(int(idx[d]) for d in range(ndim))

    # -- one 1-D array sharded over ranks -------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = (ctypes.c_ubyte * 128)()
        st = _cabi.load().nds_comm_get_unique_id(buf)
        if st != NDS_OK:
            raise NdsError(st, "nds_comm_get_unique_id failed (NCCL not loadable?)")
        return bytes(buf)

    def comm_init_rank(self, nranks, rank, unique_id):
        buf = (ctypes.c_ubyte * 128).from_buffer_copy(unique_id)
        self._raise(self._lib.nds_comm_init_rank(self._h, int(nranks), int(rank), buf))

    def comm_destroy(self):
        self._lib.nds_comm_destroy(self._h)

    def comm_info(self):
        """``nds_comm_stats``: how sharded calls exchange data between the ranks, and the counters so far."""
        out = (ctypes.c_uint64 * 4)()
        self._raise(self._lib.nds_comm_stats(self._h, out))
        return {"transport": "peer memory (own kernels over NVLink)" if out[0] else ("nccl" if out[3] > 1 else "none (one rank)"),
                "exchange_steps": int(out[1]), "host_syncs": int(out[2]), "ranks": int(out[3])}

    def quantiles_1d_sharded(self, local, qs, interpolate, skipnan=False):
        """``nds_quantiles_1d_sharded``: collective over every rank of the communicator."""
        code, ptr, shape, strides, keep, is_torch = self._describe(local)
        if len(shape) != 1 or (shape[0] > 1 and strides[0] != 1):
            raise ValueError("shards must be contiguous 1-D arrays")
        qs_arr = np.ascontiguousarray(qs, dtype=np.float64).reshape(-1)
        if is_torch:
            import torch
            out = torch.empty(qs_arr.size, dtype=local.dtype, device=local.device)
            out_ptr = out.data_ptr()
        else:
            out = np.empty(qs_arr.size, dtype=keep.dtype)
            out_ptr = out.ctypes.data
        bad = ctypes.c_size_t(0)
        st = self._lib.nds_quantiles_1d_sharded(
            self._h, code, ctypes.c_void_p(ptr if shape[0] else 0), shape[0],
            qs_arr.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), qs_arr.size, _interp_code(interpolate),
            int(bool(skipnan)), ctypes.c_void_p(out_ptr), ctypes.byref(bad))
        self._raise(st, qs_arr, bad.value)
        return out


def a_size(shape):
    n = 1
    for d in shape:
        n *= int(d)
    return n


_default = {}


def default_context(device=0):
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]


def _ctx_for(a, ctx):
    if ctx is not None:
        return ctx
    if _is_torch(a) and a.is_cuda:
        return default_context(a.device.index or 0)
    return default_context(int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("NDS_USE_LOCAL_RANK") else 0)


def _drop_axis(out, axis):
    if _is_torch(out):
        return out.select(axis, 0)
    return np.take(out, 0, axis=axis)


# ---- QuantileExt ----------------------------------------------------------------------------------------
def quantiles_axis_mut(a, axis, qs, interpolate, ctx=None):
    """QuantileExt::quantiles_axis_mut (src/quantile/mod.rs:417-493).  ``a`` is left unchanged."""
    return _ctx_for(a, ctx).quantiles_axis(a, axis, qs, interpolate, skipnan=False)


def quantile_axis_mut(a, axis, q, interpolate, ctx=None):
    """QuantileExt::quantile_axis_mut (src/quantile/mod.rs:495-508): the result drops ``axis``."""
    return _drop_axis(_ctx_for(a, ctx).quantiles_axis(a, axis, [q], interpolate, skipnan=False), int(axis))


def quantile_axis_skipnan_mut(a, axis, q, interpolate, ctx=None):
    """QuantileExt::quantile_axis_skipnan_mut (src/quantile/mod.rs:510-544).

    ``a`` may be a float array (NaN = missing) or a pair ``(values, valid_mask)`` standing for the
    reference's ``Option<T>`` elements; for the latter the result is ``(values, valid_mask)`` too."""
    if isinstance(a, tuple):
        vals, mask = a
        out, ov = _ctx_for(vals, ctx).quantiles_axis_masked(vals, mask, axis, [q], interpolate)
        return _drop_axis(out, int(axis)), _drop_axis(ov, int(axis))
    return _drop_axis(_ctx_for(a, ctx).quantiles_axis(a, axis, [q], interpolate, skipnan=True), int(axis))


# ---- Quantile1dExt --------------------------------------------------------------------------------------
def _check_1d(a):
    if len(a.shape) != 1:
        raise ValueError("Quantile1dExt is defined on 1-D arrays")


def quantile_mut(a, q, interpolate, ctx=None):
    """Quantile1dExt::quantile_mut (src/quantile/mod.rs:613-621) -> scalar."""
    _check_1d(a)
    out = quantile_axis_mut(a, 0, q, interpolate, ctx)
    return out.item() if _is_torch(out) else out[()]


def quantiles_mut(a, qs, interpolate, ctx=None):
    """Quantile1dExt::quantiles_mut (src/quantile/mod.rs:623-633) -> 1-D array."""
    _check_1d(a)
    return quantiles_axis_mut(a, 0, qs, interpolate, ctx)


# ---- Sort1dExt ------------------------------------------------------------------------------------------
def get_from_sorted_mut(a, i, ctx=None):
    """Sort1dExt::get_from_sorted_mut (src/sort.rs:99-120)."""
    out = _ctx_for(a, ctx).get_many_from_sorted(a, [i])
    return out[0].item() if _is_torch(out) else out[0]


def get_many_from_sorted_mut(a, indexes, ctx=None):
    """Sort1dExt::get_many_from_sorted_mut (src/sort.rs:121-131): dict index -> value, ascending keys
    (the reference's IndexMap iterates in that order, src/sort.rs:34-35)."""
    idx = sorted(set(int(i) for i in indexes))
    if not idx:
        return {}
    vals = _ctx_for(a, ctx).get_many_from_sorted(a, idx)
    vals = vals.cpu().numpy() if _is_torch(vals) else vals
    return {i: v for i, v in zip(idx, vals)}


def partition_mut(a, pivot_index, ctx=None):
    """Sort1dExt::partition_mut (src/sort.rs:133-168): in place; returns the new index of the pivot value.  Elements
    before it are smaller, elements after it are greater than or equal (src/sort.rs:58-67); index and arrangement are
    the reference's, bit for bit.  ``pivot_index >= len(a)`` raises IndexError (the reference panics)."""
    return _ctx_for(a, ctx).partition(a, pivot_index)


# ---- multi-GPU helper: independent lanes shard with no collective ----------------------
def shard_lanes(a, axis, rank, world):
    """The block of lanes rank ``rank`` of ``world`` owns: a view of ``a`` split along the first
    non-``axis`` dimension that is long enough, contiguous blocks, remainder to the low ranks."""
    axis = int(axis)
    dims = [d for d in range(len(a.shape)) if d != axis]
    if not dims:
        raise ValueError("a single lane cannot be sharded by lanes; use quantiles_1d_sharded")
    d = builtins.max(dims, key=lambda k: a.shape[k])  # (`max` below is the reference's QuantileExt::max)
    n = a.shape[d]
    base, rem = divmod(n, world)
    start = rank * base + builtins.min(rank, rem)
    stop = start + base + (1 if rank < rem else 0)
    index = [slice(None)] * len(a.shape)
    index[d] = slice(start, stop)
    return a[tuple(index)], d, (start, stop)


# ---- QuantileExt: min / max / argmin / argmax ----------------------------------------
def _pattern(idx):
    """ndarray's D::Pattern: a bare usize for 1-D arrays, a tuple otherwise."""
    return idx[0] if len(idx) == 1 else idx


def _minmax(a, want_max, skipnan, ctx, want_index=True):
    return _ctx_for(a, ctx).minmax(a, want_max, skipnan, want_index)


def argmin(a, ctx=None):
    """QuantileExt::argmin (src/quantile/mod.rs:283-300)."""
    st, _, idx = _minmax(a, False, False, ctx)
    if st == NDS_EMPTY_INPUT:
        raise MinMaxEmptyInput()
    if st == NDS_UNDEFINED_ORDER:
        raise UndefinedOrder()
    return _pattern(idx)


def argmax(a, ctx=None):
    """QuantileExt::argmax (src/quantile/mod.rs:347-363)."""
    st, _, idx = _minmax(a, True, False, ctx)
    if st == NDS_EMPTY_INPUT:
        raise MinMaxEmptyInput()
    if st == NDS_UNDEFINED_ORDER:
        raise UndefinedOrder()
    return _pattern(idx)


def min(a, ctx=None):  # noqa: A001  (the reference's name)
    """QuantileExt::min (src/quantile/mod.rs:320-333)."""
    st, v, _ = _minmax(a, False, False, ctx)
    if st == NDS_EMPTY_INPUT:
        raise MinMaxEmptyInput()
    if st == NDS_UNDEFINED_ORDER:
        raise UndefinedOrder()
    return v


def max(a, ctx=None):  # noqa: A001
    """QuantileExt::max (src/quantile/mod.rs:384-397)."""
    st, v, _ = _minmax(a, True, False, ctx)
    if st == NDS_EMPTY_INPUT:
        raise MinMaxEmptyInput()
    if st == NDS_UNDEFINED_ORDER:
        raise UndefinedOrder()
    return v


def argmin_skipnan(a, ctx=None):
    """QuantileExt::argmin_skipnan (src/quantile/mod.rs:302-318): EmptyInput when no non-NaN value exists."""
    st, _, idx = _minmax(a, False, True, ctx)
    if st == NDS_EMPTY_INPUT:
        raise EmptyInput()
    return _pattern(idx)


def argmax_skipnan(a, ctx=None):
    """QuantileExt::argmax_skipnan (src/quantile/mod.rs:365-382)."""
    st, _, idx = _minmax(a, True, True, ctx)
    if st == NDS_EMPTY_INPUT:
        raise EmptyInput()
    return _pattern(idx)


def min_skipnan(a, ctx=None):
    """QuantileExt::min_skipnan (src/quantile/mod.rs:335-345): NaN when no non-NaN value exists."""
    return _minmax(a, False, True, ctx)[1]


def max_skipnan(a, ctx=None):
    """QuantileExt::max_skipnan (src/quantile/mod.rs:399-409)."""
    return _minmax(a, True, True, ctx, want_index=False)[1]


from .strategies import Auto, BinsBuildError, EquiSpaced, FreedmanDiaconis, Rice, Sqrt, Sturges  # noqa: E402,F401  
from .histogram import BinNotFound, Bins, Edges, Grid, GridBuilder, Histogram, histogram  # noqa: E402,F401
from . import synth  # noqa: E402,F401  (synthetic benchmark inputs)
