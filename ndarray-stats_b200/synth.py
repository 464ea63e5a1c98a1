"""Synthetic benchmark inputs: counter-based splitmix64 streams, identical on host and device.

``host(kind, n, seed, first_index)`` restates in numpy what ``nds_synth_fill`` (csrc/nds_synth.cu) writes on the
device, so that bench.py's GPU arm and its CPU reference arm see the same bytes.
"""
import ctypes

import numpy as np

F64_UNIFORM, F32_UNIFORM, F32_UNIFORM_NAN10, U32_16VALUES, U32_16VALUES_ADV, I64_16VALUES = range(6)
KINDS = {
    "f64_uniform": F64_UNIFORM, "f32_uniform": F32_UNIFORM, "f32_uniform_nan10": F32_UNIFORM_NAN10,
    "u32_16values": U32_16VALUES, "u32_16values_adv": U32_16VALUES_ADV, "i64_16values": I64_16VALUES,
}
DTYPES = {F64_UNIFORM: np.float64, F32_UNIFORM: np.float32, F32_UNIFORM_NAN10: np.float32, U32_16VALUES: np.uint32,
          U32_16VALUES_ADV: np.uint32, I64_16VALUES: np.int64}


def splitmix64(x):
    """mix(x) on a uint64 array (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def host(kind, n, seed, first_index=0):
    """The n elements [first_index, first_index + n) of stream (kind, seed) as a numpy array."""
    return host_at(kind, np.arange(n, dtype=np.uint64) + np.uint64(first_index), seed)


def host_at(kind, index, seed):
    """The elements of stream (kind, seed) at the given element indexes (any shape)."""
    kind = KINDS.get(kind, kind)
    with np.errstate(over="ignore"):
        idx = np.asarray(index).astype(np.uint64) + np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
        z = splitmix64(idx)
        if kind == F64_UNIFORM:
            return (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
        if kind in (F32_UNIFORM, F32_UNIFORM_NAN10):
            out = (z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)
            if kind == F32_UNIFORM_NAN10:
                nan = splitmix64(idx + np.uint64(0x9E37)) % np.uint64(10) == 0
                out[nan] = np.uint32(0x7FC00000).view(np.float32)
            return out
        if kind == U32_16VALUES:
            return ((z & np.uint64(15)) * np.uint64(0x11111111)).astype(np.uint32)
        if kind == U32_16VALUES_ADV:
            return (np.uint64(0x80000000) + (z & np.uint64(15))).astype(np.uint32)
        if kind == I64_16VALUES:
            return ((z & np.uint64(15)).astype(np.int64) - 8) * (np.int64(1) << np.int64(40))
    raise ValueError(f"unknown synthetic stream {kind!r}")


def fill(ctx, tensor, kind, seed, first_index=0):
    """``nds_synth_fill`` into a CUDA tensor (any shape, C-contiguous), on the ctx stream; returns the tensor."""
    kind = KINDS.get(kind, kind)
    if not tensor.is_contiguous():
        raise ValueError("synth.fill needs a contiguous tensor")
    if tensor.element_size() != np.dtype(DTYPES[kind]).itemsize:
        raise TypeError("tensor element size does not match the stream's element type")
    st = ctx._lib.nds_synth_fill(ctx._h, int(kind), ctypes.c_void_p(tensor.data_ptr()), tensor.numel(),
                                 ctypes.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), ctypes.c_uint64(first_index))
    ctx._raise(st)
    return tensor
