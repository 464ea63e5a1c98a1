"""ctypes declarations for libndstats_b200.so — one prototype per symbol of include/ndstats_b200.h."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))

(NDS_OK, NDS_EMPTY_INPUT, NDS_INVALID_QUANTILE, NDS_NAN_IN_INPUT, NDS_UNSUPPORTED, NDS_CUDA_ERROR, NDS_NCCL_ERROR,
 NDS_BAD_ARGUMENT, NDS_CAST_OVERFLOW, NDS_INDEX_OUT_OF_BOUNDS, NDS_UNDEFINED_ORDER) = range(11)

_vp, _int, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
_psz = ctypes.POINTER(ctypes.c_size_t)
_pss = ctypes.POINTER(ctypes.c_ssize_t)
_pd = ctypes.POINTER(ctypes.c_double)

# symbol -> (restype, argtypes); must list every function the header declares (tests check this)
PROTOTYPES = {
    "nds_abi_version": (_int, []),
    "nds_status_string": (ctypes.c_char_p, [_int]),
    "nds_dtype_size": (_sz, [_int]),
    "nds_ctx_create": (_int, [_int, ctypes.POINTER(_vp)]),
    "nds_ctx_destroy": (None, [_vp]),
    "nds_ctx_set_stream": (_int, [_vp, _vp]),
    "nds_ctx_get_stream": (_vp, [_vp]),
    "nds_ctx_synchronize": (_int, [_vp]),
    "nds_ctx_last_error": (ctypes.c_char_p, [_vp]),
    "nds_ctx_kernel_launches": (ctypes.c_uint64, [_vp]),
    "nds_ctx_last_path": (ctypes.c_char_p, [_vp]),
    "nds_ctx_force_path": (_int, [_vp, ctypes.c_char_p]),
    "nds_validate_quantile_args": (_int, [_int, _psz, _int, _pd, _sz, _int, _psz, _psz]),
    "nds_quantiles_axis": (_int, [_vp, _int, _vp, _int, _psz, _pss, _int, _pd, _sz, _int, _int, _vp, _psz]),
    "nds_quantiles_axis_enqueue": (_int, [_vp, _int, _vp, _int, _psz, _pss, _int, _pd, _sz, _int, _int, _vp, _psz]),
    "nds_quantiles_axis_masked": (_int, [_vp, _int, _vp, _vp, _int, _psz, _pss, _int, _pd, _sz, _int, _vp, _vp, _psz]),
    "nds_get_many_from_sorted": (_int, [_vp, _int, _vp, _sz, ctypes.c_ssize_t, _psz, _sz, _vp]),
    "nds_partition": (_int, [_vp, _int, _vp, _sz, ctypes.c_ssize_t, _sz, _psz]),
    "nds_minmax": (_int, [_vp, _int, _vp, _int, _psz, _pss, _int, _int, _vp, _psz]),
    "nds_histogram": (_int, [_vp, _int, _vp, _sz, _sz, ctypes.c_ssize_t, ctypes.c_ssize_t, _vp, _psz, _vp]),
    "nds_comm_get_unique_id": (_int, [_vp]),
    "nds_comm_init_rank": (_int, [_vp, _int, _int, _vp]),
    "nds_comm_destroy": (_int, [_vp]),
    "nds_comm_stats": (_int, [_vp, ctypes.POINTER(ctypes.c_uint64)]),
    "nds_quantiles_1d_sharded": (_int, [_vp, _int, _vp, _sz, _pd, _sz, _int, _int, _vp, _psz]),
    "nds_synth_fill": (_int, [_vp, _int, _vp, _sz, ctypes.c_uint64, ctypes.c_uint64]),
}

_lib = None


def library_path():
    # NDS_LIB_PATH: A/B-test an alternative build of the same ABI (tuning only)
    return os.environ.get("NDS_LIB_PATH") or os.path.join(_HERE, "libndstats_b200.so")


def load():
    """Load the CUDA library.  There is deliberately no fallback: a missing build is an error."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C ndarray-stats_b200/csrc`). ndarray_stats_b200 has no CPU fallback.")
    lib = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here means the .so is stale against the header
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
