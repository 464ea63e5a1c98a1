"""CPU tests of the product's boundary: the library loads, exports every symbol the header declares,
and its host-only argument logic follows the reference's error precedence.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(nds):
    return nds._cabi.load()


def header_functions():
    text = open(os.path.join(ROOT, "include", "ndstats_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nds_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(nds, lib):
    names = header_functions()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(names) == set(nds._cabi.PROTOTYPES), "ctypes table and header disagree"


def test_no_cpu_fallback(nds):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nds.NdsError):
        nds.Context(0)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "ndarray-stats_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".rs", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"


def test_status_strings_and_sizes(lib):
    assert lib.nds_abi_version() == 1
    assert lib.nds_status_string(1) == b"Empty input."  # src/errors.rs:128
    assert b"is not between 0. and 1. (inclusive)." in lib.nds_status_string(2)  # src/errors.rs:129-131
    assert [lib.nds_dtype_size(i) for i in range(10)] == [1, 2, 4, 8, 1, 2, 4, 8, 4, 8]


def _validate(lib, shape, axis, qs, interp=0):
    shape_c = (ctypes.c_size_t * len(shape))(*shape)
    qs = np.asarray(qs, dtype=np.float64)
    bad, elems = ctypes.c_size_t(99), ctypes.c_size_t(99)
    st = lib.nds_validate_quantile_args(len(shape), shape_c, axis, qs.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                        qs.size, interp, ctypes.byref(bad), ctypes.byref(elems))
    return st, bad.value, elems.value


def test_validation_precedence_matches_reference(lib, oracle):
    # src/quantile/mod.rs:440-455: bad q first (first offender, qs order), then empty axis, then Ok(empty)
    assert _validate(lib, [5, 0], 1, [0.5, 1.5, -1.0])[:2] == (2, 1)
    assert _validate(lib, [5, 0], 1, [float("nan")])[:2] == (2, 0)
    assert _validate(lib, [5, 0], 1, [0.5])[0] == 1            # tests/quantile.rs:179-185
    assert _validate(lib, [5, 0], 0, [0.5]) == (0, 99, 0)      # tests/quantile.rs:188-192
    assert _validate(lib, [3, 4], 1, []) == (0, 99, 0)
    assert _validate(lib, [3, 4], 2, [0.5])[0] == 7            # axis out of bounds: the reference panics
    assert _validate(lib, [3, 4], 1, [0.0, 1.0])[::2] == (0, 6)
    # the oracle agrees on each of these
    a = np.zeros((5, 0))
    assert oracle.quantiles_axis(a, 1, [0.5, 1.5, -1.0], "lower")[0::2] == (2, 1)
    assert oracle.quantiles_axis(a, 1, [0.5], "lower")[0] == 1
    assert oracle.quantiles_axis(a, 0, [0.5], "lower")[0] == 0


def test_shard_lanes_partitions_every_lane_once(nds):
    a = np.arange(7 * 10).reshape(7, 10)
    for world in (1, 2, 3, 4, 8):
        parts = [nds.shard_lanes(a, 1, r, world) for r in range(world)]
        assert all(p[1] == 0 for p in parts)
        rows = np.concatenate([p[0] for p in parts], axis=0)
        assert np.array_equal(rows, a)
        parts = [nds.shard_lanes(a, 0, r, world) for r in range(world)]
        cols = np.concatenate([p[0] for p in parts], axis=1)
        assert np.array_equal(cols, a)


def test_cpp_mirror_selftest_host_only():
    """include/ndstats_b200.hpp (the C++ host mirror of the trait surface): compiles against the C ABI and passes its
    host-only checks -- validation order and error strings of the reference, and that without a device every compute call
    fails loudly (no CPU fallback).  The device part of the self-test runs when a GPU is visible; here it is masked out."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cpp = os.path.join(root, "ndarray-stats_b200", "cpp")
    subprocess.run(["make", "-C", cpp, "-s", "selftest"], check=True)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([os.path.join(cpp, "selftest")], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpp mirror self-test: ok (host-only checks" in r.stdout
