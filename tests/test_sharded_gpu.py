"""GPU tests of the multi-GPU entry point nds_quantiles_1d_sharded (Quantile1dExt::quantiles_mut over a 1-D array
sharded over ranks, src/quantile/mod.rs:623-633).

* on ONE rank (no communicator needed) the entry must agree with the oracle on every case of
  tools/check_sharded_nccl.py -- this exercises `bracket_sharded` / `long_radix_sharded` on any GPU box;
* with >= 2 GPUs the same cases run as one process per GPU (torch.distributed.run), every rank holding a different
  shard, against the oracle on the concatenation.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from kat_util import assert_bit_equal, ulp_distance

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

QS99 = np.arange(1, 100) / 100.0
CASES = [
    (np.float64, 1 << 20, QS99, "linear", False),
    (np.float32, 300001, [0.0, 0.5, 1.0, 0.25], "midpoint", False),
    (np.float64, 70001, [0.01, 0.5, 0.99], "nearest", True),
    (np.uint32, 1 << 18, [0.1, 0.5, 0.9], "lower", False),
    (np.int64, 99999, [0.1, 0.5, 0.9], "higher", False),
    (np.float32, 5, [0.5], "linear", False),
    (np.float64, 1 << 24, list(np.arange(1, 10) / 10.0), "linear", False),
    (np.float32, (1 << 23) + 77, [0.001, 0.5, 0.999], "midpoint", True),
    (np.int64, 1 << 22, [0.25, 0.75], "lower", False),
]


def make_case(dtype, n_total, skipnan):
    rng = np.random.default_rng(1234 + n_total)
    if np.dtype(dtype).kind == "f":
        full = rng.standard_normal(n_total).astype(dtype)
        if n_total == 1 << 24:
            full.sort()
        if skipnan:
            full[rng.random(n_total) < 0.2] = np.nan
    elif dtype == np.uint32:
        full = (rng.integers(0, 16, n_total).astype(np.uint64) * 0x11111111).astype(np.uint32)
    else:
        full = rng.integers(-2**40, 2**40, n_total).astype(dtype)
    return full


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{np.dtype(c[0]).name}-{c[1]}-{c[3]}")
def test_sharded_entry_one_rank(nds, gpu_ctx, oracle, case):
    import torch
    dtype, n_total, qs, interp, skipnan = case
    full = make_case(dtype, n_total, skipnan)
    shard = torch.from_numpy(full.copy()).cuda()
    got = gpu_ctx.quantiles_1d_sharded(shard, qs, interp, skipnan).cpu().numpy()
    assert gpu_ctx.last_path in ("bracket_sharded", "long_radix_sharded", "bracket->long_radix"), gpu_ctx.last_path
    st, want, _ = oracle.quantiles_axis(full, 0, qs, interp, skipnan, select_impl=1)
    assert st == 0
    if interp == "linear":
        assert ulp_distance(got, want) == 0
    else:
        assert_bit_equal(got, want, msg=f"{dtype} {n_total} {interp}")


def test_sharded_entry_one_rank_errors(nds, gpu_ctx):
    import torch
    with pytest.raises(nds.EmptyInput):
        gpu_ctx.quantiles_1d_sharded(torch.empty(0, dtype=torch.float64, device="cuda"), [0.5], "linear")
    with pytest.raises(nds.InvalidQuantile):
        gpu_ctx.quantiles_1d_sharded(torch.zeros(8, dtype=torch.float64, device="cuda"), [0.5, 1.5], "linear")


def test_sharded_entry_host_shard(nds, gpu_ctx, oracle):
    """a host (numpy) shard is staged like any other host input"""
    full = make_case(np.float64, 1 << 20, False)
    got = gpu_ctx.quantiles_1d_sharded(full, [0.1, 0.5, 0.9], "linear")
    st, want, _ = oracle.quantiles_axis(full, 0, [0.1, 0.5, 0.9], "linear", select_impl=1)
    assert ulp_distance(got, want) == 0


def test_sharded_multi_process():
    """One process per GPU, every rank a different shard (tools/check_sharded_nccl.py); needs >= 2 GPUs."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else (4 if n < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "check_sharded_nccl.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "sharded NCCL check ok" in r.stdout
