"""Host-side logic of the multi-GPU paths on CPU (gloo, world_size 2): lane sharding needs no collective and
tiles the array exactly; the sharded 1-D path's exchange step (sum of per-rank digit histograms / counts) is an
all-reduce whose result every rank uses to take the same decision.  The CUDA kernels are not involved here."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch
    import torch.distributed as dist
    import _ndstats_import
    from oracle import oracle
    nds = _ndstats_import.load()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)  # the same array on every rank
    # ---- lanes: every rank takes its block, results concatenate to the full answer ---------------------
    a = rng.standard_normal((37, 129)).astype(np.float32)
    a[rng.random(a.shape) < 0.1] = np.nan
    for axis in (0, 1):
        mine, dim, (lo, hi) = nds.shard_lanes(a, axis, rank, world)
        st, part, _ = oracle.quantiles_axis(mine, axis, [0.5], "nearest", skipnan=True)
        assert st == 0
        spans = [None] * world
        dist.all_gather_object(spans, (dim, lo, hi, part))
        assert [s[1] for s in spans] == [0] + [s[2] for s in spans[:-1]] and spans[-1][2] == a.shape[dim]
        glued = np.concatenate([s[3] for s in spans], axis=dim)
        st, full, _ = oracle.quantiles_axis(a, axis, [0.5], "nearest", skipnan=True)
        assert np.array_equal(glued.view(np.uint32), full.view(np.uint32))
    # ---- one 1-D array over ranks: MSB radix select whose per-pass histograms are summed over ranks ----
    x = rng.integers(0, 2**32, size=100003, dtype=np.uint64).astype(np.uint32)
    cuts = np.linspace(0, x.size, world + 1).astype(int)
    shard = x[cuts[rank]:cuts[rank + 1]]
    n_total = torch.tensor([shard.size], dtype=torch.int64)
    dist.all_reduce(n_total)
    ranks = [0, int(0.37 * (int(n_total) - 1)), int(n_total) - 1]
    got = []
    for r in ranks:
        prefix, mask, rem = 0, 0, r
        for shift in (24, 16, 8, 0):
            sel = shard[(shard & np.uint32(mask)) == np.uint32(prefix)]
            hist = torch.from_numpy(np.bincount((sel >> np.uint32(shift)) & np.uint32(0xFF), minlength=256).astype(np.int64))
            dist.all_reduce(hist)  # the exchange step: identical on every rank afterwards
            cum = np.cumsum(hist.numpy())
            digit = int(np.searchsorted(cum, rem, side="right"))
            rem -= int(cum[digit - 1]) if digit else 0
            prefix |= digit << shift
            mask |= 0xFF << shift
        got.append(prefix)
    assert got == [int(v) for v in np.sort(x)[ranks]]
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_sharding_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
