"""Multi-GPU check of nds_quantiles_1d_sharded (one process per GPU, launched by torch.distributed.run).

Every rank holds a different shard of ONE 1-D array; the result must equal the CPU oracle on the
concatenation (bit-exact; Linear <= 1 ulp, asserted 0).  Run on the GPU box:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded_nccl.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _ndstats_import  # noqa: E402
from kat_util import assert_bit_equal, ulp_distance  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nds = _ndstats_import.load()
    ctx = nds.Context(local)
    uid = [nds.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init_rank(world, rank, uid[0])
    from oracle import oracle

    qs99 = np.arange(1, 100) / 100.0
    cases = [
        (np.float64, 1 << 20, qs99, "linear", False),
        (np.float32, 300001, [0.0, 0.5, 1.0, 0.25], "midpoint", False),
        (np.float64, 70001, [0.01, 0.5, 0.99], "nearest", True),
        (np.uint32, 1 << 18, [0.1, 0.5, 0.9], "lower", False),
        (np.int64, 99999, [0.1, 0.5, 0.9], "higher", False),
        (np.float32, 5, [0.5], "linear", False),
        (np.float64, 1 << 24, list(np.arange(1, 10) / 10.0), "linear", False),   # the one-pass path, 9 brackets
        (np.float32, (1 << 23) + 77, [0.001, 0.5, 0.999], "midpoint", True),
        (np.int64, 1 << 22, [0.25, 0.75], "lower", False),
    ]
    paths = []
    for dtype, n_total, qs, interp, skipnan in cases:
        rng = np.random.default_rng(1234 + n_total)  # same full array on every rank, each keeps its slice
        if np.dtype(dtype).kind == "f":
            full = rng.standard_normal(n_total).astype(dtype)
            if n_total == 1 << 24:
                full.sort()  # sorted and sharded: every rank holds a different value range
            if skipnan:
                full[rng.random(n_total) < 0.2] = np.nan
        elif dtype == np.uint32:
            full = (rng.integers(0, 16, n_total).astype(np.uint64) * 0x11111111).astype(np.uint32)
        else:
            full = rng.integers(-2**40, 2**40, n_total).astype(dtype)
        # uneven shards, rank 1 may even be empty for the tiny case
        cuts = np.linspace(0, n_total, world + 1).astype(np.int64)
        if n_total == 5:
            cuts = np.array([0] + [5] * world)
        local_np = full[cuts[rank]:cuts[rank + 1]]
        shard = torch.from_numpy(local_np.copy()).cuda() if local_np.size else torch.empty(0, dtype=torch.from_numpy(full[:1]).dtype, device="cuda")
        got = ctx.quantiles_1d_sharded(shard, qs, interp, skipnan).cpu().numpy()
        st, want, _ = oracle.quantiles_axis(full, 0, qs, interp, skipnan, select_impl=1)
        assert st == 0
        if interp == "linear":
            assert ulp_distance(got, want) == 0
        else:
            assert_bit_equal(got, want, msg=f"rank {rank} {dtype} {n_total} {interp}")
        assert ctx.last_path in ("bracket_sharded", "long_radix_sharded", "bracket->long_radix"), ctx.last_path
        paths.append(ctx.last_path)
    # error paths are collective too
    try:
        ctx.quantiles_1d_sharded(torch.empty(0, dtype=torch.float64, device="cuda"), [0.5], "linear")
        raise SystemExit("expected EmptyInput")
    except nds.EmptyInput:
        pass
    dist.barrier()
    if rank == 0:
        print(f"sharded NCCL check ok on {world} GPUs; paths: {paths}; exchange: {ctx.comm_info()}")
    ctx.comm_destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
