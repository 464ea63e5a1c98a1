"""A short run of the bracket select (2^27 f64, 99 percentiles) for profiling its small fixed-cost kernels with ncu:
  python tools/prof_small_kernels.py && ncu --set full --import-source on --clock-control none \
      -k regex:'brk_(plan|refine_decide|resolve|final|sort_s0|bucket|sample)' -c 12 -o gpurun_out/prof_small python tools/prof_small_kernels.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _ndstats_import  # noqa: E402

nds = _ndstats_import.load()
ctx = nds.Context(0)
n = int(os.environ.get("N", 1 << 27))
t = torch.empty(n, dtype=torch.float64, device="cuda")
nds.synth.fill(ctx, t, "f64_uniform", 1004)
qs = np.arange(1, 100) / 100.0
for _ in range(int(os.environ.get("CALLS", 2))):
    out = ctx.quantiles_1d_sharded(t, qs, "linear")
print("ok", ctx.last_path, float(out[49]))
