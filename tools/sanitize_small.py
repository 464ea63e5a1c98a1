"""Small invocations of every kernel path for compute-sanitizer (memcheck): run as
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _ndstats_import
nds = _ndstats_import.load()
ctx = nds.Context(0)
rng = np.random.default_rng(0)
a = rng.standard_normal((70, 512)).astype(np.float32)
ctx.quantiles_axis(a, 1, [0.5, 0.9], "linear"); print(ctx.last_path)
ctx.quantiles_axis(a[:, :300], 1, [0.5], "midpoint"); print(ctx.last_path)
b = rng.standard_normal((1030, 40)).astype(np.float32); b[rng.random(b.shape) < 0.1] = np.nan
ctx.quantiles_axis(b, 0, [0.5], "nearest", skipnan=True); print(ctx.last_path)
ctx.quantiles_axis(rng.standard_normal((520, 33)), 0, [0.25], "lower"); print(ctx.last_path)
for dt in (np.float64, np.float32, np.int64, np.uint32):
    c = (rng.standard_normal(70001) * 1000).astype(dt)
    ctx.force_path("bracket"); ctx.quantiles_axis(c, 0, [0.1, 0.5, 0.9, 0.0, 1.0], "linear" if dt in (np.float64, np.float32) else "lower"); print(ctx.last_path)
    ctx.force_path("long_radix"); ctx.quantiles_axis(c, 0, [0.1, 0.5, 0.9], "lower"); print(ctx.last_path)
    ctx.force_path("auto")
d = rng.standard_normal(66000); d[rng.random(d.size) < 0.2] = np.nan
ctx.force_path("bracket"); ctx.quantiles_axis(d, 0, [0.5], "linear", skipnan=True); ctx.force_path("auto")
ctx.quantiles_axis(rng.integers(0, 9, size=(3, 5, 7)).astype(np.int16), 1, [0.3], "higher"); print(ctx.last_path)
print(nds.argmin(a, ctx=ctx), nds.max_skipnan(b, ctx=ctx), nds.argmax(a[::-1, ::2], ctx=ctx))
print(nds.get_many_from_sorted_mut(rng.integers(0, 100, 500), [0, 10, 499], ctx=ctx))
print("sanitize_small done")
