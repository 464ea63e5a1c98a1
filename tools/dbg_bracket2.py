"""One seed of tools/dbg_bracket.py with NDS_BRK_DEBUG=2: compare the device's class counts with numpy."""
import os, re, subprocess, sys
import numpy as np
seed = sys.argv[1]
env = dict(os.environ, NDS_BRK_DEBUG="2")
code = f"""
import sys, os, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import _ndstats_import
from test_parity_gpu import _long_data
nds = _ndstats_import.load(); ctx = nds.Context(0); ctx.force_path('bracket')
a = _long_data(np.random.default_rng({seed}), 'normal', 300007, np.dtype('uint32'))
np.save('gpurun_out/dbg_data.npy', a)
print(ctx.quantiles_axis(a, 0, [0.25, 0.75, 0.999, 0.001], 'nearest'))
"""
r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
print(r.stdout); err = r.stderr
a = np.load("gpurun_out/dbg_data.npy")
rows = re.findall(r"\[bracket-b\] (\d+) lo=(\d+) hi=(\d+) tie=(\d+) cls_count=(\d+) in=(\d+)", err)
print([l for l in err.splitlines() if "bracket-lut" in l or "bracket]" in l])
los = [int(r[1]) for r in rows]; his = [int(r[2]) for r in rows]
j = np.searchsorted(np.array(los, dtype=np.uint64), a.astype(np.uint64), side="right")
for b, r in enumerate(rows):
    true_cls = int((j == b).sum()); true_in = int(((a >= los[b]) & (a <= his[b])).sum())
    print("b", b, "lo", los[b], "hi", his[b], "dev cls", r[4], "true cls", true_cls, "dev in", r[5], "true in", true_in)
m = re.search(r"\[bracket-b\] (\d+) \(above\) cls_count=(\d+)", err)
print("above dev", m.group(2), "true", int((j == len(rows)).sum()))
