"""Stress the bracket path on one configuration over many seeds; print the first mismatches (GPU box only)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import _ndstats_import
from oracle import oracle
from test_parity_gpu import _long_data
nds = _ndstats_import.load()
ctx = nds.Context(0)
ctx.force_path("bracket")
dtype = np.dtype(sys.argv[1]); kind = sys.argv[2]; n = int(sys.argv[3]); interp = sys.argv[4]
qs = [float(x) for x in sys.argv[5].split(",")]
bad = 0
for seed in range(int(sys.argv[6]) if len(sys.argv) > 6 else 40):
    rng = np.random.default_rng(seed)
    a = _long_data(rng, kind, n, dtype)
    got = ctx.quantiles_axis(a, 0, qs, interp)
    st, want, _ = oracle.quantiles_axis(a, 0, qs, interp, select_impl=1)
    if os.environ.get("NDS_BRK_SEED") and int(os.environ["NDS_BRK_SEED"]) == seed:
        np.save("gpurun_out/dbg_data.npy", a)
    if not np.array_equal(got, want):
        bad += 1
        srt = np.sort(a)
        print("seed", seed, "path", ctx.last_path, "got", got, "want", want,
              "rank(got)", [int(np.searchsorted(srt, g)) for g in got], "rank(want)", [int(np.searchsorted(srt, w)) for w in want])
print("bad", bad)
