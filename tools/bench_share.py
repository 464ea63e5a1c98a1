"""One rank's share of BASELINE config 4 at 8 GPUs (2^28 f64, 99 percentiles, Linear) on ONE GPU: the sample size, the
candidate share and the 148 pieces per bracket are those of a rank of the 8-GPU run, so the stages that do not shrink with
the rank count (plan, deciding kernels, round 0 over small pieces, final sorts) can be measured without an 8-GPU box.
NDS_BRK_TIMING=1 python tools/bench_share.py [log2 n]  -> per-stage CUDA-event times on stderr + ms per enqueued call."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import _ndstats_import
nds = _ndstats_import.load()
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 28
ctx = nds.Context(0)
t = torch.rand(1 << lg, dtype=torch.float64, device="cuda")
qs = np.arange(1, 100) / 100.0
out = None
for _ in range(3):
    out = ctx.quantiles_axis(t, 0, qs, "linear")            # synchronous: the stage timer reports
want = torch.quantile(t[: 1 << 24], torch.tensor(qs, device="cuda"))  # (sanity only: same distribution)
for _ in range(3): ctx.quantiles_axis(t, 0, qs, "linear", enqueue=True)
ctx.synchronize(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): ctx.quantiles_axis(t, 0, qs, "linear", enqueue=True)
ctx.synchronize(); torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
srt = torch.sort(t).values
n = t.numel()
x = qs * (n - 1); lo = np.floor(x).astype(np.int64); hi = np.ceil(x).astype(np.int64); fr = x - np.trunc(x)
l = srt[torch.from_numpy(lo).cuda()].cpu().numpy(); h = srt[torch.from_numpy(hi).cuda()].cpu().numpy()
ref = l + fr * (h - l)
got = out.cpu().numpy() if hasattr(out, "cpu") else out
print(f"n=2^{lg} path={ctx.last_path} ms/call {dt * 1e3:.4f} bit-exact {np.array_equal(ref.view(np.uint64), got.view(np.uint64))}")
