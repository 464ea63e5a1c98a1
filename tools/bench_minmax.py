"""Device time of nds_minmax on a 1 GiB f32 / f64 array (CUDA events around K calls incl. their result read-back)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _ndstats_import
nds = _ndstats_import.load()
ctx = nds.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for dt in (torch.float32, torch.float64):
    t = torch.rand((1 << 30) // t.element_size() if False else (1 << 30) // torch.empty(0, dtype=dt).element_size(), device="cuda", dtype=dt)
    for _ in range(3): nds.argmin(t, ctx=ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 20
    e0.record(st)
    for _ in range(K): nds.argmin(t, ctx=ctx)
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    ok = int(nds.argmin(t, ctx=ctx)) == int(t.argmin()) and int(nds.argmax(t, ctx=ctx)) == int(t.argmax())
    print(f"argmin {dt}: {ms:.3f} ms/call  {t.numel() * t.element_size() / ms / 1e6:.0f} GB/s  ok={ok}")
