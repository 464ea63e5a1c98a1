"""Timing of calls that fall between the lane kernels and the grid-wide passes (device-resident, enqueued, wall clock
over 10 calls): python tools/bench_mid.py"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import _ndstats_import
nds = _ndstats_import.load()
ctx = nds.Context(0)
def bench(shape, dtype, qs, interp, label):
    t = torch.rand(shape, dtype=torch.float64, device="cuda").to(dtype) if dtype.is_floating_point else torch.randint(-10**6, 10**6, shape, device="cuda", dtype=dtype)
    for _ in range(3): ctx.quantiles_axis(t, len(shape) - 1, qs, interp, enqueue=True)
    ctx.synchronize(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): ctx.quantiles_axis(t, len(shape) - 1, qs, interp, enqueue=True)
    ctx.synchronize(); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"{label:44s} {ctx.last_path:14s} ms/call {dt * 1e3:8.4f}  GB/s {t.numel() * t.element_size() / dt / 1e9:8.1f}", flush=True)
pct = np.arange(1, 100) / 100.0
bench((1, 65536), torch.float64, pct, "linear", "smoke shape 1x65536 f64 x99")
bench((1000, 20000), torch.float32, [0.5], "midpoint", "1000x20000 f32 median")
bench((1000, 20000), torch.float32, pct, "linear", "1000x20000 f32 x99")
bench((4096, 65536), torch.float32, [0.25, 0.5, 0.75], "linear", "4096x65536 f32 quartiles (1 GiB)")
bench((2048, 65536), torch.float64, [0.5], "midpoint", "2048x65536 f64 median (1 GiB)")
bench((16, 4096), torch.float32, [0.5], "midpoint", "16x4096 f32 median")
bench((16384, 16384), torch.int64, [0.1, 0.5, 0.9], "lower", "16384x16384 i64 (2 GiB)")
bench((8192, 32768), torch.float32, [0.5], "nearest", "8192x32768 f32 median (1 GiB)")
bench((512, 524288), torch.float32, [0.5], "midpoint", "512x524288 f32 median (1 GiB, clusters of 8)")
bench((1024, 131072), torch.float64, [0.25, 0.5, 0.75], "linear", "1024x131072 f64 quartiles (1 GiB, clusters of 2)")
def bench_cols(shape, dtype, qs, interp, label):
    t = torch.rand(shape, dtype=torch.float64, device="cuda").to(dtype)
    for _ in range(3): ctx.quantiles_axis(t, 0, qs, interp, enqueue=True)
    ctx.synchronize(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): ctx.quantiles_axis(t, 0, qs, interp, enqueue=True)
    ctx.synchronize(); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"{label:44s} {ctx.last_path:22s} ms/call {dt * 1e3:8.4f}  GB/s {t.numel() * t.element_size() / dt / 1e9:8.1f}", flush=True)
bench_cols((65536, 4096), torch.float32, [0.5], "midpoint", "65536x4096 f32 column medians (1 GiB)")
bench_cols((500000, 512), torch.float32, [0.25, 0.5, 0.75], "linear", "500000x512 f32 column quartiles (1 GB)")
bench_cols((100000, 1000), torch.float64, [0.5], "nearest", "100000x1000 f64 column medians (0.8 GB)")
