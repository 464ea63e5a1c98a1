#!/usr/bin/env python
"""Benchmark of the bulk quantile path (BASELINE.json metric: quantile input GB/s & elems/s, % of the HBM
roofline, next to the host CPU reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4|c5] [--impl reference]

One "step" is one pass of the hot path over one batch of synthetic input.  Default workload (the one the
metric is quoted on, BASELINE.json configs[1]):  c2 = 65536 x 4096 f32, row-wise median
(`quantile_axis_mut(Axis(1), 0.5, &Midpoint)`).  With N > 1 (launched by torch.distributed.run, one process
per GPU) the lanes are sharded: every rank owns its own 65536 x 4096 block, no data-path collective
(`scaling: weak`); the 1-D workload c4 instead shards ONE array and combines digit histograms with NCCL.

`value` times K enqueued calls on device-resident inputs with CUDA events on the launching stream (inputs
>= 1 GiB, i.e. larger than the 126 MB L2, so every step streams from HBM); `e2e` times the same K calls
through the public host-array API (pinned host buffers, H2D + D2H inside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, dtype, shape, axis, qs, interp, skipnan)
    "c1": ("1024x1024 f64 quantiles_axis_mut(Axis(1), [.25,.5,.75], Linear)", "float64", (1024, 1024), 1,
           [0.25, 0.5, 0.75], "linear", False),
    "c2": ("65536x4096 f32 quantile_axis_mut(Axis(1), 0.5, Midpoint)", "float32", (65536, 4096), 1, [0.5],
           "midpoint", False),
    "c3": ("8192x65536 f32 quantile_axis_skipnan_mut(Axis(0), 0.5, Nearest), 10% NaN", "float32", (8192, 65536), 0,
           [0.5], "nearest", True),
    "c4": ("1-D f64 2^31 quantiles_mut(k/100, k=1..99, Linear)", "float64", (1 << 31,), 0,
           [k / 100.0 for k in range(1, 100)], "linear", False),
    "c5": ("1-D u32 2^30, 16 distinct values, quantiles_mut([.1,.5,.9], Lower)", "uint32", (1 << 30,), 0,
           [0.1, 0.5, 0.9], "lower", False),
}
SEEDS = {"c1": 42, "c2": 1002, "c3": 1003, "c4": 1004, "c5": 1005}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_device_input(torch, name, rank, shape_override=None):
    desc, dtype, shape, axis, qs, interp, skipnan = WORKLOADS[name]
    shape = shape_override or shape
    g = torch.Generator(device="cuda").manual_seed(SEEDS[name] + 7919 * rank)
    if name == "c5":
        vals = torch.arange(16, device="cuda", dtype=torch.int64) * 0x11111111
        idx = torch.randint(0, 16, shape, generator=g, device="cuda")
        t = vals[idx].to(torch.uint32) if hasattr(torch, "uint32") else vals[idx].to(torch.int64)
        return t
    tdt = {"float32": torch.float32, "float64": torch.float64}[dtype]
    t = torch.rand(shape, generator=g, device="cuda", dtype=tdt)
    if name == "c3":  # 10 % NaN
        mask = torch.rand(shape, generator=g, device="cuda", dtype=torch.float32) < 0.1
        t[mask] = float("nan")
        del mask
    return t


def parity_sample(nds, oracle, name, t, out, nlanes_check=128, nds_ctx_call=None):
    """Compare a bounded sample of the timed run's result with the oracle on the same device-generated data."""
    desc, dtype, shape, axis, qs, interp, skipnan = WORKLOADS[name]
    if t.dim() == 1:
        # size-independent rank property on the full lane (the oracle cannot hold it): for a few requested
        # quantiles the Lower order statistic v satisfies count(x < v) <= rank < count(x <= v)
        import torch
        n = t.numel()
        picks = [qs[0], qs[len(qs) // 2], qs[-1]]
        low = nds_ctx_call(t, picks)
        ok = True
        if t.dtype not in (torch.float32, torch.float64, torch.int32, torch.int64):
            t = t.to(torch.int64)  # torch has no ordered compare for the unsigned types
        for q, v in zip(picks, low.tolist()):
            rank = int(np.floor(q * (n - 1)))
            vv = torch.tensor(v, dtype=t.dtype, device=t.device) if t.dtype.is_floating_point else int(v)
            lt = int((t < vv).sum().item())
            le = int((t <= vv).sum().item())
            ok = ok and (lt <= rank < le)
        return {"ranks_checked": len(picks), "ok": bool(ok)}
    lanes = t.shape[1 - axis]
    pick = np.linspace(0, lanes - 1, num=min(nlanes_check, lanes)).astype(np.int64)
    import torch
    idx = torch.from_numpy(pick).cuda()
    sub = t.index_select(1 - axis, idx).cpu().numpy()
    got = out.index_select(1 - axis, idx).cpu().numpy()
    st, want, _ = oracle.quantiles_axis(sub, axis, qs, interp, skipnan, select_impl=0)
    if st != 0:
        return {"lanes": int(pick.size), "ok": False, "why": f"oracle status {st}"}
    ut = {4: np.uint32, 8: np.uint64}[got.dtype.itemsize]
    g, w = got.copy(), want.copy()
    if g.dtype.kind == "f":
        g[g == 0] = 0
        w[w == 0] = 0
    ok = bool(np.array_equal(np.ascontiguousarray(g).view(ut), np.ascontiguousarray(w).view(ut)))
    return {"lanes": int(pick.size), "ok": ok}


def cpu_reference_run(name, threads, budget_s, shape_lanes=None):
    """Time the oracle (C restatement of the reference algorithm; the Rust crate cannot be built here) on a
    bounded sample of the workload.  Returns (GB/s, elems/s, sample description, seconds)."""
    from oracle import oracle
    desc, dtype, shape, axis, qs, interp, skipnan = WORKLOADS[name]
    rng = np.random.default_rng(SEEDS[name])
    dt = np.dtype(dtype)
    if len(shape) == 2:
        lane_len = shape[axis]
        nl = shape_lanes or max(threads * 64, 512)
        sample_shape = (nl, lane_len) if axis == 1 else (lane_len, nl)
    else:
        # c5: the faithful quickselect is quadratic on equal runs; 2^16 keys keep the sample to seconds
        sample_shape = (min(shape[0], 1 << 16 if name == "c5" else 1 << 24),)
    if dt.kind == "f":
        a = rng.random(sample_shape, dtype=dt.type if dt == np.float32 else np.float64).astype(dt, copy=False)
        if name == "c3":
            a[rng.random(sample_shape) < 0.1] = np.nan
    else:
        a = (rng.integers(0, 16, size=sample_shape).astype(np.uint64) * 0x11111111).astype(dt)
    elems = a.size
    t_total, reps = 0.0, 0
    while t_total < budget_s and reps < 200:
        work = a.copy()
        t0 = time.perf_counter()
        st, _, _ = oracle.quantiles_axis(work, axis, qs, interp, skipnan, select_impl=0, nthreads=threads, mutate=True)
        t_total += time.perf_counter() - t0
        reps += 1
        assert st == 0, f"oracle status {st}"
    sec = t_total / reps
    sample = f"{'x'.join(map(str, sample_shape))} {dtype} of the {name} workload, {reps} repetitions, {threads} thread(s)"
    if name == "c5":
        sample += "; the reference quickselect is quadratic on equal runs: 2^24 keys of this workload ran at 3.97e4 elems/s (round 1)"
    return elems * dt.itemsize / sec / 1e9, elems / sec, sample, t_total


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle
    # all host threads the process may use: torchrun exports OMP_NUM_THREADS=1, which is not the machine
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    threads = max(threads, oracle.max_threads())
    name = args.workload
    desc = WORKLOADS[name][0]
    # `--steps K --warmup W`: each step is one bounded sample; the whole run stays within a few minutes
    per_step_budget = min(8.0, 120.0 / max(1, args.steps + args.warmup))
    for _ in range(args.warmup):
        cpu_reference_run(name, threads, per_step_budget / 4)
    vals, samples, secs = [], None, 0.0
    for _ in range(args.steps):
        gbs, eps, samples, s = cpu_reference_run(name, threads, per_step_budget)
        vals.append(gbs)
        secs += s
    value = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": "quantile input GB/s", "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": {"float32": "f32", "float64": "f64", "uint32": "u32"}[WORKLOADS[name][1]],
        "data": "synthetic",
        "config": {"workload": f"{name}: {desc}", "note": "reference CPU algorithm (C restatement; the Rust crate cannot be built in this image), lanes split over all host threads"},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": threads, "kind": "port", "sample": samples},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--scale-down", type=int, default=1, help="divide the lane count (debug only; marks the line invalid)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import _ndstats_import
    nds = _ndstats_import.load()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    name = args.workload
    desc, dtype, shape, axis, qs, interp, skipnan = WORKLOADS[name]
    one_d = len(shape) == 1
    if args.scale_down > 1:
        shape = tuple(s // args.scale_down if i == (1 - axis if not one_d else 0) else s for i, s in enumerate(shape))
    esize = np.dtype(dtype).itemsize

    ctx = nds.Context(local_rank)
    stream = torch.cuda.Stream()          # a real stream: the CUDA events below must sit on the launching stream
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    sharded_1d = one_d and distributed
    if one_d:
        # strong scaling: ONE array split over the ranks (the exchange step is the histogram allreduce)
        n_total = shape[0]
        n_local = n_total // world + (1 if rank < n_total % world else 0)
        t = make_device_input(torch, name, rank, (n_local,))
        scaling = "strong"
        if sharded_1d:
            uid = [nds.Context.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            ctx.comm_init_rank(world, rank, uid[0])
        units_total = n_total
    else:
        t = make_device_input(torch, name, rank, shape)
        scaling = "weak"
        units_total = int(np.prod(shape)) * world
    out_shape = list(t.shape)
    out_shape[axis] = len(qs)
    out = torch.empty(out_shape, dtype=t.dtype, device="cuda")

    def step_device():
        if sharded_1d:
            ctx.quantiles_1d_sharded(t, qs, interp, skipnan)  # collective, synchronous
        else:
            ctx.quantiles_axis(t, axis, qs, interp, skipnan=skipnan, enqueue=True, out=out)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    ctx.synchronize()
    path = ctx.last_path

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = ctx.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    ctx.synchronize()
    barrier()
    launches = ctx.kernel_launches - launches0
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    if distributed:
        tt = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ms_per_step = ms_total / args.steps
    value = units_total * esize / (ms_per_step * 1e-3) / 1e9
    elems_per_s = units_total / (ms_per_step * 1e-3)

    # roofline of the dominant kernel: algorithmic bytes (every input element read once + the result)
    # over the per-launch duration measured above with CUDA events on the launching stream.
    peak, peak_src = peaks()
    per_gpu_bytes = t.numel() * esize + out.numel() * esize
    achieved = per_gpu_bytes / (ms_per_step * 1e-3) / 1e9
    traffic, dominant = None, path
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, one ncu --set full capture
        with open(tpath) as f:
            rec = json.load(f).get(name)
        if rec:
            traffic, dominant = rec["dram_bytes_per_launch"], rec["kernel"]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": dominant, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": per_gpu_bytes}

    # parity of the timed result on a bounded sample of lanes (checker only)
    parity = None
    if rank == 0 and not sharded_1d:
        from oracle import oracle
        parity = parity_sample(nds, oracle, name, t, out,
                               nds_ctx_call=lambda arr, picks: ctx.quantiles_axis(arr, 0, picks, "lower", skipnan=skipnan))

    # ---- e2e: the public host-array API, pinned host buffers, copies inside the timed region ---------
    e2e = None
    if not args.no_e2e and not sharded_1d:
        host_t = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host_t.copy_(t)
        host_np = host_t.numpy()
        host_out = torch.empty(out_shape, dtype=t.dtype, pin_memory=True).numpy()
        e2e_steps = max(2, min(args.steps, 5))
        ctx.quantiles_axis(host_np, axis, qs, interp, skipnan=skipnan, out=host_out)  # warm the arena
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.quantiles_axis(host_np, axis, qs, interp, skipnan=skipnan, out=host_out)
        barrier()
        dt = (time.perf_counter() - t0) / e2e_steps
        if distributed:
            tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        same = bool(torch.equal(torch.from_numpy(host_out).cuda().view(torch.uint8), out.view(torch.uint8))) \
            if not one_d else None
        e2e = {"value": units_total * esize / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": t.numel() * esize * world,
               "d2h_bytes_per_step": out.numel() * esize * world, "ms_per_step": dt * 1e3, "steps": e2e_steps,
               "matches_device_result": same}
        del host_t, host_np

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gbs, eps, sample, _ = cpu_reference_run(name, 1, 12.0, shape_lanes=1024)
        cpu_baseline = {"value": gbs, "unit": "GB/s", "cores": 1, "kind": "port", "sample": sample,
                        "elems_per_s": eps}

    if rank == 0:
        line = {
            "metric": "quantile input GB/s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": {"float32": "f32", "float64": "f64", "uint32": "u32"}[dtype],
            "data": "synthetic",
            "config": {"workload": f"{name}: {desc}", "per_gpu_shape": list(t.shape), "axis": axis, "qs": len(qs),
                       "interp": interp, "skipnan": skipnan, "sharding": ("1-D array split over ranks + NCCL histogram allreduce" if sharded_1d else "lanes, no collective"),
                       "l2": "inputs larger than L2 (no flush needed)" if t.numel() * esize > 256e6 else "input smaller than L2: L2-resident between steps",
                       "scale_down": args.scale_down},
            "elems_per_s": elems_per_s,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks, "parity_sample": parity, "path": path,
        }
        if args.scale_down > 1:
            line["invalid"] = "scaled-down debug run"
        print(json.dumps(line))
    if distributed:
        if sharded_1d:
            ctx.comm_destroy()
        dist.destroy_process_group()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
