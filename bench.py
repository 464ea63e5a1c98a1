#!/usr/bin/env python
"""Benchmark of the bulk quantile path (BASELINE.json metric: quantile input GB/s & elems/s, % of the HBM
roofline, next to the host CPU reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" is one pass of the hot path over one batch of synthetic input.  The headline workload (the one the
metric is quoted on, BASELINE.json configs[1]) is  c2 = 65536 x 4096 f32, row-wise median
(`quantile_axis_mut(Axis(1), 0.5, &Midpoint)`).  A default run (no --workload) also measures every other
BASELINE config and appends one sub-record per config under `workloads` (c1, c3, c4, c5_u32 Lower / Higher /
adversarial, c5_i64_lanes Lower / Higher) and, under `sharded_1d`, config 4 through the multi-GPU entry point
`nds_quantiles_1d_sharded` (strong scaling: ONE 2^31-element array split over the ranks).

With N > 1 (launched by torch.distributed.run, one process per GPU): headline c2 and c3 shard their lanes, every
rank owns its own block, no data-path collective (`scaling: weak`); `sharded_1d` splits one array over the N ranks and
exchanges counts / a few keys over NVLink.

Inputs are the counter-based splitmix64 streams of SURVEY.md §8d, generated on the device by `nds_synth_fill`; the CPU
legs regenerate the same bytes with numpy (`ndarray_stats_b200.synth.host`).

`value` times K enqueued calls on device-resident inputs with CUDA events on the launching stream (inputs >= 1 GiB,
larger than the 126 MB L2, so every step streams from HBM); `e2e` times the same calls through the public host-array
API (pinned host buffers, H2D + D2H inside the timed region).
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

Q99 = [k / 100.0 for k in range(1, 100)]
# name: description, dtype, shape, axis, qs, interp, skipnan, synthetic stream, seed
WORKLOADS = {
    "c1": dict(desc="1024x1024 f64 quantiles_axis_mut(Axis(1), [.25,.5,.75], Linear)", dtype="float64",
               shape=(1024, 1024), axis=1, qs=[0.25, 0.5, 0.75], interp="linear", skipnan=False,
               stream="f64_uniform", seed=42),
    "c2": dict(desc="65536x4096 f32 quantile_axis_mut(Axis(1), 0.5, Midpoint)", dtype="float32",
               shape=(65536, 4096), axis=1, qs=[0.5], interp="midpoint", skipnan=False, stream="f32_uniform", seed=1002),
    "c3": dict(desc="8192x65536 f32 quantile_axis_skipnan_mut(Axis(0), 0.5, Nearest), 10% NaN", dtype="float32",
               shape=(8192, 65536), axis=0, qs=[0.5], interp="nearest", skipnan=True, stream="f32_uniform_nan10",
               seed=1003),
    "c4": dict(desc="1-D f64 2^31 quantiles_mut(k/100, k=1..99, Linear)", dtype="float64", shape=(1 << 31,), axis=0,
               qs=Q99, interp="linear", skipnan=False, stream="f64_uniform", seed=1004),
    "c5_u32_lower": dict(desc="1-D u32 2^30, 16 distinct values k*0x11111111, quantiles_mut([.1,.5,.9], Lower)",
                         dtype="uint32", shape=(1 << 30,), axis=0, qs=[0.1, 0.5, 0.9], interp="lower", skipnan=False,
                         stream="u32_16values", seed=1005),
    "c5_u32_higher": dict(desc="1-D u32 2^30, 16 distinct values k*0x11111111, quantiles_mut([.1,.5,.9], Higher)",
                          dtype="uint32", shape=(1 << 30,), axis=0, qs=[0.1, 0.5, 0.9], interp="higher", skipnan=False,
                          stream="u32_16values", seed=1005),
    "c5_u32_adv": dict(desc="1-D u32 2^30, 16 distinct values 0x80000000+k (ties differ in the last digit), Lower",
                       dtype="uint32", shape=(1 << 30,), axis=0, qs=[0.1, 0.5, 0.9], interp="lower", skipnan=False,
                       stream="u32_16values_adv", seed=1005),
    "c5_i64_lanes_lower": dict(desc="16384x8192 i64 lanes ((z&15)-8)*2^40, quantiles_axis_mut(Axis(1), [.1,.5,.9], Lower)",
                               dtype="int64", shape=(16384, 8192), axis=1, qs=[0.1, 0.5, 0.9], interp="lower",
                               skipnan=False, stream="i64_16values", seed=1005),
    "c5_i64_lanes_higher": dict(desc="16384x8192 i64 lanes ((z&15)-8)*2^40, quantiles_axis_mut(Axis(1), [.1,.5,.9], Higher)",
                                dtype="int64", shape=(16384, 8192), axis=1, qs=[0.1, 0.5, 0.9], interp="higher",
                                skipnan=False, stream="i64_16values", seed=1005),
    # beyond BASELINE.json (the shapes between its configs): mid-length lanes and a tall matrix's columns
    "x_mid_f64_rows": dict(desc="2048x65536 f64 row medians (quantile_axis_mut Axis(1), Midpoint): CTA per lane",
                           dtype="float64", shape=(2048, 65536), axis=1, qs=[0.5], interp="midpoint", skipnan=False,
                           stream="f64_uniform", seed=1006),
    "x_tall_f32_cols": dict(desc="65536x4096 f32 column quartiles (quantiles_axis_mut Axis(0), [.25,.5,.75], Linear): "
                                 "layout stage + CTA per lane", dtype="float32", shape=(65536, 4096), axis=0,
                            qs=[0.25, 0.5, 0.75], interp="linear", skipnan=False, stream="f32_uniform", seed=1007),
}
ALIASES = {"c5": "c5_u32_lower"}
HEADLINE = "c2"
SUB_WORKLOADS_1GPU = ["c1", "c3", "c4", "c5_u32_lower", "c5_u32_higher", "c5_u32_adv", "c5_i64_lanes_lower",
                      "c5_i64_lanes_higher", "x_mid_f64_rows", "x_tall_f32_cols"]
SUB_WORKLOADS_MULTI = ["c3"]  # lanes shard with no collective (weak scaling); c4 runs as `sharded_1d`
DTYPE_TAG = {"float32": "f32", "float64": "f64", "uint32": "u32", "int64": "i64"}


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
        self.mark = 0

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

    def section(self):
        """Summary of the samples since the previous call (the sampler keeps running)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        rows, self.mark = self.rows[self.mark:], len(self.rows)
        if not rows and self.rows:
            rows = self.rows[-1:]  # a section shorter than the sampling period: the latest sample stands for it
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
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

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()


# ---------------------------------------------------------------------------------------------------------
# inputs
# ---------------------------------------------------------------------------------------------------------
def torch_dtype(torch, name):
    return {"float32": torch.float32, "float64": torch.float64, "uint32": torch.uint32, "int64": torch.int64}[name]


def make_device_input(torch, nds, ctx, w, shape, first_index):
    """Device-resident input of workload `w`: elements [first_index, first_index + numel) of its splitmix64 stream."""
    t = torch.empty(shape, dtype=torch_dtype(torch, w["dtype"]), device="cuda")
    nds.synth.fill(ctx, t, w["stream"], w["seed"], first_index)
    return t


def host_lane_sample(nds, w, nlanes):
    """The first `nlanes` lanes of workload `w` as a host array holding the same bytes as the device input."""
    shape, axis = w["shape"], w["axis"]
    if len(shape) == 1:
        return nds.synth.host(w["stream"], min(shape[0], nlanes), w["seed"])
    rows, cols = shape
    if axis == 1:  # lanes are rows: a contiguous prefix of the stream
        r = min(rows, nlanes)
        return nds.synth.host(w["stream"], r * cols, w["seed"]).reshape(r, cols)
    c = min(cols, nlanes)  # lanes are columns: element (i, j) has stream index i * cols + j
    idx = np.arange(rows, dtype=np.uint64)[:, None] * np.uint64(cols) + np.arange(c, dtype=np.uint64)[None, :]
    return nds.synth.host_at(w["stream"], idx, w["seed"])


def ordered_view(torch, t):
    """A tensor whose signed / float order is the element order of `t` (torch cannot compare uint32)."""
    if t.dtype == torch.uint32:
        return t.view(torch.int32) ^ (-2147483648)
    return t


def ordered_scalar(torch, t, v):
    if t.dtype == torch.uint32:
        return int(np.int32(np.uint32(v) ^ np.uint32(0x80000000)))
    if t.dtype.is_floating_point:
        return float(v)
    return int(v)


# ---------------------------------------------------------------------------------------------------------
# parity of a timed result (checker only; the oracle never sits in a timed region of the GPU arm)
# ---------------------------------------------------------------------------------------------------------
def rank_property_counts(torch, t, values, reduce=None):
    """count(x < v) and count(x <= v) over the (possibly sharded) lane for every v; `reduce` sums over ranks."""
    tv = ordered_view(torch, t)
    lt = torch.empty(len(values), dtype=torch.int64, device="cuda")
    le = torch.empty(len(values), dtype=torch.int64, device="cuda")
    for i, v in enumerate(values):
        s = ordered_scalar(torch, t, v)
        lt[i] = (tv < s).sum()
        le[i] = (tv <= s).sum()
    if reduce is not None:
        reduce(lt)
        reduce(le)
    return lt.cpu().numpy(), le.cpu().numpy()


def parity_1d(torch, w, t, call, n_total, reduce=None, all_ranks=True):
    """Size-independent check of a 1-D result at full size.  `call(qs, interp)` runs the path under test.
    Every requested rank r = floor / ceil(q (n - 1)) is selected through the path (Lower and Higher) and must satisfy
    count(x < v) <= r < count(x <= v) on the whole lane; the path's result for the workload's own interpolation must
    then equal the reference formula applied to those order statistics, bit for bit."""
    qs = w["qs"] if all_ranks else [w["qs"][0], w["qs"][len(w["qs"]) // 2], w["qs"][-1]]
    qarr = np.asarray(qs, dtype=np.float64)
    lower = call(qs, "lower").cpu().numpy()
    higher = call(qs, "higher").cpu().numpy()
    x = qarr * np.float64(n_total - 1)
    rl, rh, frac = np.floor(x).astype(np.int64), np.ceil(x).astype(np.int64), x - np.trunc(x)
    ok = True
    for vals, ranks in ((lower, rl), (higher, rh)):
        lt, le = rank_property_counts(torch, t, vals.tolist(), reduce)
        ok = ok and bool(np.all((lt <= ranks) & (ranks < le)))
    got = call(qs, w["interp"]).cpu().numpy()
    if w["interp"] == "linear" and got.dtype.kind == "f":
        want = lower + (frac * (higher - lower)).astype(got.dtype)  # interpolate.rs:126-144 (f64: two roundings, no FMA)
    elif w["interp"] == "lower":
        want = lower
    elif w["interp"] == "higher":
        want = higher
    else:
        want = None
    same = None
    if want is not None:
        ut = {4: np.uint32, 8: np.uint64}[got.dtype.itemsize]
        same = bool(np.array_equal(got.view(ut), want.view(ut)))
        ok = ok and same
    return {"kind": "rank property on the full lane + reference formula", "ranks_checked": int(2 * len(qs)),
            "interp_bit_exact": same, "ok": bool(ok)}


def parity_lanes(torch, oracle, w, t, out, nlanes_check=128):
    """A bounded sample of lanes of the timed result against the oracle on the same device-generated data."""
    axis = w["axis"]
    lanes = t.shape[1 - axis]
    pick = np.linspace(0, lanes - 1, num=min(nlanes_check, lanes)).astype(np.int64)
    idx = torch.from_numpy(pick).cuda()
    sub = t.index_select(1 - axis, idx).cpu().numpy()
    got = out.index_select(1 - axis, idx).cpu().numpy()
    # tie-heavy integer lanes: the faithful quickselect is quadratic on equal runs -> sort-and-index selector
    impl = 1 if w["stream"].startswith(("u32_16", "i64_16")) else 0
    st, want, _ = oracle.quantiles_axis(sub, axis, w["qs"], w["interp"], w["skipnan"], select_impl=impl)
    if st != 0:
        return {"lanes": int(pick.size), "ok": False, "why": f"oracle status {st}"}
    ut = {4: np.uint32, 8: np.uint64}[got.dtype.itemsize]
    g, x = got.copy(), want.copy()
    if g.dtype.kind == "f":
        g[g == 0] = 0
        x[x == 0] = 0
    ok = bool(np.array_equal(np.ascontiguousarray(g).view(ut), np.ascontiguousarray(x).view(ut)))
    return {"kind": "oracle on a sample of lanes of the timed result", "lanes": int(pick.size), "ok": ok}


# ---------------------------------------------------------------------------------------------------------
# CPU legs (the oracle: a C restatement of the reference algorithm; the Rust crate cannot be built here)
# ---------------------------------------------------------------------------------------------------------
CPU_SAMPLE_LANES = 4096  # fixed, whatever the host's core count (C2 / C3 / C5 lanes); C1 runs at full size


def cpu_reference_run(nds, name, threads, budget_s, max_reps=200):
    """Time the oracle on a bounded sample of the workload, on the SAME bytes the GPU arm sees.
    Returns (GB/s, elems/s, sample description, seconds, sample bytes)."""
    from oracle import oracle
    w = WORKLOADS[name]
    dt = np.dtype(w["dtype"])
    if len(w["shape"]) == 2:
        a = host_lane_sample(nds, w, CPU_SAMPLE_LANES)
        impl = 1 if w["stream"].startswith("i64_16") else 0
    else:
        # one lane cannot be split over threads; the faithful quickselect is quadratic on equal runs, so the
        # tie-heavy lanes are timed at 2^16 keys
        n = 1 << 16 if w["stream"].startswith("u32_16") else 1 << 24
        a = host_lane_sample(nds, w, n)
        impl = 0
    elems = a.size
    t_total, reps = 0.0, 0
    while (t_total < budget_s and reps < max_reps) or reps == 0:
        work = a.copy()
        t0 = time.perf_counter()
        st, _, _ = oracle.quantiles_axis(work, w["axis"], w["qs"], w["interp"], w["skipnan"], select_impl=impl,
                                         nthreads=threads, mutate=True)
        t_total += time.perf_counter() - t0
        reps += 1
        assert st == 0, f"oracle status {st}"
    sec = t_total / reps
    sample = (f"{'x'.join(map(str, a.shape))} {w['dtype']} (first lanes / prefix of the {name} splitmix64 stream, the bytes "
              f"the GPU arm reads), {reps} repetitions, {threads} thread(s)")
    if impl == 1:
        sample += "; sort-and-index selector (the faithful quickselect is quadratic on equal runs, SURVEY F6)"
    if w["stream"].startswith("u32_16"):
        sample += "; 2^16 keys only: the reference quickselect is quadratic on equal runs (2^24 keys ran at 3.97e4 elems/s in round 1)"
    return elems * dt.itemsize / sec / 1e9, elems / sec, sample, t_total, int(a.nbytes)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import _ndstats_import
    from oracle import oracle
    nds = _ndstats_import.load()  # only the numpy restatement of the input streams is used (no GPU call)
    # all host threads the process may use: torchrun exports OMP_NUM_THREADS=1, which is not the machine
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    threads = max(threads, oracle.max_threads())
    name = args.workload
    w = WORKLOADS[name]
    # `--steps K --warmup W`: each step is one bounded sample; the whole run stays within a few minutes
    per_step_budget = min(8.0, 120.0 / max(1, args.steps + args.warmup))
    for _ in range(args.warmup):
        cpu_reference_run(nds, name, threads, per_step_budget / 4)
    vals, samples, nbytes = [], None, 0
    for _ in range(args.steps):
        gbs, eps, samples, s, nbytes = cpu_reference_run(nds, name, threads, per_step_budget)
        vals.append(gbs)
    value = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": "quantile input GB/s", "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_TAG[w["dtype"]], "data": "synthetic",
        "config": config_of(name, w, list(w["shape"]), 1,
                            note="reference CPU algorithm (C restatement; the Rust crate cannot be built in this image), "
                                 "lanes split over all host threads"),
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": threads, "kind": "port", "sample": samples,
                         "sample_bytes": nbytes},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def config_of(name, w, per_gpu_shape, world, note=None, sharded=False):
    nbytes = int(np.prod(per_gpu_shape)) * np.dtype(w["dtype"]).itemsize
    c = {"workload": f"{name}: {w['desc']}", "per_gpu_shape": list(per_gpu_shape), "axis": w["axis"], "qs": len(w["qs"]),
         "interp": w["interp"], "skipnan": w["skipnan"],
         "sharding": ("ONE 1-D array split over the ranks; counts and a few keys exchanged over NVLink" if sharded
                      else "lanes, no collective"),
         "l2": ("inputs larger than L2 (no flush needed)" if nbytes > 256e6
                else "input smaller than L2: L2-resident between steps"),
         "input": f"splitmix64 stream '{w['stream']}', seed {w['seed']} (SURVEY.md §8d), generated on the device"}
    if note:
        c["note"] = note
    return c


# ---------------------------------------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------------------------------------
class Env:
    pass


def time_steps(env, step, steps, warmup):
    """`warmup` untimed steps, then `steps` timed ones between CUDA events on the launching stream; barrier +
    synchronize on both sides; max over ranks.  Returns (ms per step, launches over all ranks)."""
    torch, ctx = env.torch, env.ctx
    for _ in range(max(warmup, 3)):
        step()
    ctx.synchronize()
    env.barrier()
    launches0 = ctx.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(env.stream)
    for _ in range(steps):
        step()
    ev1.record(env.stream)
    ctx.synchronize()
    env.barrier()
    launches = ctx.kernel_launches - launches0
    ms_total = ev0.elapsed_time(ev1)
    if env.distributed:
        tt = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        env.dist.all_reduce(tt, op=env.dist.ReduceOp.MAX)
        ms_total = float(tt.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        env.dist.all_reduce(lt, op=env.dist.ReduceOp.SUM)
        launches = int(lt.item())
    return ms_total / steps, launches


def run_lane_workload(env, name, steps, warmup, want_parity=True):
    """One workload through nds_quantiles_axis_enqueue on device-resident input.  2-D workloads: every rank owns its own
    block of lanes (weak scaling, no collective); 1-D workloads run on one GPU only."""
    torch, nds, ctx = env.torch, env.nds, env.ctx
    w = WORKLOADS[name]
    shape = w["shape"]
    esize = np.dtype(w["dtype"]).itemsize
    numel = int(np.prod(shape))
    t = make_device_input(torch, nds, ctx, w, shape, env.rank * numel if len(shape) == 2 else 0)
    out_shape = list(shape)
    out_shape[w["axis"]] = len(w["qs"])
    out = torch.empty(out_shape, dtype=t.dtype, device="cuda")

    def step():
        ctx.quantiles_axis(t, w["axis"], w["qs"], w["interp"], skipnan=w["skipnan"], enqueue=True, out=out)

    env.sampler.section()
    ms, launches = time_steps(env, step, steps, warmup)
    clocks = env.sampler.section()
    path = ctx.last_path
    world = env.world if len(shape) == 2 else 1
    units = numel * world
    rec = {"ms_per_step": ms, "value": units * esize / (ms * 1e-3) / 1e9, "unit": "GB/s",
           "elems_per_s": units / (ms * 1e-3), "path": path, "gpu_launches": launches, "steps": steps,
           "clocks": clocks, "config": config_of(name, w, shape, world)}
    per_gpu_bytes = numel * esize + out.numel() * esize
    peak, peak_src = peaks()
    rec["frac"] = per_gpu_bytes / (ms * 1e-3) / 1e9 / peak
    rec["algorithmic_bytes_per_launch"] = per_gpu_bytes
    if want_parity and env.rank == 0:
        from oracle import oracle
        if len(shape) == 1:
            rec["parity_sample"] = parity_1d(
                torch, w, t, lambda qs, interp: ctx.quantiles_axis(t, 0, qs, interp, skipnan=w["skipnan"]), shape[0])
        else:
            rec["parity_sample"] = parity_lanes(torch, oracle, w, t, out)
    return rec, t, out


def run_sharded_1d(env, steps, warmup, want_parity=True):
    """BASELINE config 4 through nds_quantiles_1d_sharded: ONE 2^31-element f64 array split over the ranks (strong
    scaling).  On N > 1 ranks rank 0 also times the whole array on its own GPU through the same entry point, so that the
    record carries the speed-up measured inside one run."""
    torch, nds, ctx, dist = env.torch, env.nds, env.ctx, env.dist
    name = "c4"
    w = WORKLOADS[name]
    n_total = w["shape"][0]
    world, rank = env.world, env.rank
    base, rem = divmod(n_total, world)
    n_local = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    t = make_device_input(torch, nds, ctx, w, (n_local,), first)
    if env.distributed:
        uid = [nds.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init_rank(world, rank, uid[0])
    result = {}

    def step():
        result["out"] = ctx.quantiles_1d_sharded(t, w["qs"], w["interp"], w["skipnan"])  # collective, synchronous

    env.sampler.section()
    for _ in range(2):
        step()  # (first calls size the arena)
    info0 = ctx.comm_info()
    ms, launches = time_steps(env, step, steps, warmup)
    info1 = ctx.comm_info()
    calls = steps + max(warmup, 3)
    exchange = {"transport": info1["transport"], "ranks": info1["ranks"],
                "exchange_steps_per_call": (info1["exchange_steps"] - info0["exchange_steps"]) / calls,
                "host_syncs_per_call": (info1["host_syncs"] - info0["host_syncs"]) / calls}
    clocks = env.sampler.section()
    path = ctx.last_path
    reduce = (lambda x: dist.all_reduce(x)) if env.distributed else None
    parity = None
    if want_parity:  # collective: every rank takes part
        parity = parity_1d(torch, w, t, lambda qs, interp: ctx.quantiles_1d_sharded(t, qs, interp, w["skipnan"]), n_total,
                           reduce=reduce)
    esize = 8
    peak, _ = peaks()
    rec = {"workload": f"{name}: {w['desc']}", "entry": "nds_quantiles_1d_sharded", "n_gpus": world, "scaling": "strong",
           "ms_per_step": ms, "value": n_total * esize / (ms * 1e-3) / 1e9, "unit": "GB/s",
           "elems_per_s": n_total / (ms * 1e-3), "frac_of_n_gpu_roofline": n_total * esize / (ms * 1e-3) / 1e9 / (peak * world),
           "path": path, "gpu_launches": launches, "steps": steps, "clocks": clocks, "parity_sample": parity,
           "exchange": exchange,
           "config": config_of(name, w, (n_local,), world, sharded=True)}
    del t
    torch.cuda.empty_cache()
    if env.distributed:
        # the same array on ONE GPU through the same entry point (rank 0 only, a second context without a communicator)
        n1_ms = None
        if rank == 0:
            solo = nds.Context(env.local_rank)
            solo.set_stream(env.stream.cuda_stream)
            full = make_device_input(torch, nds, solo, w, (n_total,), 0)
            for _ in range(3):
                solo.quantiles_1d_sharded(full, w["qs"], w["interp"], w["skipnan"])
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(env.stream)
            for _ in range(steps):
                solo.quantiles_1d_sharded(full, w["qs"], w["interp"], w["skipnan"])
            ev1.record(env.stream)
            torch.cuda.synchronize()
            n1_ms = ev0.elapsed_time(ev1) / steps
            del full
            solo.close()
            torch.cuda.empty_cache()
        env.barrier()
        if rank == 0:
            rec["n1_ms_per_step_same_run"] = n1_ms
            rec["speedup_vs_n1"] = n1_ms / ms
    return rec


def run_e2e(env, name, t, out, steps, pinned=True):
    """The public host-array API: numpy in, numpy out, H2D + D2H inside the timed region (wall clock around
    synchronous calls, barrier on both sides, max over ranks)."""
    torch, ctx = env.torch, env.ctx
    w = WORKLOADS[name]
    out_shape = list(out.shape)
    if pinned:
        host_t = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host_t.copy_(t)
        host_np = host_t.numpy()
        host_out = torch.empty(out_shape, dtype=t.dtype, pin_memory=True).numpy()
    else:
        host_np = t.cpu().numpy().copy()   # plain pageable memory, what a numpy caller passes
        host_out = np.empty(out_shape, dtype=host_np.dtype)
    e2e_steps = max(2, min(steps, 5))
    ctx.quantiles_axis(host_np, w["axis"], w["qs"], w["interp"], skipnan=w["skipnan"], out=host_out)  # warm the arena
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.quantiles_axis(host_np, w["axis"], w["qs"], w["interp"], skipnan=w["skipnan"], out=host_out)
    env.barrier()
    dt = (time.perf_counter() - t0) / e2e_steps
    if env.distributed:
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        env.dist.all_reduce(tt, op=env.dist.ReduceOp.MAX)
        dt = float(tt.item())
    same = bool(torch.equal(torch.from_numpy(host_out).cuda().view(torch.uint8), out.view(torch.uint8)))
    esize = t.element_size()
    units_total = t.numel() * env.world
    return {"value": units_total * esize / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": t.numel() * esize * env.world,
            "d2h_bytes_per_step": out.numel() * esize * env.world, "ms_per_step": dt * 1e3, "steps": e2e_steps,
            "host_memory": "pinned" if pinned else "pageable", "matches_device_result": same, "path": ctx.last_path}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS) + sorted(ALIASES) + ["sharded_1d"],
                    help="measure only this workload (default: headline c2 + every other config under `workloads`)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="headline only")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    only = ALIASES.get(args.workload, args.workload)
    if args.impl == "reference":
        args.workload = only if only in WORKLOADS else HEADLINE
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import _ndstats_import
    nds = _ndstats_import.load()

    env = Env()
    env.torch, env.dist, env.nds = torch, dist, nds
    env.world = int(os.environ.get("WORLD_SIZE", "1"))
    env.rank = int(os.environ.get("RANK", "0"))
    env.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    env.distributed = env.world > 1
    torch.cuda.set_device(env.local_rank)
    if env.distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", env.local_rank))
    env.ctx = nds.Context(env.local_rank)
    env.stream = torch.cuda.Stream()      # a real stream: the CUDA events must sit on the launching stream
    torch.cuda.set_stream(env.stream)
    env.ctx.set_stream(env.stream.cuda_stream)

    def barrier():
        if env.distributed:
            dist.barrier()
        torch.cuda.synchronize()
    env.barrier = barrier
    env.sampler = ClockSampler(env.local_rank)
    env.sampler.start()
    time.sleep(0.3)
    rank = env.rank
    sub_steps = max(3, min(args.steps, 10))
    want_parity = not args.no_parity

    if only == "sharded_1d":
        rec = run_sharded_1d(env, args.steps, args.warmup, not args.no_parity)
        if rank == 0:
            print(json.dumps({"metric": "quantile input GB/s", "value": rec["value"], "unit": "GB/s", "n_gpus": env.world,
                              "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": rec["ms_per_step"],
                              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                              "data": "synthetic", "config": rec["config"], "sharded_1d": rec,
                              "gpu_launches": rec["gpu_launches"], "clocks": rec["clocks"]}))
        return finish(env)

    name = only or HEADLINE
    w = WORKLOADS[name]
    one_d = len(w["shape"]) == 1
    if one_d and env.distributed:
        raise SystemExit("1-D workloads on several GPUs run as --workload sharded_1d")
    rec, t, out = run_lane_workload(env, name, args.steps, args.warmup, want_parity)
    peak, peak_src = peaks()
    traffic, dominant = None, rec["path"]
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, one ncu --set full capture
        with open(tpath) as f:
            table = json.load(f)
        trec = table.get(name) or table.get(name.split("_")[0])
        if trec:
            traffic, dominant = trec["dram_bytes_per_launch"], trec["kernel"]
    achieved = rec["algorithmic_bytes_per_launch"] / (rec["ms_per_step"] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": dominant, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": rec["algorithmic_bytes_per_launch"],
                "frac_of_nominal_8000": achieved / 8000.0}   # against the nominal HBM3e figure as well

    e2e = e2e_pageable = None
    if not args.no_e2e and not one_d:
        e2e = run_e2e(env, name, t, out, args.steps, pinned=True)
        if not env.distributed:
            e2e_pageable = run_e2e(env, name, t, out, args.steps, pinned=False)
    del t, out
    torch.cuda.empty_cache()

    cpu_baseline = None
    if rank == 0 and env.world == 1 and not args.no_cpu_baseline:
        gbs, eps, sample, _, nbytes = cpu_reference_run(nds, name, 1, 12.0)
        cpu_baseline = {"value": gbs, "unit": "GB/s", "cores": 1, "kind": "port", "sample": sample,
                        "sample_bytes": nbytes, "elems_per_s": eps}

    workloads, sharded = {}, None
    if only is None and not args.no_workloads:
        for sub in (SUB_WORKLOADS_MULTI if env.distributed else SUB_WORKLOADS_1GPU):
            r, st, so = run_lane_workload(env, sub, sub_steps, args.warmup, want_parity)
            del st, so
            torch.cuda.empty_cache()
            workloads[sub] = r
        sharded = run_sharded_1d(env, sub_steps, args.warmup, want_parity)

    if rank == 0:
        line = {
            "metric": "quantile input GB/s", "value": rec["value"], "unit": "GB/s", "n_gpus": env.world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": rec["ms_per_step"],
            "higher_is_better": True, "scaling": "weak" if not one_d else "strong", "vs_baseline": None,
            "dtype": DTYPE_TAG[w["dtype"]], "data": "synthetic", "config": rec["config"],
            "elems_per_s": rec["elems_per_s"], "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "e2e_pageable": e2e_pageable, "gpu_launches": rec["gpu_launches"], "clocks": rec["clocks"],
            "parity_sample": rec.get("parity_sample"), "path": rec["path"],
        }
        if workloads:
            line["workloads"] = workloads
        if sharded:
            line["sharded_1d"] = sharded
        print(json.dumps(line))
    return finish(env)


def finish(env):
    env.sampler.stop()
    if env.distributed:
        env.ctx.comm_destroy()
        env.dist.destroy_process_group()
    env.ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
