#!/bin/bash
# every BASELINE workload once on one GPU (device-resident timing only), one JSON line each
OUT=gpurun_out/${1:-bench_all}; mkdir -p $OUT
for wl in c1 c2 c3 c4 c5; do
  timeout 600 python bench.py --workload $wl --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --no-e2e > $OUT/$wl.json 2> $OUT/$wl.err
  echo "$wl rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$OUT/$wl.json").read().strip().splitlines()[-1])
    print("$wl", d["path"], "ms", round(d["ms_per_step"],4), "GB/s", round(d["value"],1), "frac", round(d["roofline"]["frac"],4), "launches", d["gpu_launches"], d["parity_sample"])
except Exception as e:
    print("$wl failed", e); print(open("$OUT/$wl.err").read()[-2000:])
PY
done
