#!/bin/bash
# First GPU call of the next round: the experiments that were prepared but could not be run in round 1 (no GPU minutes
# left).  Everything runs under `timeout`; the experimental strip kernel has never executed on a GPU.
# usage (1 GPU): gpurun --timeout 600 -- 'bash tools/next_experiments.sh'
set -u
OUT=gpurun_out/next; mkdir -p $OUT
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
# 1. cols_pipelined_kernel (NDS_COLS_V2=1): parity first, then C3 with and without it
NDS_COLS_V2=1 timeout 120 python -m pytest tests -m gpu -q -x -k "cols or skipnan" > $OUT/pytest_cols_v2.log 2>&1; echo "cols v2 pytest rc=$?"; tail -3 $OUT/pytest_cols_v2.log
for v in 0 1; do
  NDS_COLS_V2=$v timeout 60 python bench.py --workload c3 $B > $OUT/bench_c3_v2_$v.json 2> $OUT/bench_c3_v2_$v.err; echo "c3 v2=$v rc=$? $(cut -c1-140 $OUT/bench_c3_v2_$v.json)"
done
# 2. bracket select: a larger sample (fewer candidates) against the default (~n/256)
for s in 0 12582912 16777216; do
  NDS_BRK_SAMPLE=$s timeout 60 python bench.py --workload c4 $B > $OUT/bench_c4_s$s.json 2> $OUT/bench_c4_s$s.err; echo "c4 sample=$s rc=$? $(cut -c1-140 $OUT/bench_c4_s$s.json)"
done
# 3. the tighter bracket (4.5 sigma): candidates -10 %, fallback probability ~1e-3 per call
NDS_BRK_SIGMA=4.5 timeout 60 python bench.py --workload c4 $B > $OUT/bench_c4_sigma45.json 2> $OUT/bench_c4_sigma45.err; echo "c4 sigma=4.5 rc=$? $(cut -c1-140 $OUT/bench_c4_sigma45.json)"
