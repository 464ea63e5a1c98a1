#!/bin/bash
# Round-end evidence on one B200, most important first: GPU parity tests, smoke, the default bench (both arms), the
# other workloads' bench lines, ncu launch lists and one `--set full` capture of each workload's dominant kernel.
# usage: tools/evidence_final.sh <tag>
set -u
TAG=${1:-r01e}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.csv 2>&1
timeout 400 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -1 $OUT/smoke.log
timeout 300 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench c2 rc=$?"; cut -c1-200 $OUT/bench_c2.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_c2_reference.json 2>> $OUT/bench_c2.err; echo "ref rc=$?"
for w in c3 c4 c5 c1; do
  timeout 200 python bench.py --workload $w --steps 10 --no-e2e > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "bench $w rc=$?"; cut -c1-160 $OUT/bench_$w.json
done
for w in c2 c3 c4 c5 c1; do
  CMD="python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$w.csv $CMD > $OUT/ncu_launches_$w.log 2>&1; echo "launches $w rc=$?"
done
for wk in "c2 short_bitslice" "c4 brk_stream" "c3 cols_bitslice" "c5 brk_stream" "c1 short_bitslice"; do
  set -- $wk
  CMD="python bench.py --workload $1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o $OUT/prof_$1 $CMD > $OUT/ncu_full_$1.log 2>&1; echo "full $1 rc=$?"
done
