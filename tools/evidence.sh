#!/bin/bash
# Round evidence pass on one B200: GPU parity tests, smoke, the default bench (both arms), the ncu launch
# list of the same bench command and one `--set full` capture of its dominant kernel.
# usage: tools/evidence.sh <tag> [workload]
set -u
TAG=${1:-r01}; WL=${2:-c2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.csv 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
python bench.py --workload $WL > $OUT/bench_$WL.json 2> $OUT/bench_$WL.err; echo "bench rc=$?"
python bench.py --workload $WL --impl reference --steps 3 --warmup 1 > $OUT/bench_${WL}_reference.json 2>> $OUT/bench_$WL.err
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > $OUT/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$WL.csv $CMD > $OUT/ncu_launches.log 2>&1
$CMD > $OUT/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'bitslice|long_|bracket' -s 3 -c 1 -f -o $OUT/prof_$WL $CMD > $OUT/ncu_full.log 2>&1
tail -2 $OUT/pytest_gpu.log; cat $OUT/smoke.log | tail -2; cat $OUT/bench_$WL.json
