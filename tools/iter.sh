#!/bin/bash
# One GPU iteration: GPU parity tests, bench lines of the given workloads, launch list + source-level capture of one
# workload's dominant kernel, short bracket stress.  usage: tools/iter.sh <tag> "<workloads>" <profiled-workload> <kernel-regex>
set -u
TAG=${1:-it}; WLS=${2:-"c4 c5"}; PW=${3:-c4}; KR=${4:-brk_stream}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q --maxfail=12 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
for w in $WLS; do
  timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "bench $w rc=$?"; cut -c1-160 $OUT/bench_$w.json
done
CMD="python bench.py --workload $PW --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$PW.csv $CMD > $OUT/ncu_launches_$PW.log 2>&1
SKIP=3; [ "$KR" = brk_stream ] || SKIP=3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$KR" -s $SKIP -c 1 -f -o $OUT/prof_$PW $CMD > $OUT/ncu_full_$PW.log 2>&1
SEEDS=${SEEDS:-1} timeout 200 bash tools/stress_bracket.sh > $OUT/stress.log 2>&1; grep -c "^==" $OUT/stress.log; grep -E "^bad|rror" $OUT/stress.log | sort | uniq -c | head
