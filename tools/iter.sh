#!/bin/bash
# One GPU iteration on the bracket path: GPU parity tests, C4/C5 bench lines, launch list and a source-level capture of
# the streaming kernel.  usage: tools/iter.sh <tag> [pytest-args]
set -u
TAG=${1:-it}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q --maxfail=12 ${2:-} > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
timeout 60 python tools/dbg_bracket2.py 7 > $OUT/dbg2.log 2>&1; tail -12 $OUT/dbg2.log
for w in c4 c5; do
  timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "bench $w rc=$?"; cut -c1-330 $OUT/bench_$w.json
done
CMD="python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c4.csv $CMD > $OUT/ncu_launches_c4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'brk_stream' -s 3 -c 1 -f -o $OUT/prof_c4 $CMD > $OUT/ncu_full_c4.log 2>&1
SEEDS=2 timeout 240 bash tools/stress_bracket.sh > $OUT/stress.log 2>&1; grep -c "^==" $OUT/stress.log; grep -E "^bad|rror" $OUT/stress.log | head
