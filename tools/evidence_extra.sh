#!/bin/bash
# bench line + ncu launch list + one `--set full` capture of the dominant kernel for one non-default workload
# usage: tools/evidence_extra.sh <tag> <workload> <kernel-regex>
set -u
TAG=$1; WL=$2; KRE=$3
OUT=gpurun_out/$TAG; mkdir -p $OUT
python bench.py --workload $WL --steps 10 --no-e2e > $OUT/bench_$WL.json 2> $OUT/bench_$WL.err; echo "bench $WL rc=$?"
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > $OUT/plain_$WL.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$WL.csv $CMD > $OUT/ncu_launches_$WL.log 2>&1
$CMD > $OUT/plain2_$WL.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -f -o $OUT/prof_$WL $CMD > $OUT/ncu_full_$WL.log 2>&1
cat $OUT/bench_$WL.json | cut -c1-300
