# A/B: default library vs the NDS_AB_VARIANT build, interleaved
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['parity_sample'])"; }
for i in 1 2 3; do echo A; run; echo B; NDS_LIB_PATH=$PWD/ndarray-stats_b200/libndstats_b200_ab.so run; done
