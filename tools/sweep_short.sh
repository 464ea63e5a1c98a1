# tuning sweep for the short-lane kernel on the C2 workload: ring stages x warps per SM
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_sample'])"; }
for s in 1 2 3; do for w in 8 12 16; do echo "stages=$s warps<=$w"; NDS_SHORT_STAGES=$s NDS_SHORT_WARPS=$w run; done; done
echo default; run
