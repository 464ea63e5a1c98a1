# tuning sweep for the short-lane kernel on the C2 workload: ring stages x warps per SM
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_sample'])"; }
for cfg in "1 16" "2 12" "1 12" "1 16"; do set -- $cfg; echo "stages=$1 warps<=$2"; NDS_SHORT_STAGES=$1 NDS_SHORT_WARPS=$2 run; done
echo default; run
