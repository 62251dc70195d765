# A/B of GDRF_BENCH_FLAGS values on one box: bash tools/ab_flags.sh 0 524288 0 524288
for f in "$@"; do
  GDRF_BENCH_FLAGS=$f python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']['all_contractions_ms_per_step']; print('flags', $f, round(d['ms_per_step'],1), d['clocks']['sm_mhz'], {k: round(v,1) for k,v in r.items()})"
done
