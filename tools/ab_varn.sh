# A/B of the narrow diagonal-block MMAs on one box: bash tools/ab_varn.sh   (16384 = GDRF_FLAG_FULL_WIDTH)
for f in 16384 0 16384 0; do
  GDRF_BENCH_FLAGS=$f python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']['all_contractions_ms_per_step']; print('flags', $f, round(d['ms_per_step'],1), d['clocks']['sm_mhz'], {k: round(v,1) for k,v in r.items()})"
done
