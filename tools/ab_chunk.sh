for c in 0 37888 56832 0 37888; do
  GDRF_BENCH_CHUNK_ROWS=$c python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']['all_contractions_ms_per_step']; print('chunk', $c, round(d['ms_per_step'],1), d['clocks']['sm_mhz'], d['gpu_launches'], {k: round(v,1) for k,v in r.items()})"
done
