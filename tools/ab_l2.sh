# A/B of GDRF_BENCH_FLAGS values on one box, with the DRAM bytes of the big contractions from an ncu metrics pass:
#   bash tools/ab_l2.sh 0 524288 1572864
# (used for the L2 cache-hint experiments of round 2 on builds that carried the hints behind flag bits 19 / 20:
#  profiles/r02b_ab_l2_*.txt; the store-side hint that worked is now unconditional, the flag bits are gone)
mkdir -p gpurun_out
bash tools/ab_flags.sh "$@" "$@"
for f in "$@"; do
  GDRF_BENCH_FLAGS=$f ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k "regex:gemm_tc2" -s 28 -c 7 --csv --log-file gpurun_out/ab_l2_$f.csv python bench.py --obs 75776 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ab_l2_ncu_$f.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ab_l2_$f.csv')) if len(r)>10]
h=rows[0]
agg={}
for r in rows[1:]:
    d=dict(zip(h,r)); agg.setdefault(d['Kernel Name'][:34],{})[d['Metric Name'].split('.')[0].replace('__','_')[:28]]=d['Metric Value']
for k,v in agg.items():
    if 'G2' in k or 'G3' in k or 'G6' in k: print($f, k, v)
PY
done
