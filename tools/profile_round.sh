# Round profile capture (B200_PROFILING.md recipe): bash tools/profile_round.sh r01i
# 1) the plain run must exit 0 first; 2) launch list; 3) one --set full capture of the pair kernels of one chunk.
set -e
TAG=$1
mkdir -p gpurun_out
CMD="python bench.py --n 75776 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${TAG}_plain.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 260 --csv --log-file gpurun_out/${TAG}_launches_n75776.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 72 -c 6 -o gpurun_out/${TAG}_gemm_tc2 -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ncu -i gpurun_out/${TAG}_gemm_tc2.ncu-rep --page raw --csv > gpurun_out/${TAG}_gemm_tc2_ncu_raw.csv
python tools/ncu_summary.py gpurun_out/${TAG}_gemm_tc2_ncu_raw.csv > gpurun_out/${TAG}_gemm_ncu_summary.json
python tools/launch_summary.py gpurun_out/${TAG}_launches_n75776.csv | head -24
python -c "
import json
for k, v in json.load(open('gpurun_out/${TAG}_gemm_ncu_summary.json')).items():
    print(k[:48], round(v['duration_ms'], 3), 'ms  tensor', round(v['tensor_pipe_active_pct'], 1), '/', round(v['tensor_pipe_elapsed_pct'], 1), '%  clk', round(v['sm_ghz'], 2), 'GHz  dram', round(v['dram_read_bytes'] / 1e9, 2), '+', round(v['dram_write_bytes'] / 1e9, 2), 'GB')
"
