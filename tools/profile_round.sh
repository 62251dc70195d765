# Round profile capture (B200_PROFILING.md recipe): bash tools/profile_round.sh r02
# 1) the plain run must exit 0 first; 2) launch list of one step; 3) one --set full capture of every per-chunk kernel of
# one 18 944-observation chunk of the C4 shape.  Summaries land in gpurun_out/; copy what is to be judged to profiles/.
set -e
TAG=$1
mkdir -p gpurun_out
CMD="python bench.py --obs 75776 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${TAG}_plain.json
# one whole step: everything between two M x M prologues (k_kuu<double> opens a prologue); -s skips the warm-up steps
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 260 --csv --log-file gpurun_out/${TAG}_launches_n75776.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
# the prologue alone (C4 / C1 / M = 2048 shapes, tools/time_prologue.py), launch by launch
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_chol|k_trinv|k_kuu|k_pack|k_merge" -c 50 --csv --log-file gpurun_out/${TAG}_prologue_launches.csv python tools/time_prologue.py > gpurun_out/${TAG}_ncu0.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:gemm_tc2|k_scale_w|k_likelihood|k_reduce_dphi|k_dw_finalize|k_kxz_planes|k_kxz_backward|k_obs_" -s 180 -c 15 -o gpurun_out/${TAG}_chunk -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ncu -i gpurun_out/${TAG}_chunk.ncu-rep --page raw --csv > gpurun_out/${TAG}_chunk_ncu_raw.csv
python tools/ncu_summary.py gpurun_out/${TAG}_chunk_ncu_raw.csv > gpurun_out/${TAG}_gemm_ncu_summary.json
python tools/launch_summary.py gpurun_out/${TAG}_launches_n75776.csv | head -40
python -c "
import json
for k, v in json.load(open('gpurun_out/${TAG}_gemm_ncu_summary.json')).items():
    print(k[:44].ljust(44), round(v['duration_ms'], 3), 'ms  tensor', round(v['tensor_pipe_active_pct'], 1), '%  clk', round(v['sm_ghz'], 2), 'GHz  dram', round(v['dram_read_bytes'] / 1e9, 2), '+', round(v['dram_write_bytes'] / 1e9, 2), 'GB', ' x distinct', v.get('traffic_over_distinct'), ' L2 hit', v.get('l2_hit_pct'), ' dram%', v.get('dram_pct_of_peak'))
"
