"""Condenses `ncu -i X.ncu-rep --page raw --csv` into the per-kernel JSON bench.py reads for `roofline.traffic`.

    python tools/ncu_summary.py gpurun_out/r02_chunk_ncu_raw.csv > profiles/r02_gemm_ncu_summary.json

Besides the measured counters every kernel gets `distinct_bytes`: the bytes of distinct operands it has to read plus the
bytes it has to write for one 18 944-observation chunk of the C4 shape (K = 32, V = 512, M = Mp = 1024, D = 3) -- what a
launch would move if nothing were read twice -- and `traffic_over_distinct` = measured DRAM bytes / that.
"""
import csv
import json
import sys

COLS = {"duration_ms": "gpu__time_duration.sum",
        "tensor_pipe_active_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "tensor_pipe_elapsed_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram_read_bytes": "dram__bytes_read.sum", "dram_write_bytes": "dram__bytes_write.sum",
        "sm_cycles_active": "sm__cycles_active.avg", "regs": "launch__registers_per_thread",
        "l2_hit_pct": "lts__t_sector_hit_rate.pct", "dram_pct_of_peak": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "issue_active_pct": "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "achieved_occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active"}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}

N, K, V, M, D = 18944, 32, 512, 1024, 3
NM = N * M
TRI = 10.0 / 16.0            # share of the 64-row k-blocks of a lower-triangular [M x M] operand that 256-wide tiles touch
DISTINCT = {                 # bytes: distinct reads + writes per launch
    "G1T": 3 * 2 * NM + 3 * 2 * M * M * TRI + (3 + 2) * 2 * NM + 8 * N,             # Kxz, Linv -> W (3 planes) + W16 + wsq
    "G2<2>": 2 * 2 * NM + 2 * 2 * K * M * M * TRI + 2 * 2 * NM * K + 8 * K * N,     # W16, ST16 -> T (2 planes), q
    "<G3>": 2 * 2 * NM * K + 2 * 2 * K * M * M * TRI + 4 * NM + 4 * K * N,    # T, ST16N, g2 -> dW
    "<G6>": 2 * 2 * NM * K * 2 + 4 * K * M * M * TRI * 2,                     # WG, T -> dS (read-modify-write)
    "G4T": 2 * 2 * NM + 2 * 2 * M * M * TRI + 4 * NM,                               # dWtot, Linv16 -> dKxz
    "G5T": 2 * 2 * NM * 2 + 8 * M * M + 2 * 2 * N * 128,                            # dWtot (+ g_loc), W16 -> C5, du_loc
    "k_scale_w": 3 * 2 * NM + 4 * K * N + 2 * 2 * NM * K,                           # W, g2 -> WG
    "k_likelihood": 4 * N * V + 4 * K * N + 6 * 64 * V + 4 * K * N + 4 * (N // 128) * K * V,   # ws, theta, phi^T -> g1, dphi slabs
    "<GF>": 2 * 2 * NM + 8 * K * N,                                                 # W16 -> f_loc
    "k_reduce_dphi": 4 * (N // 128) * K * V + 8 * K * V,
    "k_dw_finalize": 3 * 2 * NM + 4 * NM + 2 * 2 * NM, "k_kxz_planes": 4 * D * N + 3 * 2 * NM,
    "k_kxz_backward": 4 * NM + 4 * D * N, "k_obs_prepare": (8 + 8 + 4 + 4 + 4) * K * N, "k_obs_finalize": (4 * 6) * K * N,
}


def main(path):
    rows = list(csv.reader(open(path)))
    head, units = rows[0], rows[1]
    out = {}
    for r in rows[2:]:
        name = r[head.index("Kernel Name")]
        rec = {}
        for key, col in COLS.items():
            if col not in head:
                continue
            i = head.index(col)
            try:
                rec[key] = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
            except ValueError:
                continue
        rec["dram_bytes"] = rec["dram_read_bytes"] + rec["dram_write_bytes"]
        rec["sm_ghz"] = rec["sm_cycles_active"] / (rec["duration_ms"] * 1e6)
        rec["sm_mhz"] = 1e3 * rec["sm_ghz"]
        rec["dram_gbs"] = rec["dram_bytes"] / (rec["duration_ms"] * 1e-3) / 1e9
        for tag, b in DISTINCT.items():
            if tag in name:
                rec["distinct_bytes"] = b
                rec["traffic_over_distinct"] = round(rec["dram_bytes"] / b, 2)
        if name not in out:          # first launch of each kernel (one chunk)
            out[name] = rec
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1])
