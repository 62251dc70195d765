"""Condenses `ncu -i X.ncu-rep --page raw --csv` into the per-kernel JSON bench.py reads for `roofline.traffic`.

    python tools/ncu_summary.py profiles/r01i_gemm_tc2_ncu_raw.csv > profiles/r01i_gemm_ncu_summary.json
"""
import csv
import json
import sys

COLS = {"duration_ms": "gpu__time_duration.sum",
        "tensor_pipe_active_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "tensor_pipe_elapsed_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram_read_bytes": "dram__bytes_read.sum", "dram_write_bytes": "dram__bytes_write.sum",
        "sm_cycles_active": "sm__cycles_active.avg", "regs": "launch__registers_per_thread"}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}


def main(path):
    rows = list(csv.reader(open(path)))
    head, units = rows[0], rows[1]
    out = {}
    for r in rows[2:]:
        name = r[head.index("Kernel Name")]
        rec = {}
        for key, col in COLS.items():
            i = head.index(col)
            rec[key] = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
        rec["dram_bytes"] = rec["dram_read_bytes"] + rec["dram_write_bytes"]
        rec["sm_ghz"] = rec["sm_cycles_active"] / (rec["duration_ms"] * 1e6)
        out[name] = rec
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1])
