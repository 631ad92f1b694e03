#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / bench.py cite:
   python profiles/scripts/summarize_ncu.py gpurun_out/r01_lstm_rec.ncu-rep > profiles/r01_lstm_rec_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_bytes.sum", "L2 bytes (all)"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 read sectors from SMs"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory throughput %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (realtime)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster size"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
]


def traffic(reps):
    """--traffic rec.ncu-rep gemm.ncu-rep loss.ncu-rep -> JSON {category: mean DRAM bytes (read+write) per launch}.
    The gemm capture must hold ALL tensor-core GEMM launches of one step (capture.sh: --launch-skip 72 -c 24)."""
    import json
    out = {}
    for cat, rep in zip(("recurrence", "gemm_tc", "loss"), reps):
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        col = {n: i for i, n in enumerate(hdr)}
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot, n = 0.0, 0
        for r in rows[2:]:
            b = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                b += float(r[col[key]]) * scale[units[col[key]]]
            tot += b
            n += 1
        out[cat] = tot / max(1, n)
        out[cat + "_launches_captured"] = n
    out["source"] = "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, mean per captured launch"
    print(json.dumps(out, indent=1))
    return 0


def main():
    if sys.argv[1] == "--traffic":
        return traffic(sys.argv[2:5])
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    print(f"# {rep}: {len(rows) - 2} captured launch(es); ncu --set full --clock-control none (replayed, cold caches: use shares, not absolutes)")
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0]
        print(f"\n== {name}   [launch id {r[col['ID']]}]")
        for key, label in KEYS:
            for n, i in col.items():
                if n == key:
                    print(f"  {label:38s} {r[i]:>18s} {units[i]}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
