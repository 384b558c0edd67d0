#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--out profiles/name.json] [--all-matching REGEX]
"""
import argparse
import csv
import io
import json
import re
import subprocess
import sys

KEEP = [
    r"^gpu__time_duration\.sum$", r"^dram__bytes_(read|write)\.sum$", r"^dram__bytes_(read|write)\.sum\.per_second$",
    r"^launch__(grid_size|block_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit_.*|waves_per_multiprocessor)$",
    r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__inst_executed_pipe_tensor_subpipe_dmma\.avg\.pct_of_peak_sustained_active$",
    r"^sm__inst_executed_pipe_(fp64|lsu|alu|fma|xu)\.avg\.pct_of_peak_sustained_active$",
    r"^sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active$",
    r"^sm__issue_active\.avg\.pct_of_peak_sustained_elapsed$",
    r"^lts__t_sector_hit_rate\.pct$", r"^l1tex__t_sector_hit_rate\.pct$",
    r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared(_op_(ld|st))?\.sum$",
    r"^l1tex__data_pipe_lsu_wavefronts_mem_shared(_op_(ld|st))?\.sum$",
    r"^smsp__average_warps?_issue_stalled_.*_per_issue_active\.ratio$",
    r"^smsp__average_warp_latency_issue_stalled_.*\.ratio$",
    r"^smsp__inst_executed\.sum$", r"^smsp__sass_thread_inst_executed_op_(dfma|dmul|dadd)_pred_on\.sum$",
    r"^sm__sass_inst_executed_op_shared_(ld|st)\.sum$", r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^l1tex__throughput\.avg\.pct_of_peak_sustained_active$", r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__cycles_elapsed\.max$", r"^smsp__cycles_active\.avg$",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--out")
    ap.add_argument("--grep", default=None, help="extra regex of metric names to keep")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    pats = [re.compile(p) for p in KEEP] + ([re.compile(args.grep)] if args.grep else [])
    out = []
    for vals in rows[2:]:
        rec = {}
        for h, u, v in zip(hdr, units, vals):
            if h in ("Kernel Name", "Block Size", "Grid Size") or any(p.search(h) for p in pats):
                try:
                    v = float(v.replace(",", ""))
                except ValueError:
                    pass
                rec[h + (" [" + u + "]" if u else "")] = v
        # keep only the non-trivial stall reasons
        rec = {k: v for k, v in rec.items() if not ("issue_stalled" in k and isinstance(v, float) and v < 0.05)}
        out.append(rec)
    text = json.dumps(out, indent=1)
    if args.out:
        open(args.out, "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
