#!/usr/bin/env python3
"""Compact JSON of the metrics worth keeping from an `ncu --set full` report (read here with `ncu -i ... --page raw --csv`):
    python tools/ncu_rep_metrics.py gpurun_out/r02_rle_full.ncu-rep "what was captured" > profiles/r02_ncu_rle_full_metrics.json
One entry per captured kernel launch."""
import csv
import io
import json
import subprocess
import sys

KEEP = ("Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_blocks", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.avg.per_cycle_active")
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = {"what": sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], "launches": []}
for r in rows[2:]:
    e = {}
    for i, h in enumerate(hdr):
        if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h):
            v = r[i]
            try:
                if h != "Kernel Name" and float(v.replace(",", "")) == 0.0:
                    continue
            except ValueError:
                pass
            e[h] = v if h == "Kernel Name" else [v, units[i]]
    e["Kernel Name"] = e["Kernel Name"].split("(")[0]
    out["launches"].append(e)
print(json.dumps(out, indent=1))
