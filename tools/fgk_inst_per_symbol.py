#!/usr/bin/env python3
"""Turns the per-class ncu metric captures of the FGK kernels into profiles/r02_fgk_inst_per_symbol.json.

Capture (one GPU, see tools/ncu_fgk_classes.sh): for every class c of the C3 workload
    HC_BENCH_CLASSES=c python bench.py --files 296 --steps 1 --warmup 1 --no-cpu --no-e2e --no-extras --overlap 1 > r2_cls_c.json
    ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,... -k regex:fgk_.*code -s 2 -c 2 --csv --log-file r2_cls_c.csv <same command>
bench.py multiplies these per-symbol instruction counts with the symbol counts of its own run (roofline of an
issue-bound kernel: warp instructions per second against 148 SMs x 4 schedulers x SM clock)."""
import csv
import json
import os
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
out = {"fgk_encode": {}, "fgk_decode": {}, "_detail": {}}
for cls in ("random", "walk", "smooth", "const"):
    d = json.load(open(os.path.join(src, "r2_cls_%s.json" % cls)))
    nsym = d["fgk"]["fgk_encode"]["symbols"]
    met = {}
    for row in csv.reader(open(os.path.join(src, "r2_cls_%s.csv" % cls))):
        if len(row) > 14 and row[4].startswith("fgk_"):
            met.setdefault(row[4].split("(")[0], {})[row[12]] = float(row[14])
    for k, name in (("fgk_encode_kernel", "fgk_encode"), ("fgk_decode_kernel", "fgk_decode")):
        m = met[k]
        out[name][cls] = m["smsp__inst_executed.sum"] / nsym
        out["_detail"]["%s/%s" % (name, cls)] = {
            "symbols": nsym, "streams": 296, "warp_inst": m["smsp__inst_executed.sum"], "ms_under_ncu": m["gpu__time_duration.sum"] / 1e6,
            "ns_per_symbol_two_streams_per_sm": d["fgk"][name]["ns_per_symbol_longest_stream"],
            "issue_active_pct": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "dram_bytes": m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)}
json.dump(out, open("profiles/r02_fgk_inst_per_symbol.json", "w"), indent=1)
print(json.dumps({k: out[k] for k in ("fgk_encode", "fgk_decode")}, indent=1))
