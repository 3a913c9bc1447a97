#!/usr/bin/env python3
"""Aggregates an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list by
kernel and by pipeline stage: launches, mean duration (cold cache, serialised: compare SHARES), DRAM bytes per launch.
    python tools/launch_summary.py gpurun_out/r02_ncu_launches_ma.csv [more.csv ...] > profiles/r02_ncu_launch_summary.json
Also writes profiles/r02_traffic.json (DRAM bytes per stage and step) that bench.py puts into roofline.stages[].traffic."""
import csv
import json
import sys
from collections import OrderedDict, defaultdict

STAGE = [("diff_apply", ("diff_apply_kernel",)), ("diff_revert", ("diff_segsum_kernel", "diff_revert_kernel")),
         ("adapt_encode", ("adapt_cost_mask_kernel", "adapt_cost_kernel", "adapt_select_kernel", "adapt_emit_kernel", "adapt_gather_large_kernel",
                           "adapt_emit_large_kernel", "adapt_emit_small_kernel")),
         ("adapt_decode", ("adapt_index_warp_kernel", "adapt_index_cta_kernel", "adapt_expand_kernel", "adapt_expand_large_kernel",
                           "adapt_scatter_large_kernel", "adapt_expand_small_kernel", "adapt_decode_kernel")),
         ("fgk_encode", ("fgk_encode_kernel",)), ("fgk_decode", ("fgk_decode_kernel",)),
         ("rle_encode", ("rle_encode_kernel",)), ("rle_decode", ("rle_decode_kernel",))]

per = defaultdict(lambda: defaultdict(list))
for path in sys.argv[1:]:
    ids = {}
    for row in csv.reader(open(path)):
        if len(row) < 15 or not row[0].isdigit():
            continue
        name = row[4].split("(")[0].replace("hcd::", "")
        per[name][row[12]].append(float(row[14]))
out = OrderedDict()
for name, m in sorted(per.items(), key=lambda kv: -sum(kv[1].get("gpu__time_duration.sum", [0]))):
    t = m.get("gpu__time_duration.sum", [])
    n = len(t)
    out[name] = {"launches": n, "mean_ms": sum(t) / n / 1e6 if n else None, "total_ms": sum(t) / 1e6,
                 "dram_read_MB_per_launch": sum(m.get("dram__bytes_read.sum", [0])) / max(n, 1) / 1e6,
                 "dram_write_MB_per_launch": sum(m.get("dram__bytes_write.sum", [0])) / max(n, 1) / 1e6}
tot = sum(v["total_ms"] for v in out.values())
for v in out.values():
    v["share_of_listed_time"] = v["total_ms"] / tot
stages, traffic = OrderedDict(), {}
# launches per step of the stage's main kernel = number of steps captured
for st, kerns in STAGE:
    ks = [k for k in kerns if k in out]
    if not ks:
        continue
    steps = max(out[ks[-1]]["launches"] if st.startswith("fgk") else max(out[k]["launches"] for k in ks), 1)
    if st in ("adapt_encode", "adapt_decode", "diff_revert"):
        steps = max(out[k]["launches"] for k in ks)
    ms = sum(out[k]["total_ms"] for k in ks) / steps
    by = sum((out[k]["dram_read_MB_per_launch"] + out[k]["dram_write_MB_per_launch"]) * out[k]["launches"] for k in ks) * 1e6 / steps
    stages[st] = {"kernels": ks, "steps_captured": steps, "ms_per_step": ms, "dram_bytes_per_step": by}
    traffic[st] = int(by)
json.dump(traffic, open("profiles/r02_traffic.json", "w"), indent=1)
print(json.dumps({"kernels": out, "stages": stages, "note": "ncu per-launch times are cold-cache and serialised: shares, not absolutes"}, indent=1))
