#!/bin/bash
# Partial evidence pack after a change of the FGK kernels only (run under gpurun): per-class instruction counts, the default bench line.
bash tools/ncu_fgk_classes.sh
python bench.py --steps 6 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err || tail -5 gpurun_out/r02_bench_n1.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_n1.json')); print('value', round(d['value'],3), round(d['ms_per_step'],2), 'seq', round(d['sequential_ms_per_step'],2), 'e2e', round(d['e2e']['value'],3), {k:(round(v['ms'],2), round(v['frac'],3)) for k,v in d['fgk'].items()})
P
