#!/bin/bash
# Partial evidence pack after a change of the RLE kernels only (run under gpurun): the default bench line, codec-stream
# counts beyond 4, the -m launch list with DRAM bytes, one full ncu capture of each RLE kernel.
python bench.py --steps 6 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err || tail -5 gpurun_out/r02_bench_n1.err
for ov in 6 8; do python bench.py --steps 16 --warmup 3 --no-cpu --no-e2e --no-extras --overlap $ov > gpurun_out/r02_overlap$ov.json 2>> gpurun_out/r02_bench_n1.err; done
python bench.py --steps 16 --warmup 3 --no-cpu --no-e2e --no-extras --overlap 4 > gpurun_out/r02_overlap4.json 2>> gpurun_out/r02_bench_n1.err
CMDM="python bench.py --workload c3m --steps 2 --warmup 1 --no-cpu --no-e2e --no-extras --overlap 1"
$CMDM > gpurun_out/r02_plain_m.json 2> gpurun_out/r02_plain.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:rle_ -c 40 --csv --log-file gpurun_out/r02_ncu_launches_m.csv $CMDM > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:rle_ -s 4 -c 2 -f -o gpurun_out/r02_rle_full $CMDM > /dev/null 2>&1
python - <<'P'
import json
for f in ('r02_bench_n1','r02_overlap4','r02_overlap6','r02_overlap8'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, 'value', round(d['value'],3), 'ms', round(d['ms_per_step'],2), 'e2e', (d.get('e2e') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e)
P
ls -la gpurun_out/r02_*
