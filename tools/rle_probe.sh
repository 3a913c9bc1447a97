#!/bin/bash
# quick probe of the MNP-5 RLE kernels on one GPU (run under gpurun): stage parity tests, the -m bench line, one full
# ncu capture of each RLE kernel; build variants under tools/_variants/ are timed as well
python -m pytest tests/test_stages.py -m gpu -q -x 2>&1 | tail -3
CMDM="python bench.py --workload c3m --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras --overlap 1"
show() { python - "$1" <<'P'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], {k:round(v,3) for k,v in d['stage_ms']['compress'].items()}, {k:round(v,3) for k,v in d['stage_ms']['decompress'].items()})
P
}
$CMDM > gpurun_out/rle_probe_m.json 2> gpurun_out/rle_probe_m.err || tail -5 gpurun_out/rle_probe_m.err
show gpurun_out/rle_probe_m.json
for v in tools/_variants/*.so; do
  [ -f "$v" ] || continue
  HC_B200_DEBUG=1 HC_B200_LIB=$PWD/$v $CMDM > gpurun_out/rle_probe_$(basename $v .so).json 2>> gpurun_out/rle_probe_m.err && show gpurun_out/rle_probe_$(basename $v .so).json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rle_ -s 4 -c 2 -o gpurun_out/rle_probe_full -f $CMDM > /dev/null 2>&1
ls -la gpurun_out/rle_probe*
