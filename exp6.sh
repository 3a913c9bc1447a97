for g in 1 2 3 4 8; do
HC_GROUPS=$g timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_g$g.json 2>/dev/null
python -c "
import json;d=json.load(open('gpurun_out/bench_g$g.json'));print($g, d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'])"
done
