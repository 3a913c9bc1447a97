for i in 1 2; do timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-adapt > gpurun_out/bench_m_$i.json 2>/dev/null
python -c "
import json;d=json.load(open('gpurun_out/bench_m_$i.json'));print(d['value'],d['ms_per_step'],d['stage_ms'])"; done
