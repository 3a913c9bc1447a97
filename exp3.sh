timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-adapt > gpurun_out/bench26m.json 2> gpurun_out/bench26m.err
python -c "
import json;d=json.load(open('gpurun_out/bench26m.json'));print(d['value'],d['ms_per_step'],d['stage_ms'])"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench26.json 2> gpurun_out/bench26.err
python -c "
import json;d=json.load(open('gpurun_out/bench26.json'));print(d['value'],d['ms_per_step'],d['stage_ms'])"
