timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-adapt > gpurun_out/bench22m.json 2> gpurun_out/bench22m.err
python -c "
import json;d=json.load(open('gpurun_out/bench22m.json'));print(d['value'],d['ms_per_step'],d['stage_ms'])"
timeout 500 ncu --set full --import-source on --clock-control none -k regex:rle_encode -c 1 -o gpurun_out/rle_v3 -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-adapt > gpurun_out/ncu_rle3.log 2>&1; tail -2 gpurun_out/ncu_rle3.log
