timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 500 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench28.json 2> gpurun_out/bench28.err
python -c "
import json;d=json.load(open('gpurun_out/bench28.json'));print(d['value'],d['ms_per_step'],d['stage_ms'])"
for cfg in "random 148" "walk 1024"; do
  set -- $cfg
  HC_BENCH_CLASSES=$1 timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --files $2 > gpurun_out/exp_$1_$2.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/exp_$1_$2.json"))
print("$1 $2", d["ms_per_step"], {k:(round(v["ms"],2), v["symbols"], round(v["ns_per_symbol_longest_stream"],1)) for k,v in d["fgk"].items()})
PY
done
