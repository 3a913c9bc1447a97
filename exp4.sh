timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | cut -c1-700
timeout 300 python bench.py --no-adapt --no-cpu --no-e2e > gpurun_out/bench_final_m.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_ncu_launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_lf.log 2>&1
