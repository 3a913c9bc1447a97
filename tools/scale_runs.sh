#!/bin/bash
# The multi-GPU evidence pack (run under `gpurun --gpus 8`): N-GPU parity test, then the bench at N = 8 for
# C3ma weak, C3 strong (fixed 4096-file batch), C5 (16384 files in total) and C4 (1024 files in total).
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
python -m pytest tests/test_multigpu.py -m gpu -q 2>&1 | tail -3
$TR --master-port 29601 bench.py --gpus $N --steps 6 --warmup 3 --no-cpu > gpurun_out/r2_n${N}_c3ma_weak.json 2> gpurun_out/r2_n${N}.err || tail -5 gpurun_out/r2_n${N}.err
$TR --master-port 29602 bench.py --gpus $N --steps 6 --warmup 3 --no-cpu --no-extras --total-files 4096 > gpurun_out/r2_n${N}_c3ma_strong4096.json 2> gpurun_out/r2_n${N}.err || tail -5 gpurun_out/r2_n${N}.err
$TR --master-port 29603 bench.py --gpus $N --steps 6 --warmup 3 --no-cpu --no-extras --workload c5 --total-files 16384 > gpurun_out/r2_n${N}_c5_strong16384.json 2> gpurun_out/r2_n${N}.err || tail -5 gpurun_out/r2_n${N}.err
$TR --master-port 29604 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu --no-extras --workload c4 --total-files 1024 --overlap 1 --depth 2 > gpurun_out/r2_n${N}_c4_strong1024.json 2> gpurun_out/r2_n${N}.err || tail -5 gpurun_out/r2_n${N}.err
for f in gpurun_out/r2_n${N}_*.json; do python - "$f" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], "value %.2f e2e %s ms/step %.1f seq %.2f scaling %s parity %d" % (d["value"], d["e2e"] and round(d["e2e"]["value"], 2), d["ms_per_step"], d["sequential_value"], d["scaling"], d["parity_checked_files"]))
PY
done
