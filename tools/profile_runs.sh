#!/bin/bash
# Round-2 evidence pack on ONE GPU (run under gpurun): tests, the default bench line, per-class FGK instruction
# counts, the launch list with DRAM bytes of one step (-m -a and -m), and one full ncu capture of the top kernel.
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 6 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err || tail -5 gpurun_out/r02_bench_n1.err
bash tools/ncu_fgk_classes.sh
for ov in 6 8; do python bench.py --steps 16 --warmup 3 --no-cpu --no-e2e --no-extras --overlap $ov > gpurun_out/r02_overlap$ov.json 2>> gpurun_out/r02_bench_n1.err; done
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-extras --overlap 1"
$CMD > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_ncu_launches_ma.csv $CMD > /dev/null 2>&1
CMDM="python bench.py --workload c3m --steps 2 --warmup 1 --no-cpu --no-e2e --no-extras --overlap 1"
$CMDM > gpurun_out/r02_plain_m.json 2> gpurun_out/r02_plain.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:rle_ -c 40 --csv --log-file gpurun_out/r02_ncu_launches_m.csv $CMDM > /dev/null 2>&1
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fgk_decode -s 4 -c 1 -o gpurun_out/r02_fgk_decode_full $CMD > /dev/null 2>&1
CMDM="python bench.py --workload c3m --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras --overlap 1"
ncu --set full --clock-control none --import-source on -k regex:rle_ -s 4 -c 2 -f -o gpurun_out/r02_rle_full $CMDM > /dev/null 2>&1
ls -la gpurun_out/r02_*
