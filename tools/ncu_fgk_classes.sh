#!/bin/bash
# Per-class warp-instruction counts of the FGK kernels (run under gpurun on one GPU; results in gpurun_out/).
for cls in random walk smooth const; do
  export HC_BENCH_CLASSES=$cls
  CMD="python bench.py --files 296 --steps 1 --warmup 1 --no-cpu --no-e2e --no-extras --overlap 1"
  $CMD > gpurun_out/r2_cls_$cls.json 2> gpurun_out/r2_cls.err &&
  ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none -k regex:fgk_.*code -s 2 -c 2 --csv --log-file gpurun_out/r2_cls_$cls.csv $CMD > /dev/null 2>&1
done
