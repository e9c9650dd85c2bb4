#!/bin/bash
# ncu launch list (device time per kernel) of one eager train step.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/launches_run.log 2>&1
tail -1 gpurun_out/launches_run.log | cut -c1-300
