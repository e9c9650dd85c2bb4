#!/bin/bash
# ncu evidence for profiles/: (1) launch list with device time per kernel, (2) full-set capture of
# the dominant kernel.  One GPU only.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/launches_run.log 2>&1
tail -2 gpurun_out/launches_run.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 60 -c 6 \
  -o gpurun_out/gemm_tc python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/full_run.log 2>&1
tail -2 gpurun_out/full_run.log | cut -c1-300
ls -la gpurun_out
