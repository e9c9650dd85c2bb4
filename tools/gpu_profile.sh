#!/bin/bash
# ncu evidence for profiles/: (1) launch list with device time per kernel, (2) full-set capture of
# the dominant kernel.  One GPU only.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/launches_run.log 2>&1
tail -2 gpurun_out/launches_run.log | cut -c1-300
# (2) every GEMM launch of the 4th step (51 per step): DRAM traffic + tensor-pipe activity (light metric set)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
  --clock-control none -k regex:gemm_tc_kernel -s 153 -c 51 --csv --log-file gpurun_out/gemm_traffic.csv \
  python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/traffic_run.log 2>&1
tail -1 gpurun_out/traffic_run.log | cut -c1-200
# (3) full-set capture of 10 launches around the NerfMLP forward/backward turn (report must stay < 64 MiB)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 158 -c 10 \
  -o gpurun_out/gemm_tc python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/full_run.log 2>&1
tail -1 gpurun_out/full_run.log | cut -c1-200
ls -la gpurun_out
