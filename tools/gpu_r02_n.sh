#!/bin/bash
# round 2, call n: source-level ncu capture of encode_fast_kernel (what paces it?)
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:encode_fast_kernel -s 9 -c 2 -o gpurun_out/enc \
  python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/enc_run.log 2>&1
ls -la gpurun_out/enc* 
ncu -i gpurun_out/enc.ncu-rep --page details --csv 2>/dev/null | grep -iE "Issue Slot|Pipe|Stall|Eligible|Executed Ipc|Warp Cycles Per|Active Warps|XU|Theoretical Occ|Achieved Occ|Registers" | cut -c1-260 | head -60
