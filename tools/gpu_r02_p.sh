#!/bin/bash
# round 2, call p: training on the procedural scene with the final kernels (convergence / r/s log); ncu of the final encode kernel
mkdir -p gpurun_out
timeout 200 python tools/train_synthetic.py --steps 400 --batch 16384 --cast_rays --graph > gpurun_out/train_synthetic.log 2>&1; tail -14 gpurun_out/train_synthetic.log
timeout 200 ncu --set full --import-source on --clock-control none -k regex:encode_fast_kernel -s 9 -c 1 -o gpurun_out/enc2 \
  python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/enc2_run.log 2>&1
ls -la gpurun_out/enc2*
