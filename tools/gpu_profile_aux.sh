#!/bin/bash
# ncu full-set capture of the non-GEMM kernels of one train step (encode, heads, sampling, compositing).
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none \
  -k regex:'encode_fast_kernel|encode_kernel|head_fwd_kernel|pixels_to_rays|viewdir_enc|clip_adam|pack_weights|head_bwd_kernel|sample_level_kernel|composite' -s 44 -c 24 \
  -o gpurun_out/aux python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/aux_run.log 2>&1
tail -2 gpurun_out/aux_run.log | cut -c1-300
ls -la gpurun_out
