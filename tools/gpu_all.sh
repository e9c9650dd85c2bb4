#!/bin/bash
# tests + smoke + short bench on one B200
mkdir -p gpurun_out
bash tools/gpu_kernel_tests.sh > /dev/null 2>&1
tail -60 gpurun_out/kernel_tests.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
echo "=== bench"; timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 --cpu_rays 128 2>&1 | tail -5 | tee gpurun_out/bench.log
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-900 | tee gpurun_out/bench_ref.log
