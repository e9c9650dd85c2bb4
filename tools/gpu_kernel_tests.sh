#!/bin/bash
# First-contact GPU run: every kernel test, GEMM groups in separate processes so a trap in one
# tcgen05 variant does not take the other results down with it.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { echo "=== $*" ; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider "$@" 2>&1 | tail -40; }
{
run -k "not gemm"
run -k "gemm_fwd"
run -k "gemm_dgrad"
run -k "gemm_wgrad"
} > gpurun_out/kernel_tests.log 2>&1
tail -150 gpurun_out/kernel_tests.log
