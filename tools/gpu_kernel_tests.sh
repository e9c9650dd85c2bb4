#!/bin/bash
# GPU run of the kernel + model tests; GEMM groups in separate processes so a trap in one
# tcgen05 variant does not take the other results down with it.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { echo "=== $*" ; timeout 900 python -m pytest -m gpu -q -p no:cacheprovider --tb=short "$@" 2>&1 | grep -vE "^\s*$" | tail -${TAILN:-60}; }
{
run tests/test_gpu_kernels.py -k "not gemm"
run tests/test_gpu_kernels.py -k "gemm"
TAILN=120 run tests/test_gpu_model.py
run tests/test_gpu_camera.py
run tests/test_gpu_train_loop.py
} > gpurun_out/kernel_tests.log 2>&1
tail -250 gpurun_out/kernel_tests.log
