#!/bin/bash
# quick GPU check: non-GEMM kernel tests, model tests, short bench
mkdir -p gpurun_out
timeout 900 python -m pytest -m gpu -q -p no:cacheprovider --tb=short tests/test_gpu_kernels.py -k "not gemm" 2>&1 | tail -15
timeout 900 python -m pytest -m gpu -q -p no:cacheprovider --tb=short tests/test_gpu_model.py 2>&1 | tail -15
timeout 600 python bench.py --steps ${STEPS:-20} --warmup 3 --no_cpu_baseline 2>&1 | tail -1 | cut -c1-1400
