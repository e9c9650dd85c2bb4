#!/bin/bash
# round 2, run A: all GPU tests, smoke, the four bench workloads, CPU thread sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { echo "=== $*" ; timeout 1200 python -m pytest -m gpu -q -p no:cacheprovider --tb=short "$@" 2>&1 | grep -vE "^\s*$" | tail -${TAILN:-40}; }
{
run tests/test_gpu_kernels.py -k "not gemm"
run tests/test_gpu_kernels.py -k "gemm"
TAILN=80 run tests/test_gpu_model.py
TAILN=80 run tests/test_gpu_fullwidth.py -s
run tests/test_gpu_render.py
run tests/test_gpu_camera.py
run tests/test_gpu_train_loop.py
} > gpurun_out/tests_a.log 2>&1
grep -E "passed|failed|error|===" gpurun_out/tests_a.log | tail -30
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
for w in train360 render refnerf raw; do
  echo "=== bench $w"; timeout 900 python bench.py --workload $w --steps ${STEPS:-10} --warmup 3 --no_cpu_baseline 2>&1 | tail -1 | tee gpurun_out/bench_$w.log | cut -c1-1500
done
echo "=== cpu sweep"; timeout 900 python tools/cpu_sweep.py 1024 16 32 64 128 2>&1 | tail -6
