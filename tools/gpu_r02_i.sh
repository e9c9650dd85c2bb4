#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*" ; timeout 900 python -m pytest -m gpu -q -p no:cacheprovider --tb=short "$@" 2>&1 | grep -vE "^\s*$" | tail -${TAILN:-25}; }
{
run tests/test_gpu_kernels.py -k "wgrad or dgrad"
run tests/test_gpu_chain.py
TAILN=40 run tests/test_gpu_model.py
TAILN=30 run tests/test_gpu_fullwidth.py -s
} > gpurun_out/tests_i.log 2>&1
grep -E "passed|failed|error|===|fullwidth|Error|assert" gpurun_out/tests_i.log | cut -c1-500 | tail -30
echo "=== chain bench"; timeout 300 python tools/chain_bench.py --only bwd 2>&1 | tail -3
for w in train360 raw refnerf; do
  echo "=== bench $w"; timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no_cpu_baseline 2>&1 | tail -1 | tee gpurun_out/bench_$w.log | cut -c1-300
done
