#!/bin/bash
# round 2, run D: chain kernel after the hand-off change; entry-script tests
mkdir -p gpurun_out
run() { echo "=== $*" ; timeout 900 python -m pytest -m gpu -q -p no:cacheprovider --tb=short "$@" 2>&1 | grep -vE "^\s*$" | tail -${TAILN:-40}; }
{
TAILN=60 run tests/test_gpu_chain.py
TAILN=60 run tests/test_gpu_scripts.py
TAILN=30 run tests/test_gpu_fullwidth.py -s
run tests/test_gpu_train_loop.py tests/test_gpu_render.py
} > gpurun_out/tests_d.log 2>&1
grep -E "passed|failed|error|===|fullwidth|Error|assert" gpurun_out/tests_d.log | cut -c1-400 | tail -40
echo "=== chain bench"; timeout 300 python tools/chain_bench.py 2>&1 | tail -7
for w in train360 raw; do
  echo "=== bench $w"; timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no_cpu_baseline 2>&1 | tail -1 | tee gpurun_out/bench_$w.log | cut -c1-700
done
