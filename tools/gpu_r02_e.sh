#!/bin/bash
# round 2, run E: chain kernel with the pipelined epilogue
mkdir -p gpurun_out
run() { echo "=== $*" ; timeout 900 python -m pytest -m gpu -q -p no:cacheprovider --tb=short "$@" 2>&1 | grep -vE "^\s*$" | tail -${TAILN:-40}; }
{
TAILN=60 run tests/test_gpu_chain.py
TAILN=60 run tests/test_gpu_scripts.py
} > gpurun_out/tests_e.log 2>&1
grep -E "passed|failed|error|===|Error|assert" gpurun_out/tests_e.log | cut -c1-400 | tail -30
echo "=== chain bench"; timeout 300 python tools/chain_bench.py 2>&1 | tail -7
echo "=== chain bench 8x256 skip"; timeout 300 python tools/chain_bench.py --depth 8 --fpad 128 --only fwd 2>&1 | tail -3
for w in train360 raw; do
  echo "=== bench $w"; timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no_cpu_baseline 2>&1 | tail -1 | tee gpurun_out/bench_$w.log | cut -c1-400
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_chain_kernel -s 3 -c 1 \
  -o gpurun_out/chain2 python tools/chain_bench.py --only fwd --iters 2 > gpurun_out/chain_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
