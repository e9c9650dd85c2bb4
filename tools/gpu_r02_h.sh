#!/bin/bash
mkdir -p gpurun_out
echo "=== tests"; timeout 600 python -m pytest -m gpu -q -p no:cacheprovider tests/test_gpu_chain.py 2>&1 | tail -3
echo "=== chain bench"; timeout 300 python tools/chain_bench.py 2>&1 | tail -6
echo "=== chain bench 8x256"; timeout 300 python tools/chain_bench.py --depth 8 --fpad 128 2>&1 | tail -6
MNRF_LIB=$PWD/multinerf_b200/libmnrf_b200_knobs.so timeout 300 python tools/chain_trace.py > gpurun_out/chain_trace_inf.txt 2>&1
MNRF_LIB=$PWD/multinerf_b200/libmnrf_b200_knobs.so timeout 300 python tools/chain_trace.py --train > gpurun_out/chain_trace_train.txt 2>&1
for w in train360 raw; do
  echo "=== bench $w"; timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no_cpu_baseline 2>&1 | tail -1 | tee gpurun_out/bench_$w.log | cut -c1-300
done
