#!/bin/bash
# round 2, run F: chain hand-off with CTA-scope arrive; epilogue variants (v1 = x32 single buffer, default = x16 pipelined)
mkdir -p gpurun_out
which nvcc; ls /usr/local/cuda/bin/nvcc
echo "=== default (x16 pipelined)"; timeout 300 python tools/chain_bench.py 2>&1 | tail -6
echo "=== v1 (x32 single-buffer)"; MNRF_LIB=$PWD/multinerf_b200/libmnrf_b200_v1.so timeout 300 python tools/chain_bench.py 2>&1 | tail -6
echo "=== tests (default)"; timeout 600 python -m pytest -m gpu -q -p no:cacheprovider tests/test_gpu_chain.py 2>&1 | tail -3
echo "=== tests (v1)"; MNRF_LIB=$PWD/multinerf_b200/libmnrf_b200_v1.so timeout 600 python -m pytest -m gpu -q -p no:cacheprovider tests/test_gpu_chain.py 2>&1 | tail -3
for w in train360 raw; do
  echo "=== bench $w"; timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no_cpu_baseline 2>&1 | tail -1 | tee gpurun_out/bench_$w.log | cut -c1-300
done
