#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*" ; timeout 900 python -m pytest -m gpu -q -p no:cacheprovider --tb=short "$@" 2>&1 | grep -vE "^\s*$" | tail -${TAILN:-15}; }
run tests/test_gpu_chain.py tests/test_gpu_model.py tests/test_gpu_fullwidth.py tests/test_gpu_train_loop.py
echo "=== chain bench"; timeout 300 python tools/chain_bench.py 2>&1 | tail -6
for i in 1 2; do echo "=== bench train360"; timeout 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline 2>&1 | tail -1 | tee gpurun_out/bench_train360.log | cut -c1-250; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/launches_run.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
  --clock-control none -k regex:"gemm_tc_kernel|mlp_chain_kernel" -s 123 -c 41 --csv --log-file gpurun_out/gemm_traffic.csv \
  python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/traffic_run.log 2>&1
echo done
