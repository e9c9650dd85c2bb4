#!/bin/bash
# round 2 final evidence run on one B200 (second session): full GPU test suite, smoke, the four bench workloads + reference arm,
# per-launch DRAM traffic of the tensor-core kernels, full-set capture of chain fwd/bwd + GEMMs, chain micro-bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short -x 2>&1 | grep -vE "^\s*$" | tail -8 | tee gpurun_out/tests_final.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke_final.log
for w in train360 render refnerf raw; do
  echo "=== bench $w"; timeout 600 python bench.py --workload $w --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_$w.log | cut -c1-400
done
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.log | cut -c1-300
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/launches_run.log 2>&1
echo "=== ncu per-launch traffic of the tensor-core kernels (4th step)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
  --clock-control none -k regex:"gemm_tc_kernel|mlp_chain_kernel" -s 123 -c 41 --csv --log-file gpurun_out/gemm_traffic.csv \
  python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/traffic_run.log 2>&1
echo "=== ncu full-set: chain fwd/bwd + GEMMs of the NerfMLP level"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|mlp_chain_kernel" -s 130 -c 10 \
  -o gpurun_out/r02_tc python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/full_run.log 2>&1
timeout 200 python tools/chain_bench.py > gpurun_out/chain_bench.txt 2>&1; tail -6 gpurun_out/chain_bench.txt
timeout 200 python tools/gemm_bench.py --bottleneck > gpurun_out/bottleneck_final.txt 2>&1; cat gpurun_out/bottleneck_final.txt
ls -la gpurun_out | head -40
