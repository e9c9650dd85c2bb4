#!/bin/bash
# round 2 final evidence run on one B200: full GPU test suite, smoke, the four bench workloads + reference arm,
# ncu launch list, per-launch DRAM traffic of the tensor-core kernels, full-set capture, chain event traces
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short -x 2>&1 | grep -vE "^\s*$" | tail -15 | tee gpurun_out/tests_final.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
for w in train360 render refnerf raw; do
  echo "=== bench $w"; timeout 900 python bench.py --workload $w --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_$w.log | cut -c1-300
done
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.log | cut -c1-300
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/launches_run.log 2>&1
echo "=== ncu launch list, side sums as separate passes (A/B)"
MNRF_SIDE_SUMS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches_noside.csv python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > /dev/null 2>&1
echo "=== bench train360 A/B same box"
for v in 1 0 1 0; do MNRF_SIDE_SUMS=$v timeout 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('side_sums=$v', round(j['ms_per_step'],3), 'ms/step', j['clocks'])"; done
echo "=== ncu per-launch traffic of the tensor-core kernels (4th step)"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
  --clock-control none -k regex:"gemm_tc_kernel|mlp_chain_kernel" -s 123 -c 41 --csv --log-file gpurun_out/gemm_traffic.csv \
  python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/traffic_run.log 2>&1
echo "=== ncu full-set: chain fwd/bwd + 8 GEMMs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|mlp_chain_kernel" -s 130 -c 12 \
  -o gpurun_out/r02_tc python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/full_run.log 2>&1
if [ -f multinerf_b200/libmnrf_b200_knobs.so ]; then
  MNRF_LIB=$PWD/multinerf_b200/libmnrf_b200_knobs.so timeout 300 python tools/chain_trace.py > gpurun_out/chain_trace_inf.txt 2>&1
  MNRF_LIB=$PWD/multinerf_b200/libmnrf_b200_knobs.so timeout 300 python tools/chain_trace.py --train > gpurun_out/chain_trace_train.txt 2>&1
fi
timeout 200 python tools/chain_bench.py > gpurun_out/chain_bench.txt 2>&1; timeout 200 python tools/chain_bench.py --depth 8 --fpad 128 >> gpurun_out/chain_bench.txt 2>&1; cat gpurun_out/chain_bench.txt
ls -la gpurun_out | head -40
