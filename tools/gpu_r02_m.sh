#!/bin/bash
# round 2, call m: explicit shared-space accesses in gemm_tc (staging stores, side-sum tile reads), check-free IPE degree loop
mkdir -p gpurun_out
echo "=== pytest -m gpu (full)"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short -x 2>&1 | grep -vE "^\s*$" | tail -12 | tee gpurun_out/tests_m.log
echo "=== gemm_bench --bottleneck"; timeout 300 python tools/gemm_bench.py --bottleneck 2>&1 | tee gpurun_out/bottleneck_m.txt
echo "=== gemm_bench"; timeout 300 python tools/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench_m.txt
for v in 1 2; do
  echo "=== bench train360"; timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline 2>&1 | tail -1 | tee gpurun_out/bench_m$v.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(j['ms_per_step'],3), 'ms/step', round(j['value']), 'rays/s', j['clocks'], j['roofline']['frac'], j['roofline']['whole_step_frac'])"
done
echo "=== bench batch 2048"; timeout 300 python bench.py --steps 30 --warmup 5 --batch_size 2048 --no_cpu_baseline 2>&1 | tail -1 | tee gpurun_out/bench_b2048_m.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(j['ms_per_step'],3), 'ms/step', round(j['value']), 'rays/s', j['clocks'])"
echo "=== ncu launch list, batch 2048"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches_b2048.csv python bench.py --steps 1 --warmup 3 --batch_size 2048 --no_cpu_baseline --no_graph > gpurun_out/launches_b2048_run.log 2>&1
echo "=== ncu launch list, default"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/launches_run.log 2>&1
ls -la gpurun_out | head -30
