#!/bin/bash
# round 2, call l: wgrad split heuristic, encode ray segments, per-ray view-direction rows, head_bwd flush; short-K GEMM variants A/B
mkdir -p gpurun_out
echo "=== pytest -m gpu (full)"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short -x 2>&1 | grep -vE "^\s*$" | tail -12 | tee gpurun_out/tests_l.log
for v in 0 5 4; do
  echo "=== gemm_bench --bottleneck SMALLK=$v"; MNRF_GEMM_SMALLK=$v timeout 300 python tools/gemm_bench.py --bottleneck 2>&1 | tee gpurun_out/bottleneck_smallk$v.txt
done
for v in 0 5 4 0 5 4; do
  echo "=== bench train360 SMALLK=$v"; MNRF_GEMM_SMALLK=$v timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline 2>&1 | tail -1 | tee gpurun_out/bench_smallk$v.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(j['ms_per_step'],3), 'ms/step', round(j['value']), 'rays/s', j['clocks'], j['roofline']['frac'], j['roofline']['whole_step_frac'])"
done
for v in 0 1; do
  echo "=== bench batch 2048 PDL=$v"; MNRF_PDL=$v timeout 300 python bench.py --steps 30 --warmup 5 --batch_size 2048 --no_cpu_baseline 2>&1 | tail -1 | tee gpurun_out/bench_b2048_l_pdl$v.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(j['ms_per_step'],3), 'ms/step', round(j['value']), 'rays/s', j['clocks'])"
done
echo "=== ncu launch list, batch 2048"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches_b2048.csv python bench.py --steps 1 --warmup 3 --batch_size 2048 --no_cpu_baseline --no_graph > gpurun_out/launches_b2048_run.log 2>&1
echo "=== ncu launch list, default"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/launches_run.log 2>&1
ls -la gpurun_out | head -30
