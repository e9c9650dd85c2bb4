#!/bin/bash
# round 2, run B: the layer-chained trunk kernel -- unit tests, model tests, benches, launch list
mkdir -p gpurun_out
run() { echo "=== $*" ; timeout 900 python -m pytest -m gpu -q -p no:cacheprovider --tb=short "$@" 2>&1 | grep -vE "^\s*$" | tail -${TAILN:-40}; }
{
TAILN=60 run tests/test_gpu_chain.py -x
TAILN=60 run tests/test_gpu_fullwidth.py -s
run tests/test_gpu_model.py
run tests/test_gpu_render.py tests/test_gpu_train_loop.py
} > gpurun_out/tests_b.log 2>&1
grep -E "passed|failed|error|===|fullwidth" gpurun_out/tests_b.log | cut -c1-600 | tail -30
for w in train360 raw refnerf; do
  echo "=== bench $w"; timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no_cpu_baseline 2>&1 | tail -1 | tee gpurun_out/bench_$w.log | cut -c1-900
done
echo "=== bench train360 MNRF_CHAIN=0"; MNRF_CHAIN=0 timeout 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline 2>&1 | tail -1 | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no_cpu_baseline --no_graph > gpurun_out/launches_run.log 2>&1
python tools/summarize_profile.py r02_chain1 2>/dev/null | head -40
echo "=== cpu sweep"; timeout 600 python tools/cpu_sweep.py 1024 4 8 12 16 24 2>&1 | tail -6
