#!/bin/bash
# round 2, call o: branch-free large-argument sin in the IPE loop -- parity tests, step time, launch list
mkdir -p gpurun_out
echo "=== pytest kernels + model"; timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullwidth.py tests/test_gpu_model.py -m gpu -q -p no:cacheprovider --tb=short -x 2>&1 | grep -vE "^\s*$" | tail -8 | tee gpurun_out/tests_o.log
echo "=== bench train360"; timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline 2>&1 | tail -1 | tee gpurun_out/bench_o.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(j['ms_per_step'],3), 'ms/step', round(j['value']), 'rays/s', j['clocks'], j['roofline']['frac'], j['roofline']['whole_step_frac'])"
echo "=== bench batch 2048"; timeout 300 python bench.py --steps 30 --warmup 5 --batch_size 2048 --no_cpu_baseline 2>&1 | tail -1 | tee gpurun_out/bench_b2048_o.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(j['ms_per_step'],3), 'ms/step', round(j['value']), 'rays/s', j['clocks'])"
echo "=== ncu launch list, batch 2048"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches_b2048.csv python bench.py --steps 1 --warmup 3 --batch_size 2048 --no_cpu_baseline --no_graph > gpurun_out/launches_b2048_run.log 2>&1
python - <<'P'
import csv
lines=[l for l in open('gpurun_out/launches_b2048.csv') if not l.startswith('==')]
rows=[(r['Kernel Name'], float(r['Metric Value'].replace(',',''))) for r in csv.DictReader(lines) if r.get('Metric Name')=='gpu__time_duration.sum']
idx=[i for i,(n,_) in enumerate(rows) if 'sample_level' in n][::3]
st=rows[idx[-2]:idx[-1]]
print('b2048 step total ms', sum(t for _,t in st)/1e6, 'encode us', [round(t/1e3,1) for n,t in st if 'encode' in n])
P
