#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
echo "=== render check x$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/render_check.py 2>&1 | grep -vE "^\s*$|Warning|warn" | tail -12
echo "=== bench x$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_$N.log
echo "=== bench x1"
timeout 900 python bench.py --steps 10 --warmup 3 --no_cpu_baseline 2>&1 | tail -1 | tee gpurun_out/bench_1.log
