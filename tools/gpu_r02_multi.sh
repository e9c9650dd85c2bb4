#!/bin/bash
# round 2: N-GPU checks (N = $1): render path equivalence, train-step bench with the gradient exchange captured
# in the step graph (and the two-graph fallback), render bench
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== render check x$N"
timeout 600 $TR --master-port 29511 tools/render_check.py 2>&1 | grep -vE "^\s*$|Warning|warn" | tail -8 | tee gpurun_out/render_check_$N.log
echo "=== bench train360 x$N (one graph, NCCL captured, NerfMLP exchange overlapped)"
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_train360_$N.log | cut -c1-600
echo "=== bench train360 x$N MNRF_GRAPH_NCCL=0 (two graphs around an eager all-reduce)"
MNRF_GRAPH_NCCL=0 timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_train360_${N}_twograph.log | cut -c1-300
echo "=== bench render x$N"
timeout 900 $TR --master-port 29514 bench.py --gpus $N --workload render --steps 5 --warmup 2 2>&1 | tail -1 | tee gpurun_out/bench_render_$N.log | cut -c1-900
echo "=== bench render x$N --no_graph"
timeout 900 $TR --master-port 29515 bench.py --gpus $N --workload render --steps 5 --warmup 2 --no_graph 2>&1 | tail -1 | cut -c1-300
