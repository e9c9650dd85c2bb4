#!/bin/bash
# round 2: N-GPU checks (N = $1): render path equivalence, train-step bench (graphs around the NCCL exchange, the
# NerfMLP segment reduced while the PropMLP levels are in backward), render bench.  Tight timeouts: a hang costs N x.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== render check x$N"
timeout 240 $TR --master-port 29511 tools/render_check.py 2>&1 | grep -vE "^\s*$|Warning|warn|OMP_NUM|\*\*\*" | tail -6 | tee gpurun_out/render_check_$N.log
echo "=== bench train360 x$N"
timeout 240 $TR --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_train360_$N.log | cut -c1-2500
echo "=== bench render x$N"
timeout 300 $TR --master-port 29514 bench.py --gpus $N --workload render --steps 5 --warmup 2 2>&1 | tail -1 | tee gpurun_out/bench_render_$N.log | cut -c1-600
if [ "$2" == "nccl_graph" ]; then
  echo "=== bench train360 x$N MNRF_GRAPH_NCCL=1"
  MNRF_GRAPH_NCCL=1 timeout 120 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | tail -2 | cut -c1-300
fi
