#!/bin/bash
# round 2, run C: where does the chained-trunk kernel spend its time?
mkdir -p gpurun_out
echo "=== product build"; timeout 300 python tools/chain_bench.py 2>&1 | tail -8
echo "=== knob build"
MNRF_TIMING_KNOBS=1 python multinerf_b200/build.py > /dev/null 2>&1
for dbg in 0 1 2 4 6 7; do
  echo "--- MNRF_CHAIN_DEBUG=$dbg"; MNRF_CHAIN_DEBUG=$dbg timeout 300 python tools/chain_bench.py --only fwd 2>&1 | tail -2
done
MNRF_CHAIN_DEBUG=1 timeout 300 python tools/chain_bench.py --only bwd 2>&1 | tail -2
python multinerf_b200/build.py > /dev/null 2>&1
echo "=== ncu full capture of the forward chain"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_chain_kernel -s 3 -c 2 \
  -o gpurun_out/chain python tools/chain_bench.py --only fwd --iters 2 > gpurun_out/chain_ncu.log 2>&1
tail -3 gpurun_out/chain_ncu.log
ls -la gpurun_out/*.ncu-rep
