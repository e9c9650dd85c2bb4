#!/bin/bash
mkdir -p gpurun_out
MNRF_LIB=$PWD/multinerf_b200/libmnrf_b200_knobs.so timeout 300 python tools/chain_trace.py > gpurun_out/chain_trace_inf.txt 2>&1
MNRF_LIB=$PWD/multinerf_b200/libmnrf_b200_knobs.so timeout 300 python tools/chain_trace.py --train > gpurun_out/chain_trace_train.txt 2>&1
head -120 gpurun_out/chain_trace_inf.txt
