#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/psnr_parity.py > gpurun_out/psnr_parity.txt 2>&1; cat gpurun_out/psnr_parity.txt | tail -20
