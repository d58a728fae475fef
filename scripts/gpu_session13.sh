#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/gpu_determinism.py 8 > gpurun_out/s13_det_plain.log 2>&1
timeout 900 python scripts/gpu_determinism.py 8 warm > gpurun_out/s13_det_warm.log 2>&1
MODLE_B200_MID=0 timeout 900 python scripts/gpu_determinism.py 8 warm > gpurun_out/s13_det_mid0.log 2>&1
tail -n 12 gpurun_out/s13_det_plain.log gpurun_out/s13_det_warm.log gpurun_out/s13_det_mid0.log | cut -c1-400
