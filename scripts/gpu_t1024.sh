#!/bin/bash
mkdir -p gpurun_out
L=$PWD/modle_b200/libmodle_b200_t1024.so
{
MODLE_B200_LIB=$L timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c3_chr1_shape or c4_high or c1_chr20" 2>&1 | tail -3
MODLE_B200_LIB=$L timeout 300 python scripts/gpu_phases.py c3 296 2
MODLE_B200_LIB=$L timeout 300 python scripts/gpu_phases.py c4 296 1
} > gpurun_out/t1024.log 2>&1
cat gpurun_out/t1024.log
