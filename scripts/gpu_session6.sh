#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "register" 2>&1 | tail -5
for path in direct binned; do
MODLE_B200_REGISTER_PATH=$path timeout 600 python scripts/bench_register.py --reps 3 --out gpurun_out/register_$path.json > gpurun_out/register_$path.log 2>&1
grep c5_hbm gpurun_out/register_$path.log | cut -c1-330
done
