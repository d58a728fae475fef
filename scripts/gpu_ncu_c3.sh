#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_simulate_cells -c 1 -f -o gpurun_out/prof_c3 python scripts/gpu_phases.py c3 148 1 > gpurun_out/ncu_c3.log 2>&1
tail -5 gpurun_out/ncu_c3.log
