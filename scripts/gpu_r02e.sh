#!/bin/bash
TAG=${1:-r02e}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_throughput_mode.py tests/test_gpu_full_geometry.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
for mode in 0 1; do
  (MODLE_B200_RNG_MODE=$mode timeout 200 python scripts/gpu_phases.py c3 148; MODLE_B200_RNG_MODE=$mode timeout 200 python scripts/gpu_phases.py c1 444) > gpurun_out/${TAG}_phases_mode$mode.txt 2>&1; echo "phases mode $mode rc=$?"; grep product gpurun_out/${TAG}_phases_mode$mode.txt
done
MODLE_B200_BENCH_CHROMS=chr1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_simulate_cells -c 1 \
    -o gpurun_out/${TAG}_ncu_chr1_mode1 python bench.py --steps 1 --warmup 0 --cells 148 --rng-mode throughput --no-cpu-baseline --no-extras --streams 1 > gpurun_out/${TAG}_ncu_mode1.log 2>&1; echo "ncu rc=$?"
