#!/bin/bash
# Session 9: band -> pixels kernels (tests, bench, ncu) + per-kernel times of the binned register path at full density.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pixels.py -m gpu -x -q > gpurun_out/s9_pytest_pixels.log 2>&1; echo "rc=$?" >> gpurun_out/s9_pytest_pixels.log
timeout 900 python scripts/bench_pixels.py --out gpurun_out/s9_pixels.json > gpurun_out/s9_pixels.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fill_pixels|k_count_row" -c 2 -f -o gpurun_out/s9_prof_pixels python scripts/bench_pixels.py --reps 0 --no-cpu --cases chr1_loop > gpurun_out/s9_ncu_pixels.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum --clock-control none -k regex:"k_bin|k_scatter|k_register" -c 40 --csv --log-file gpurun_out/s9_register_launches.csv python scripts/bench_register.py --reps 0 > gpurun_out/s9_register_ncu.log 2>&1
timeout 600 python scripts/bench_register.py --out gpurun_out/s9_register.json > gpurun_out/s9_register.log 2>&1
tail -3 gpurun_out/s9_pytest_pixels.log; cat gpurun_out/s9_pixels.log | cut -c1-700; tail -4 gpurun_out/s9_register.log | cut -c1-400
