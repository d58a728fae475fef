#!/bin/bash
# GPU session: parity tests, smoke, bench, ncu launch list + full capture of the top kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.log 2>&1
nproc >> gpurun_out/smi.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --workload c1 --steps 2 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c2_short.json 2> gpurun_out/bench_c2_short.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c1.csv python bench.py --workload c1 --cells 148 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_simulate_cells -c 1 -o gpurun_out/prof_c1 python bench.py --workload c1 --cells 148 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/bench_c1.json; cat gpurun_out/bench_c2_short.json
