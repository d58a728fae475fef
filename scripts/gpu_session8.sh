#!/bin/bash
# Session 8: re-validation after container re-creation; fresh launch list + full captures of the current kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s8_smi.log 2>&1; nproc >> gpurun_out/s8_smi.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s8_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s8_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s8_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/s8_bench_default.json 2> gpurun_out/s8_bench_default.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/s8_bench_reference.json 2> gpurun_out/s8_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s8_launches_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s8_ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_simulate_cells -c 1 -f -o gpurun_out/s8_prof_chr1 python scripts/gpu_phases.py c3 148 1 > gpurun_out/s8_ncu_chr1.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_simulate_cells -c 1 -f -o gpurun_out/s8_prof_chr20 python scripts/gpu_phases.py c1 444 1 > gpurun_out/s8_ncu_chr20.log 2>&1
python scripts/gpu_phases.py c3 148 2 > gpurun_out/s8_phases_chr1.txt 2>&1
python scripts/gpu_phases.py c1 444 2 > gpurun_out/s8_phases_chr20.txt 2>&1
tail -3 gpurun_out/s8_pytest_gpu.log; tail -2 gpurun_out/s8_smoke.log; cat gpurun_out/s8_bench_default.json | cut -c1-600; cat gpurun_out/s8_bench_reference.json | cut -c1-300
