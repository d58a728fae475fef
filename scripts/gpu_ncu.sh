#!/bin/bash
# ncu captures behind profiles/ (one gpurun call on ONE GPU):
#   gpurun --timeout 2400 -- 'bash scripts/gpu_ncu.sh <tag>'
# * one full capture (--set full, source imported) of k_simulate_cells<1024,1> per RNG mode:
#   chr1 shape, 148 cells, one launch  ->  gpurun_out/<tag>_ncu_chr1_mode{0,1}.ncu-rep
#   (read here with `ncu -i ... --page raw --csv` / `--page source --csv`, scripts/ncu_lines.py)
# * the launch list of the default bench command (gpu__time_duration per launch)
# * per-kernel DRAM bytes / reduction sectors of the isolated contact-register path (C5 geometry)
# * compute-sanitizer racecheck of both modes over scripts/gpu_small_cases.py
TAG=${1:-ncu}
mkdir -p gpurun_out
for mode in 0 1; do
  MODLE_B200_BENCH_CHROMS=chr1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_simulate_cells -c 1 \
    -o gpurun_out/${TAG}_ncu_chr1_mode$mode python bench.py --steps 1 --warmup 0 --cells 148 --rng-mode $([ $mode = 1 ] && echo throughput || echo deterministic) --no-cpu-baseline --no-extras --streams 1 > gpurun_out/${TAG}_ncu_mode$mode.log 2>&1; echo "ncu mode $mode rc=$?"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_default_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum,l1tex__t_requests_pipe_lsu_mem_global_op_red.sum --clock-control none -k regex:"k_bin|k_scatter|k_register|k_calibrate" --csv --log-file gpurun_out/${TAG}_register_ncu.csv python scripts/bench_register.py --reps 1 > gpurun_out/${TAG}_register_ncu.log 2>&1; echo "register ncu rc=$?"
for mode in 0 1; do
  MODLE_B200_RNG_MODE=$mode timeout 500 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --print-limit 20 \
    python scripts/gpu_small_cases.py burnin,sampling,whole_small,whole_c4,whole_pblock,mid,large > gpurun_out/${TAG}_racecheck_mode$mode.log 2>&1; echo "racecheck mode $mode rc=$?"; grep -c "Race reported" gpurun_out/${TAG}_racecheck_mode$mode.log; tail -1 gpurun_out/${TAG}_racecheck_mode$mode.log
done
