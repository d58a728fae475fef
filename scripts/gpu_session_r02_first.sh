#!/bin/bash
# First GPU session of round 2: everything built after round 1's GPU budget ran out.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_session_r02_first.sh'
# 1. the whole -m gpu suite (the throughput-mode tests and the statistical gate have never run on
#    a device; the deterministic parity tests re-validate the four barrier moves of DESIGN.md 3)
# 2. bench lines: deterministic (headline) and throughput mode, same workload
# 3. per-phase cycles of both modes on the chr1 and chr20 shapes
# 4. racecheck of the throughput-mode kernels on a small case; determinism repeats
# 5. ncu launch list + one full capture of the throughput-mode <1024,1> kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02a_pytest.log
timeout 300 python bench.py --steps 2 --warmup 3 > gpurun_out/r02a_bench_det.json 2> gpurun_out/r02a_bench_det.err; echo "bench det rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --rng-mode throughput --no-cpu-baseline > gpurun_out/r02a_bench_thr.json 2> gpurun_out/r02a_bench_thr.err; echo "bench thr rc=$?"
cut -c1-220 gpurun_out/r02a_bench_det.json gpurun_out/r02a_bench_thr.json
for mode in 0 1; do
  (MODLE_B200_RNG_MODE=$mode timeout 200 python scripts/gpu_phases.py c3 148; MODLE_B200_RNG_MODE=$mode timeout 200 python scripts/gpu_phases.py c1 444) > gpurun_out/r02a_phases_mode$mode.txt 2>&1; echo "phases mode $mode rc=$?"; grep product gpurun_out/r02a_phases_mode$mode.txt
done
MODLE_B200_RNG_MODE=1 timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --print-limit 20 \
  python scripts/gpu_small_cases.py burnin,sampling,whole_small,whole_c4,whole_pblock,mid,large > gpurun_out/r02a_racecheck_thr.log 2>&1; echo "racecheck rc=$?"; grep -c "Race reported" gpurun_out/r02a_racecheck_thr.log
MODLE_B200_RNG_MODE=1 timeout 300 python scripts/gpu_determinism.py > gpurun_out/r02a_determinism_thr.log 2>&1; echo "determinism rc=$?"; tail -2 gpurun_out/r02a_determinism_thr.log
MODLE_B200_BENCH_CHROMS=chr1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_simulate_cells -c 1 \
  -o gpurun_out/r02a_ncu_thr_chr1 python bench.py --steps 1 --warmup 0 --cells 148 --rng-mode throughput --no-cpu-baseline --streams 1 > gpurun_out/r02a_ncu_thr.log 2>&1; echo "ncu rc=$?"
# 6. A/B of the deterministic mode with the window repair of rank_lefs (sim_core.hpp switch
#    MODLE_B200_WINDOW_RANK_REPAIR=1; bit-identical results, CPU-verified): build the variant
#    BEFORE the gpurun call (python -c "from modle_b200 import build; build.build(variant='winrank',
#    defines=['MODLE_B200_WINDOW_RANK_REPAIR=1'])"), then parity + phases + bench through it.
if [ -f modle_b200/libmodle_b200_winrank.so ]; then
  export MODLE_B200_LIB=$PWD/modle_b200/libmodle_b200_winrank.so
  timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02a_pytest_winrank.log 2>&1; echo "winrank parity rc=$?"; tail -2 gpurun_out/r02a_pytest_winrank.log
  (timeout 200 python scripts/gpu_phases.py c3 148; timeout 200 python scripts/gpu_phases.py c1 444) > gpurun_out/r02a_phases_winrank.txt 2>&1; echo "winrank phases rc=$?"; grep product gpurun_out/r02a_phases_winrank.txt
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_winrank.json 2> gpurun_out/r02a_bench_winrank.err; echo "bench winrank rc=$?"; cut -c1-220 gpurun_out/r02a_bench_winrank.json
  unset MODLE_B200_LIB
fi
