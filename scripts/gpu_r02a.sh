#!/bin/bash
# Round 2, first GPU call: suite, both RNG modes timed, phases, window-rank A/B.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02a_pytest.log
timeout 300 python bench.py --steps 2 --warmup 3 > gpurun_out/r02a_bench_det.json 2> gpurun_out/r02a_bench_det.err; echo "bench det rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --rng-mode throughput --no-cpu-baseline > gpurun_out/r02a_bench_thr.json 2> gpurun_out/r02a_bench_thr.err; echo "bench thr rc=$?"
cut -c1-220 gpurun_out/r02a_bench_det.json gpurun_out/r02a_bench_thr.json
for mode in 0 1; do
  (MODLE_B200_RNG_MODE=$mode timeout 200 python scripts/gpu_phases.py c3 148; MODLE_B200_RNG_MODE=$mode timeout 200 python scripts/gpu_phases.py c1 444) > gpurun_out/r02a_phases_mode$mode.txt 2>&1; echo "phases mode $mode rc=$?"; grep product gpurun_out/r02a_phases_mode$mode.txt
done
if [ -f modle_b200/libmodle_b200_winrank.so ]; then
  export MODLE_B200_LIB=$PWD/modle_b200/libmodle_b200_winrank.so
  timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02a_pytest_winrank.log 2>&1; echo "winrank parity rc=$?"; tail -2 gpurun_out/r02a_pytest_winrank.log
  (timeout 200 python scripts/gpu_phases.py c3 148; timeout 200 python scripts/gpu_phases.py c1 444) > gpurun_out/r02a_phases_winrank.txt 2>&1; echo "winrank phases rc=$?"; grep winrank gpurun_out/r02a_phases_winrank.txt
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_winrank.json 2> gpurun_out/r02a_bench_winrank.err; echo "bench winrank rc=$?"; cut -c1-220 gpurun_out/r02a_bench_winrank.json
  unset MODLE_B200_LIB
fi
