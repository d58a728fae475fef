#!/bin/bash
# The standard single-GPU check after a kernel change (one gpurun call, ~3 minutes):
#   gpurun --timeout 1200 -- 'bash scripts/gpu_suite.sh <tag>'
# 1. the whole `-m gpu` suite (parity vs the oracle, both RNG modes, full-geometry C4 / C5, ...)
# 2. per-phase cycles of both modes on the chr1 (<1024,1>) and chr20 (<256,3>) shapes
# 3. bench lines of both modes (short: 2 steps; the driver's run is the record)
TAG=${1:-suite}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
for mode in 0 1; do
  (MODLE_B200_RNG_MODE=$mode timeout 200 python scripts/gpu_phases.py c3 148; MODLE_B200_RNG_MODE=$mode timeout 200 python scripts/gpu_phases.py c1 444) > gpurun_out/${TAG}_phases_mode$mode.txt 2>&1; echo "phases mode $mode rc=$?"; grep product gpurun_out/${TAG}_phases_mode$mode.txt
done
timeout 300 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench_det.json 2> gpurun_out/${TAG}_bench_det.err; echo "bench det rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --rng-mode throughput > gpurun_out/${TAG}_bench_thr.json 2> gpurun_out/${TAG}_bench_thr.err; echo "bench thr rc=$?"
cut -c1-200 gpurun_out/${TAG}_bench_det.json gpurun_out/${TAG}_bench_thr.json
