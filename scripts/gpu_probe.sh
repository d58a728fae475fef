#!/bin/bash
# Finer phase split of the throughput-mode kernel (variant library built with -DMODLE_B200_PROBE=1:
#   python -c "from modle_b200 import build; build.build(variant='probe', defines=['MODLE_B200_PROBE=1'])").
# Slots borrowed in that build: rng_refill = rank: slots + count scan; mv.ensure = rank: merge;
# mv.scan = rank: verify sweep (the rest of `rank` = repair rounds); mv.exceptions = collisions:
# clearing the collision words; sec.scan = boundaries (leader); sec.draws = LEF-BAR walk.
TAG=${1:-probe}
mkdir -p gpurun_out
export MODLE_B200_LIB=$PWD/modle_b200/libmodle_b200_probe.so
(MODLE_B200_RNG_MODE=1 timeout 200 python scripts/gpu_phases.py c3 148; MODLE_B200_RNG_MODE=1 timeout 200 python scripts/gpu_phases.py c1 444) > gpurun_out/${TAG}_phases_probe.txt 2>&1; echo "rc=$?"; cat gpurun_out/${TAG}_phases_probe.txt
