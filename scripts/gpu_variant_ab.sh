#!/bin/bash
# A/B of a variant library against the product on the chr1 / C4 shapes (both RNG modes).
#   bash scripts/gpu_variant_ab.sh <variant> <tag>
V=$1; TAG=${2:-ab}
mkdir -p gpurun_out
for lib in "" $PWD/modle_b200/libmodle_b200_$V.so; do
  for mode in 0 1; do
    for wl in "c3 148" "c4 148"; do
      MODLE_B200_LIB=$lib MODLE_B200_RNG_MODE=$mode timeout 300 python scripts/gpu_phases.py $wl 2>&1 | grep -E "product|libmodle" | sed "s/^/mode $mode: /"
    done
  done
done | tee gpurun_out/${TAG}_variant_$V.txt
