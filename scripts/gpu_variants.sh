#!/bin/bash
# Compares library variants: per-phase cycle breakdown on C1 (chr20) and a chr1 sample.
mkdir -p gpurun_out
L=$PWD/modle_b200
{
for v in "" _ldcg _mb2; do
  for wl in "c1 512" "c3 296"; do
    MODLE_B200_LIB=$L/libmodle_b200$v.so timeout 300 python scripts/gpu_phases.py $wl 2 2>&1
  done
done
} > gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log
