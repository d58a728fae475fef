#!/bin/bash
mkdir -p gpurun_out
for set in chr1,chr15,chr21 chr2,chr14,chr22 chr7,chr9,chr16; do
  n=$(echo $set | tr ',' '_')
  MODLE_B200_BENCH_CHROMS=$set timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --streams 3 > gpurun_out/s21_$n.json 2> gpurun_out/s21_$n.err
  echo "== $set rc=$?"; grep -h "Error\|error" gpurun_out/s21_$n.err | head -3 | cut -c1-250; cut -c1-150 gpurun_out/s21_$n.json | tail -n 1
done
MODLE_B200_BENCH_CHROMS=chr1,chr15,chr21 timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 8 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --streams 3 --cells 6 > gpurun_out/s21_memcheck.log 2>&1
grep -v "Host Frame" gpurun_out/s21_memcheck.log | grep -v "^W10\|^\[W" | head -60 | cut -c1-220
