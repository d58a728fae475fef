#!/bin/bash
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $CS --tool racecheck --racecheck-report all --print-limit 60 python scripts/gpu_small_cases.py burnin,sampling > gpurun_out/s14_racecheck_small.log 2>&1
timeout 1200 $CS --tool racecheck --racecheck-report all --print-limit 60 python scripts/gpu_small_cases.py mid > gpurun_out/s14_racecheck_mid.log 2>&1
timeout 900 $CS --tool initcheck --print-limit 30 python scripts/gpu_small_cases.py burnin,sampling > gpurun_out/s14_initcheck.log 2>&1
timeout 900 $CS --tool memcheck --print-limit 30 python scripts/gpu_small_cases.py burnin,sampling,mid > gpurun_out/s14_memcheck.log 2>&1
for f in racecheck_small racecheck_mid initcheck memcheck; do echo "== $f"; grep -c "=========" gpurun_out/s14_$f.log; tail -n 6 gpurun_out/s14_$f.log | cut -c1-250; done
