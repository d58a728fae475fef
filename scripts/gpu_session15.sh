#!/bin/bash
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for c in burnin,sampling whole_small whole_c4,whole_pblock mid large; do
  n=$(echo $c | tr ',' '_')
  timeout 1500 $CS --tool racecheck --racecheck-report all --print-limit 1000000 python scripts/gpu_small_cases.py $c 2>&1 | grep "Thread (\|RACECHECK SUMMARY\|epochs" | grep -o "in [a-z_]*\.[a-z]*:[0-9]*\|RACECHECK.*\|^[a-z_0-9]* epochs.*" | paste -sd' \n' > gpurun_out/s15_race_$n.raw
  grep -o "in [a-z_.]*:[0-9]* in [a-z_.]*:[0-9]*" gpurun_out/s15_race_$n.raw | sort | uniq -c > gpurun_out/s15_race_$n.txt
  grep -o "RACECHECK.*" gpurun_out/s15_race_$n.raw >> gpurun_out/s15_race_$n.txt
  rm -f gpurun_out/s15_race_$n.raw
  echo "== $n"; cat gpurun_out/s15_race_$n.txt
done
for i in 1 2 3; do timeout 600 python scripts/gpu_determinism.py 8 warm > gpurun_out/s15_det_$i.log 2>&1; tail -n 4 gpurun_out/s15_det_$i.log | cut -c1-300; done
