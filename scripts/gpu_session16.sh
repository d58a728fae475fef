#!/bin/bash
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for c in burnin,sampling whole_c4,whole_pblock; do
  n=$(echo $c | tr ',' '_')
  timeout 1500 $CS --tool racecheck --racecheck-report analysis --print-limit 100000 python scripts/gpu_small_cases.py $c 2>&1 | grep -v "Host Frame\|Saved host\|^=========\s*$" | grep "Race reported\|Thread (\|RACECHECK\|epochs" | sed 's/modle_b200::CellSim:://; s/+0x[0-9a-f]*//' | cut -c1-230 | sort | uniq -c | sort -rn | head -60 > gpurun_out/s16_race_analysis_$n.txt
  echo "== $n"; cat gpurun_out/s16_race_analysis_$n.txt
done
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s16_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s16_pytest.log; tail -n 3 gpurun_out/s16_pytest.log
for l in 128 64 32; do
  MODLE_B200_GEN_PER_THREAD=$l timeout 600 python scripts/gpu_phases.py c3 148 2 2>&1 | grep -i "product\|rng_refill\|mv.ensure\|fault" | sed "s/^/[l=$l] /" >> gpurun_out/s16_gen.log
  MODLE_B200_GEN_PER_THREAD=$l timeout 600 python scripts/gpu_phases.py c1 444 2 2>&1 | grep -i "product\|rng_refill\|mv.ensure" | sed "s/^/[l=$l] /" >> gpurun_out/s16_gen.log
  MODLE_B200_GEN_PER_THREAD=$l timeout 600 python scripts/gpu_chrom.py chr13,chr8 512 2 2>&1 | sed "s/^/[l=$l] /" >> gpurun_out/s16_gen.log
done
cat gpurun_out/s16_gen.log | cut -c1-220
timeout 600 python scripts/bench_register.py --out gpurun_out/s16_register.json > gpurun_out/s16_register.log 2>&1
python -c "
import json
for r in json.load(open('gpurun_out/s16_register.json')): print(r['case'], r['stream'], '%.2f ms'%r['ms'], 'frac %.3f'%r['frac_of_hbm_sector_ceiling'])
"
for i in 1 2; do timeout 600 python scripts/gpu_determinism.py 8 warm > gpurun_out/s16_det_$i.log 2>&1; tail -n 3 gpurun_out/s16_det_$i.log | cut -c1-300; done
