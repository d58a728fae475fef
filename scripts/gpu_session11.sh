#!/bin/bash
# Session 11: ep-in-global spill (MID=2) vs MID=1 for chr8..chr12; parity under both.
mkdir -p gpurun_out
for mid in 1 2; do
  MODLE_B200_MID=$mid timeout 600 python scripts/gpu_chrom.py chr8,chr10,chr12,chr13,chrX 512 2 >> gpurun_out/s11_mid.log 2>&1
  MODLE_B200_MID=$mid timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mid_ or c1_chr20 or c4_high or defaults" > gpurun_out/s11_pytest_mid$mid.log 2>&1; echo "rc=$?" >> gpurun_out/s11_pytest_mid$mid.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_count|k_scan|k_fill" -c 12 --csv --log-file gpurun_out/s11_pixels_launches.csv python scripts/bench_pixels.py --reps 0 --no-cpu --cases chr1_loop,c5_loop > gpurun_out/s11_ncu_pixels.log 2>&1
cat gpurun_out/s11_mid.log; tail -2 gpurun_out/s11_pytest_mid1.log gpurun_out/s11_pytest_mid2.log; grep -v "^==" gpurun_out/s11_pixels_launches.csv | cut -d, -f1,5,12- | tail -14
