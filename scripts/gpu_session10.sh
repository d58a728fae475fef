#!/bin/bash
# Session 10: mid kernel configuration (<512,2>) vs the previous split; register tile size sweep; pixels bench.
mkdir -p gpurun_out
for mid in 0 1; do
  MODLE_B200_MID=$mid timeout 600 python scripts/gpu_chrom.py chr8,chr10,chr13,chr15,chr16 512 2 >> gpurun_out/s10_mid.log 2>&1
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/s10_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s10_pytest.log
for ts in 21 22 23; do
  MODLE_B200_TILE_SHIFT=$ts timeout 600 python scripts/bench_register.py --out gpurun_out/s10_register_ts$ts.json > gpurun_out/s10_register_ts$ts.log 2>&1
done
timeout 900 python scripts/bench_pixels.py --out gpurun_out/s10_pixels.json > gpurun_out/s10_pixels.log 2>&1
cat gpurun_out/s10_mid.log; tail -2 gpurun_out/s10_pytest.log
for ts in 21 22 23; do echo ts=$ts; python -c "
import json
for r in json.load(open('gpurun_out/s10_register_ts$ts.json')): print(r['case'], r['stream'], '%.2f ms'%r['ms'], 'frac %.3f'%r['frac_of_hbm_sector_ceiling'])
"; done
cut -c1-420 gpurun_out/s10_pixels.log | tail -4
