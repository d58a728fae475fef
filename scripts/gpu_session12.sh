#!/bin/bash
# Session 12: paced cooperative replay of the binned register path; look-back row offsets; all GPU tests.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s12_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s12_pytest.log
for ts in 22 23 24; do
  MODLE_B200_TILE_SHIFT=$ts timeout 600 python scripts/bench_register.py --out gpurun_out/s12_register_ts$ts.json > gpurun_out/s12_register_ts$ts.log 2>&1
done
timeout 900 python scripts/bench_pixels.py --out gpurun_out/s12_pixels.json > gpurun_out/s12_pixels.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum --clock-control none -k regex:"k_bin|k_scatter|k_register" -c 16 --csv --log-file gpurun_out/s12_register_launches.csv python scripts/bench_register.py --reps 0 > gpurun_out/s12_register_ncu.log 2>&1
tail -n 3 gpurun_out/s12_pytest.log
for ts in 22 23 24; do echo ts=$ts; python -c "
import json
for r in json.load(open('gpurun_out/s12_register_ts$ts.json')): print(r['case'], r['stream'], '%.2f ms'%r['ms'], 'frac %.3f'%r['frac_of_hbm_sector_ceiling'])
"; tail -n 2 gpurun_out/s12_register_ts$ts.log | cut -c1-200; done
cut -c1-420 gpurun_out/s12_pixels.log | tail -n 4
