#!/bin/bash
mkdir -p gpurun_out
for w in 32768 16384; do
  MODLE_B200_LARGE_WINDOW=$w timeout 600 python scripts/gpu_phases.py c3 148 2 2>&1 | grep -i "product\|rng_refill\|fault" | sed "s/^/[W=$w] /" >> gpurun_out/s17_window.log
  MODLE_B200_LARGE_WINDOW=$w timeout 600 python scripts/gpu_chrom.py chr1,chr5,chr8 512 2 2>&1 | sed "s/^/[W=$w] /" >> gpurun_out/s17_window.log
  MODLE_B200_LARGE_WINDOW=$w timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_simulate_cells -c 1 --csv --log-file gpurun_out/s17_ncu_w$w.csv python scripts/gpu_phases.py c3 148 1 > /dev/null 2>&1
  grep -v "^==" gpurun_out/s17_ncu_w$w.csv | cut -d, -f12- | tail -n 3 | sed "s/^/[W=$w] /" >> gpurun_out/s17_window.log
done
cat gpurun_out/s17_window.log | cut -c1-220
MODLE_B200_LARGE_WINDOW=16384 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c3_chr1 or large_chr8 or c4_high" > gpurun_out/s17_pytest_w16k.log 2>&1; tail -n 2 gpurun_out/s17_pytest_w16k.log
timeout 900 python bench.py > gpurun_out/s17_bench_default.json 2> gpurun_out/s17_bench_default.err
MODLE_B200_LARGE_WINDOW=16384 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/s17_bench_w16k.json 2> gpurun_out/s17_bench_w16k.err
python -c "
import json
for f in ('s17_bench_default','s17_bench_w16k'):
    d=json.loads([l for l in open('gpurun_out/'+f+'.json') if l.startswith('{')][-1]); print(f, 'value %.1f M/s'%(d['value']/1e6), 'ms %.1f'%d['ms_per_step'], 'e2e %.1f M/s'%(d['e2e']['value']/1e6), d.get('cpu_baseline',{}).get('value'))
"
