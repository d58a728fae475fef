#!/bin/bash
mkdir -p gpurun_out
for st in 2 3 4 6 8 12; do
  timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --streams $st > gpurun_out/s18_bench_streams$st.json 2> gpurun_out/s18_bench_streams$st.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/s18_bench_streams$st.json') if l.startswith('{')][-1]); print('streams $st', 'value %.1f M/s'%(d['value']/1e6), 'ms %.1f'%d['ms_per_step'], 'e2e %.1f M/s'%(d['e2e']['value']/1e6), 'e2e ms %.1f'%d['e2e']['ms_per_step'])
"
done
timeout 600 python -m pytest tests/test_pixels.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/s18_pytest.log 2>&1; tail -n 2 gpurun_out/s18_pytest.log
