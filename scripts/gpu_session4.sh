#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
{
timeout 300 python scripts/gpu_phases.py c1 512 2
timeout 300 python scripts/gpu_phases.py c3 296 2
timeout 300 python scripts/gpu_phases.py c4 296 1
} > gpurun_out/phases4.log 2>&1
cat gpurun_out/phases4.log
timeout 900 python bench.py --steps 1 --warmup 1 --streams 3 --no-cpu-baseline > gpurun_out/bench_c2_s3.json 2> gpurun_out/bench_c2_s3.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c2_s3.json')):
    try:
        d=json.load(open(f)); print(f, 'value %.1f M/s'%(d['value']/1e6), 'ms %.1f'%d['ms_per_step'], 'e2e %.1f M/s'%(d['e2e']['value']/1e6), 'e2e ms %.1f'%d['e2e']['ms_per_step'])
    except Exception as e: print(f, 'FAILED', e)
PY
tail -3 gpurun_out/bench_c2_s3.err
