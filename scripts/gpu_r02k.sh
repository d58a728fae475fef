#!/bin/bash
TAG=${1:-r02k}
mkdir -p gpurun_out
for st in 1 3; do
timeout 600 python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --streams $st 2>/dev/null > gpurun_out/${TAG}_bench_streams$st.json
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_streams$st.json').read().strip().splitlines()[-1])
print('streams $st', round(d['ms_per_step']), 'e2e', round(d['e2e']['ms_per_step']))
print(json.dumps(d['config']['rank0_launch_ms']))
PY
done
