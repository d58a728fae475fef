#!/bin/bash
# What the driver runs at round end on one GPU, for the record under profiles/:
#   gpurun --timeout 1500 -- 'bash scripts/gpu_driver_like.sh <tag>'
TAG=${1:-driver}
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err; echo "reference arm rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_reference_arm.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench_default.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_default.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/${TAG}_bench_reference_arm.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('reference arm', r['value'], r['cpu_baseline']['cores'], 'ratio e2e', d['e2e']['value']/r['value'], 'ratio value', d['value']/r['value'])
print('cpu_baseline in line', d.get('cpu_baseline',{}).get('value'))
x=d.get('extra',{})
print('thr', x.get('throughput_mode',{}).get('value'))
print('c3', {k:x.get('c3',{}).get(k) for k in ('value','ms_per_step','checks')})
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','traffic','avg_launch_ms')})
print('clocks', d['clocks'])
PY
