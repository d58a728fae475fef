#!/bin/bash
# Quick GPU check: parity tests + short benches of a small (chr20) and a large (chr1) interval.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --workload c1 --steps 2 --warmup 1 --no-cpu-baseline 2> gpurun_out/q_c1.err | tee gpurun_out/q_c1.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c1', d['value']/1e6, 'M LU/s', d['ms_per_step'], 'ms', 'e2e', d['e2e']['value']/1e6)"
python bench.py --workload c3 --cells 296 --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/q_c3.err | tee gpurun_out/q_c3.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3/296', d['value']/1e6, 'M LU/s', d['ms_per_step'], 'ms', 'e2e', d['e2e']['value']/1e6)"
