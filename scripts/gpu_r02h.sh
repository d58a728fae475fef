#!/bin/bash
# 2-GPU call: NCCL parity tests, the C++ multi-GPU consumer, bench at N=2 (C2 + extra.c3 split 2 x 4096).
TAG=${1:-r02h}
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/${TAG}_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; echo "bench n2 rc=$?"; tail -2 gpurun_out/${TAG}_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02h_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
x=d.get('extra',{})
print('thr', x.get('throughput_mode',{}).get('value'))
print('c3', json.dumps(x.get('c3'))[:1200])
print(d['config']['parallelism'])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 2 --warmup 2 --plan slices --streams 8 --no-extras > gpurun_out/${TAG}_bench_n2_slices.json 2> gpurun_out/${TAG}_bench_n2_slices.err; echo "bench n2 slices rc=$?"; cut -c1-260 gpurun_out/${TAG}_bench_n2_slices.json
for t in 1 4 8; do
  MODLE_B200_HOST_ADD_THREADS=$t timeout 300 python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('host add threads $t: value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
done
