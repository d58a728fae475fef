#!/bin/bash
TAG=${1:-r02l}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
for st in 3; do
timeout 600 python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --streams $st 2>/dev/null > gpurun_out/${TAG}_bench_streams$st.json
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_streams$st.json').read().strip().splitlines()[-1])
print('streams $st', round(d['ms_per_step']), 'e2e', round(d['e2e']['ms_per_step']))
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 2 --warmup 1 --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=2 whole', 'ms', round(d['ms_per_step']), 'rank_ms', d['config']['rank_ms_per_step'], 'e2e ms', round(d['e2e']['ms_per_step']))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 --steps 2 --warmup 1 --no-extras --pre-split 2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=2 pre-split 2', 'ms', round(d['ms_per_step']), 'rank_ms', d['config']['rank_ms_per_step'], 'e2e ms', round(d['e2e']['ms_per_step']), d['config']['parallelism'][:90])"
