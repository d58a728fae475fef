#!/bin/bash
# 8-GPU call: NCCL parity tests at 2/4/8, C++ consumer at 8, bench at N=8 (default plan with extras), plan/stream variants.
TAG=${1:-r02j}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/${TAG}_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_n8.json 2> gpurun_out/${TAG}_bench_n8.err; echo "bench n8 rc=$?"; tail -2 gpurun_out/${TAG}_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02j_bench_n8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['rank_ms_per_step'])
x=d.get('extra',{})
print('thr', x.get('throughput_mode',{}))
print('c3', json.dumps(x.get('c3'))[:1500])
PY
for variant in "--streams 6" "--streams 3 --plan slices" "--streams 8 --plan slices" "--streams 24 --plan slices"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 8 --steps 3 --warmup 3 --no-extras $variant 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$variant', 'ms', round(d['ms_per_step']), 'G/s', round(d['value']/1e9,2), 'rank_ms', d['config']['rank_ms_per_step'], 'e2e ms', round(d['e2e']['ms_per_step']))"
done | tee gpurun_out/${TAG}_n8_variants.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 8 --steps 3 --warmup 3 --no-extras --plan slices --streams 8 --rng-mode throughput 2>/dev/null | cut -c1-300
