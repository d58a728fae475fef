#!/bin/bash
# Multi-GPU check (N = 2, 4 or 8 GPUs of one box; charged N x):
#   gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_multi.sh N <tag>'
# 1. tests/test_gpu_multi.py: the sharded Python path over NCCL vs the oracle, and the C++ host
#    (tests/cabi/consumer_multi.cpp: planner + device-resident simulate + modle_b200_reduce_band)
# 2. the bench line at N (C2 + extra.c3 = chr1 x 8192 cells split over the ranks, reduce proven
#    against the single-GPU golden and the oracle) and the plan variants
N=${1:-2}; TAG=${2:-multi}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/${TAG}_pytest_multi.log
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29631 --steps 5 --warmup 5 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e ms', d['e2e']['ms_per_step'], d['config']['rank_ms_per_step'])
x=d.get('extra',{})
print('thr', x.get('throughput_mode',{}).get('value'), x.get('throughput_mode',{}).get('ms_per_step'))
print('c3', {k:x.get('c3',{}).get(k) for k in ('value','ms_per_step','reduce_ms','parallelism','checks')})
PY
for variant in "--plan whole" "--plan slices --streams 8"; do
  run 29632 --steps 5 --warmup 5 --no-extras $variant 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$variant', 'ms', round(d['ms_per_step']), 'G/s', round(d['value']/1e9,2), 'rank_ms', d['config']['rank_ms_per_step'], 'e2e ms', round(d['e2e']['ms_per_step']))"
done | tee gpurun_out/${TAG}_n${N}_variants.txt
