#!/bin/bash
# Multi-GPU session (N = $1): NCCL parity test + bench for C2 (whole chromosomes per rank) and C3
# (one chromosome, cells split over ranks + reduce).
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi_$N.log 2>&1; tail -3 gpurun_out/pytest_multi_$N.log
for wl in "c2 512" "c3 2048"; do
  set -- $wl
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $1 --cells $2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_$1_n$N.json 2> gpurun_out/bench_$1_n$N.err
  tail -2 gpurun_out/bench_$1_n$N.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_$1_n$N.json') if l.startswith('{')][-1]); print('$1 n=$N', 'value %.1f M/s'%(d['value']/1e6), 'ms %.1f'%d['ms_per_step'], 'e2e %.1f M/s'%(d['e2e']['value']/1e6), 'e2e ms %.1f'%d['e2e']['ms_per_step'], d['config']['parallelism'])
except Exception as e: print('FAILED', e)
PY
done
