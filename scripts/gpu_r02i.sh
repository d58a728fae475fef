#!/bin/bash
TAG=${1:-r02i}
mkdir -p gpurun_out
for variant in "--streams 3" "--streams 6" "--streams 3 --plan slices" "--streams 12 --plan slices"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 2 --warmup 2 --no-extras $variant 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$variant', 'ms', round(d['ms_per_step']), 'rank_ms', d['config']['rank_ms_per_step'], 'rank_lu', [round(x/1e9,2) for x in d['config']['rank_lef_updates_per_step']], 'e2e ms', round(d['e2e']['ms_per_step']), 'avg launch', round(d['roofline']['avg_launch_ms']))"
done | tee gpurun_out/${TAG}_n2_variants.txt
