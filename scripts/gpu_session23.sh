#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s23_plain.json 2> gpurun_out/s23_plain.err
echo "plain rc=$?"; grep -h "Error\|error -" gpurun_out/s23_plain.err | sort | uniq -c | head -8 | cut -c1-300; cut -c1-200 gpurun_out/s23_plain.json | tail -n 1
