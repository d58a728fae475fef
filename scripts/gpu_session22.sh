#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline --cells 32 > gpurun_out/s22_plain.json 2> gpurun_out/s22_plain.err
echo "plain rc=$?"; grep -h "Error" gpurun_out/s22_plain.err | sort | uniq -c | head -5 | cut -c1-250; cut -c1-200 gpurun_out/s22_plain.json | tail -n 1
CUDA_LAUNCH_BLOCKING=1 timeout 900 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 6 python bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline --cells 4 > gpurun_out/s22_memcheck.log 2>&1
grep -v "^W10\|^\[W" gpurun_out/s22_memcheck.log | grep -v "Host Frame: \(_Py\|Py\|\[0x\|python\|cfunction\|method\)" | head -80 | cut -c1-230
