#!/bin/bash
# 8-GPU diagnostic, kept short (charged 8x).
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=3 timeout 60 python scripts/gpu_chrom.py chr1 148 1 > gpurun_out/s24_single_gpu3.log 2>&1 &
tr() { # name, env...
  name=$1; shift
  env "$@" timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 --workload c3 --cells 2048 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s24_$name.json 2> gpurun_out/s24_$name.err
  echo "$name rc=$? $(grep -h 'Error\|error -' gpurun_out/s24_$name.err | sort | uniq -c | head -3 | cut -c1-200) $(cut -c1-160 gpurun_out/s24_$name.json | tail -n 1)"
}
PORT=29551 tr oldcfg MODLE_B200_LARGE_WINDOW=32768 MODLE_B200_MID=0
PORT=29552 tr nonvls NCCL_NVLS_ENABLE=0
PORT=29553 tr default X=1
wait
tail -n 2 gpurun_out/s24_single_gpu3.log | cut -c1-200
