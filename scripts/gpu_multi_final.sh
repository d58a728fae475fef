N=$1; TAG=$2
mkdir -p gpurun_out
if [ "$3" = "tests" ]; then timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/${TAG}_pytest_multi.log; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus $N --steps 8 --warmup 5 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e ms', d['e2e']['ms_per_step'], d['config']['rank_ms_per_step'])
x=d.get('extra',{})
print('thr', x.get('throughput_mode',{}).get('value'), x.get('throughput_mode',{}).get('ms_per_step'))
print('c3', {k:x.get('c3',{}).get(k) for k in ('value','ms_per_step','reduce_ms','parallelism','checks')})
print(d['config']['parallelism'])
PY
