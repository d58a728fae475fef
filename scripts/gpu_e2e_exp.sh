for w in 3 4 6 8; do
  timeout 300 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --e2e-workers $w 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e2e workers $w: value ms', round(d['ms_per_step']), 'e2e ms', round(d['e2e']['ms_per_step']))"
done
for t in 1 8; do
  MODLE_B200_HOST_ADD_THREADS=$t timeout 300 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --e2e-workers 4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('host add threads $t: value ms', round(d['ms_per_step']), 'e2e ms', round(d['e2e']['ms_per_step']))"
done
