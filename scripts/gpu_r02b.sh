#!/bin/bash
# Round 2, second GPU call (1 GPU): new tests, the bench line with its extras (+ the C3 golden),
# phases of C4 / C5, ncu of both RNG modes and of the register path, racecheck of both modes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02b_pytest.log
timeout 600 python bench.py --steps 2 --warmup 3 --write-c3-golden > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02b_bench.err
cp tests/golden/c3_chr1_8192cells_checksums.json gpurun_out/ 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02b_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'))
x=d.get('extra',{})
print('thr', x.get('throughput_mode',{}).get('value'))
print('c3', {k:x.get('c3',{}).get(k) for k in ('value','ms_per_step','reduce_ms','checks')})
print('reg', json.dumps(x.get('register'))[:1500])
PY
for wl in "c4 148" "c5 148"; do
  for mode in 0 1; do
    MODLE_B200_RNG_MODE=$mode timeout 300 python scripts/gpu_phases.py $wl 1 > gpurun_out/r02b_phases_${wl%% *}_mode$mode.txt 2>&1; echo "phases $wl mode $mode rc=$?"; grep product gpurun_out/r02b_phases_${wl%% *}_mode$mode.txt
  done
done
for mode in 0 1; do
  MODLE_B200_BENCH_CHROMS=chr1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_simulate_cells -c 1 \
    -o gpurun_out/r02b_ncu_chr1_mode$mode python bench.py --steps 1 --warmup 0 --cells 148 --rng-mode $([ $mode = 1 ] && echo throughput || echo deterministic) --no-cpu-baseline --no-extras --streams 1 > gpurun_out/r02b_ncu_mode$mode.log 2>&1; echo "ncu mode $mode rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum,l1tex__t_requests_pipe_lsu_mem_global_op_red.sum --clock-control none -k regex:"k_bin|k_scatter|k_register|k_calibrate" --csv --log-file gpurun_out/r02b_register_ncu.csv python scripts/bench_register.py --reps 1 > gpurun_out/r02b_register_ncu.log 2>&1; echo "register ncu rc=$?"
for mode in 0 1; do
  MODLE_B200_RNG_MODE=$mode timeout 500 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --print-limit 20 \
    python scripts/gpu_small_cases.py burnin,sampling,whole_small,whole_c4,whole_pblock,mid,large > gpurun_out/r02b_racecheck_mode$mode.log 2>&1; echo "racecheck mode $mode rc=$?"; grep -c "Race reported" gpurun_out/r02b_racecheck_mode$mode.log; tail -2 gpurun_out/r02b_racecheck_mode$mode.log
done
