#!/bin/bash
# Round-end record on one GPU: the `-m gpu` suite, the per-phase cycle tables of both RNG modes and
# the default bench line with its extras (the command the driver runs, 20 steps).
#   gpurun --timeout 900 -- 'bash scripts/gpu_final.sh <tag>'
TAG=${1:-final}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
for mode in 0 1; do
  (MODLE_B200_RNG_MODE=$mode timeout 200 python scripts/gpu_phases.py c3 148; MODLE_B200_RNG_MODE=$mode timeout 200 python scripts/gpu_phases.py c1 444) > gpurun_out/${TAG}_phases_mode$mode.txt 2>&1; echo "phases mode $mode rc=$?"; grep product gpurun_out/${TAG}_phases_mode$mode.txt
done
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_default.json").read().strip().split("\n")[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "cpu", d.get("cpu_baseline", {}).get("value"))
x = d.get("extra", {})
print("thr", x.get("throughput_mode", {}).get("value"), "c3", x.get("c3", {}).get("value"), x.get("c3", {}).get("checks", {}).get("errors"))
PY
