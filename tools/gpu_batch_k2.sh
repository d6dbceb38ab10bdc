#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2G}
timeout 900 python bench.py --fraction 50 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference --breakdown gpurun_out/${T}_bd_f50.json > gpurun_out/${T}_bench_f50.json 2> gpurun_out/${T}_bench_f50.err; echo "bench f50 rc=$?"; cut -c1-200 gpurun_out/${T}_bench_f50.json
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_f50.json")); print(d["config"]["fit_solver"], d["config"]["fit_solver_stats"]); print(d["stages_ms_per_step"])
PY
