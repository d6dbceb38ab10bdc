#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2I}
timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/${T}_bd_f10.json > gpurun_out/${T}_bench_f10.json 2> gpurun_out/${T}_bench_f10.err; echo "bench f10 rc=$?"; cut -c1-200 gpurun_out/${T}_bench_f10.json; tail -3 gpurun_out/${T}_bench_f10.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_f10.json")); c=d["config"]
print(c["fit_rel_vs_fp64_solve_on_own_factor"]); print(c["parity_vs_reference"])
PY
