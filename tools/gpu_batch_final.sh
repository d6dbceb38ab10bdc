#!/bin/bash
# final records of the round: full GPU suite, smoke, both bench lines, the reference arm, launch lists of one timed step
mkdir -p gpurun_out
T=${1:-r2M}
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/${T}_gpu_tests.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log | cut -c1-300
timeout 1200 python bench.py --steps 3 --warmup 3 --breakdown gpurun_out/${T}_bd_f10.json > gpurun_out/${T}_bench_f10.json 2> gpurun_out/${T}_bench_f10.err; echo "bench f10 rc=$?"; cut -c1-200 gpurun_out/${T}_bench_f10.json
timeout 900 python bench.py --fraction 50 --steps 5 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/${T}_bd_f50.json > gpurun_out/${T}_bench_f50.json 2> gpurun_out/${T}_bench_f50.err; echo "bench f50 rc=$?"; cut -c1-200 gpurun_out/${T}_bench_f50.json
for F in 10 50; do
  B="python bench.py --fraction $F --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-gpu-reference --no-sweep"
  timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/${T}_launches_f${F}.csv $B > /dev/null 2>&1; echo "launch list f$F rc=$?"
  python tools/summarize_launches.py gpurun_out/${T}_launches_f${F}.csv "ncu --nvtx --nvtx-include timed/ --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv: $B (one timed step; per-launch times are cold-cache and serialised: compare SHARES)" > gpurun_out/${T}_launches_f${F}_summary.txt 2>&1
  gzip -f gpurun_out/${T}_launches_f${F}.csv
  head -12 gpurun_out/${T}_launches_f${F}_summary.txt | cut -c1-130
done
