#!/bin/bash
# round-end style validation: full GPU suite, smoke, both bench lines, the reference arm
mkdir -p gpurun_out
T=${1:-r2A}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log | cut -c1-300
timeout 1200 python bench.py --steps 3 --warmup 3 --breakdown gpurun_out/${T}_bd_f10.json > gpurun_out/${T}_bench_f10.json 2> gpurun_out/${T}_bench_f10.err; echo "bench f10 rc=$?"; cut -c1-200 gpurun_out/${T}_bench_f10.json
timeout 900 python bench.py --fraction 50 --steps 5 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/${T}_bd_f50.json > gpurun_out/${T}_bench_f50.json 2> gpurun_out/${T}_bench_f50.err; echo "bench f50 rc=$?"; cut -c1-200 gpurun_out/${T}_bench_f50.json
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${T}_ref_arm_f10.json 2> gpurun_out/${T}_ref_arm_f10.err; echo "reference arm rc=$?"; cut -c1-300 gpurun_out/${T}_ref_arm_f10.json
