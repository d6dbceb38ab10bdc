#!/bin/bash
# GPU-box batch: full GPU test suite, smoke, the default bench (fraction 10), the fraction-50 bench, the CPU reference arm.
mkdir -p gpurun_out
T=${1:-r2a}
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_gpu_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 --breakdown gpurun_out/${T}_bd_f10.json > gpurun_out/${T}_bench_f10.json 2> gpurun_out/${T}_bench_f10.err; echo "bench f10 rc=$?"
timeout 600 python bench.py --fraction 50 --steps 5 --warmup 3 --breakdown gpurun_out/${T}_bd_f50.json > gpurun_out/${T}_bench_f50.json 2> gpurun_out/${T}_bench_f50.err; echo "bench f50 rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/${T}_ref_arm_f10.json 2> gpurun_out/${T}_ref_arm_f10.err; echo "ref arm rc=$?"
tail -3 gpurun_out/${T}_gpu_tests.log; tail -2 gpurun_out/${T}_smoke.log; nproc; free -g | sed -n 2p
