#!/bin/bash
# Krylov Nystrom root: operator tests, the probe at the benchmark's landmark block, then the bench lines
mkdir -p gpurun_out
T=${1:-r2r}
timeout 900 python -m pytest tests/test_gpu_dense64.py -x -q > gpurun_out/${T}_dense64_tests.log 2>&1; echo "dense64 tests rc=$?"; tail -15 gpurun_out/${T}_dense64_tests.log
timeout 900 python tools/root_probe.py 10 > gpurun_out/${T}_root_probe.log 2>&1; echo "probe rc=$?"; tail -5 gpurun_out/${T}_root_probe.log | cut -c1-1500
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sweep --breakdown gpurun_out/${T}_bd_f10.json > gpurun_out/${T}_bench_f10.json 2> gpurun_out/${T}_bench_f10.err; echo "bench f10 rc=$?"; cut -c1-1200 gpurun_out/${T}_bench_f10.json; tail -5 gpurun_out/${T}_bench_f10.err
