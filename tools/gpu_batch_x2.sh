#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2C}
timeout 900 python -m pytest tests/test_gpu_dense64.py tests/test_gpu_parity.py -x -q -k "krylov" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
timeout 900 python tools/root_probe.py 10 > gpurun_out/${T}_root_probe.log 2>&1; echo "probe rc=$?"; tail -1 gpurun_out/${T}_root_probe.log | cut -c1-700
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sweep --no-gpu-reference --breakdown gpurun_out/${T}_bd_f10.json > gpurun_out/${T}_bench_f10.json 2> gpurun_out/${T}_bench_f10.err; echo "bench f10 rc=$?"; cut -c1-200 gpurun_out/${T}_bench_f10.json
