#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2F}
timeout 900 python -m pytest tests/test_gpu_dense64.py tests/test_gpu_parity.py -x -q -k "krylov or band_solve or dgemm" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
timeout 600 python tools/bandfit_probe.py > gpurun_out/${T}_bandfit_probe.log 2>&1; echo "bandfit probe rc=$?"; tail -3 gpurun_out/${T}_bandfit_probe.log | cut -c1-500
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/${T}_bd_f10.json > gpurun_out/${T}_bench_f10.json 2> gpurun_out/${T}_bench_f10.err; echo "bench f10 rc=$?"; cut -c1-200 gpurun_out/${T}_bench_f10.json; tail -2 gpurun_out/${T}_bench_f10.err
timeout 900 python bench.py --fraction 50 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference --breakdown gpurun_out/${T}_bd_f50.json > gpurun_out/${T}_bench_f50.json 2> gpurun_out/${T}_bench_f50.err; echo "bench f50 rc=$?"; cut -c1-200 gpurun_out/${T}_bench_f50.json
