#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2x}
timeout 900 python -m pytest tests/test_gpu_dense64.py tests/test_gpu_ops.py tests/test_gpu_parity.py -x -q -k "dgemm or potrf or qr_rotate or krylov or band_fit or helpers" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${T}_tests.log
timeout 900 python tools/root_probe.py 10 > gpurun_out/${T}_root_probe.log 2>&1; echo "probe rc=$?"; tail -1 gpurun_out/${T}_root_probe.log | cut -c1-1300
GNNGP_BAND_LOOKAHEAD=0 timeout 600 python tools/bandfit_probe.py > gpurun_out/${T}_bandfit_probe_serial.log 2>&1; echo "bandfit probe (serial) rc=$?"; tail -3 gpurun_out/${T}_bandfit_probe_serial.log | cut -c1-120
timeout 600 python tools/bandfit_probe.py > gpurun_out/${T}_bandfit_probe.log 2>&1; echo "bandfit probe (look-ahead) rc=$?"; tail -3 gpurun_out/${T}_bandfit_probe.log | cut -c1-120
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sweep --no-gpu-reference --breakdown gpurun_out/${T}_bd_f10.json > gpurun_out/${T}_bench_f10.json 2> gpurun_out/${T}_bench_f10.err; echo "bench f10 rc=$?"; cut -c1-300 gpurun_out/${T}_bench_f10.json; tail -3 gpurun_out/${T}_bench_f10.err
