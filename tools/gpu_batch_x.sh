#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2y}
timeout 900 python -m pytest tests/test_gpu_dense64.py tests/test_gpu_ops.py tests/test_gpu_parity.py -x -q -k "dgemm or potrf or qr_rotate or krylov or band_fit or helpers" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
timeout 600 python tools/potrf_probe.py > gpurun_out/${T}_potrf_probe.log 2>&1; echo "potrf probe rc=$?"; cat gpurun_out/${T}_potrf_probe.log
timeout 900 python tools/root_probe.py 10 > gpurun_out/${T}_root_probe.log 2>&1; echo "probe rc=$?"; tail -1 gpurun_out/${T}_root_probe.log | cut -c1-700
