#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2u}
timeout 900 python -m pytest tests/test_gpu_dense64.py tests/test_gpu_ops.py -x -q -k "dgemm or potrf or qr_rotate or krylov or band_fit or helpers" > gpurun_out/${T}_dense64_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/${T}_dense64_tests.log
timeout 600 python tools/dgemm_probe.py > gpurun_out/${T}_dgemm_probe.log 2>&1; echo "dgemm probe rc=$?"; cat gpurun_out/${T}_dgemm_probe.log | cut -c1-220
timeout 900 python tools/root_probe.py 10 > gpurun_out/${T}_root_probe.log 2>&1; echo "probe rc=$?"; tail -1 gpurun_out/${T}_root_probe.log | cut -c1-1300
timeout 600 python tools/bandfit_probe.py 20000 15405 > gpurun_out/${T}_bandfit_probe.log 2>&1; echo "bandfit probe rc=$?"; tail -1 gpurun_out/${T}_bandfit_probe.log | cut -c1-400
