#!/bin/bash
# fp64 GEMM probe + ncu full capture of dgemm_kernel (source-level), then the whole GPU suite
mkdir -p gpurun_out
T=${1:-r2t}
timeout 600 python tools/dgemm_probe.py > gpurun_out/${T}_dgemm_probe.log 2>&1; echo "dgemm probe rc=$?"; cat gpurun_out/${T}_dgemm_probe.log | cut -c1-220
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dgemm_kernel -c 6 -f -o gpurun_out/${T}_dgemm_full python tools/dgemm_probe.py 15404 1 > gpurun_out/${T}_ncu_dgemm.log 2>&1; echo "ncu dgemm rc=$?"
ls -la gpurun_out/${T}_dgemm_full.ncu-rep
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/${T}_gpu_tests.log
