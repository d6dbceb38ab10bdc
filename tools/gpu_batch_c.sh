#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2c}
timeout 900 python -m pytest tests/test_gpu_ops.py -q -k "band_fit" > gpurun_out/${T}_bandfit_tests.log 2>&1; echo "bandfit tests rc=$?"
tail -15 gpurun_out/${T}_bandfit_tests.log
timeout 600 python tools/bandfit_probe.py > gpurun_out/${T}_bandfit_probe.log 2>&1; echo "probe rc=$?"
tail -5 gpurun_out/${T}_bandfit_probe.log | cut -c1-500
