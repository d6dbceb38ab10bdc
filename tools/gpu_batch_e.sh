#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2e}
timeout 900 python -m pytest tests/test_gpu_ops.py -q -k "band_fit" > gpurun_out/${T}_bandfit_tests.log 2>&1; echo "bandfit tests rc=$?"
tail -8 gpurun_out/${T}_bandfit_tests.log | cut -c1-300
timeout 600 python tools/bandfit_probe.py > gpurun_out/${T}_bandfit_probe.log 2>&1; echo "probe rc=$?"
tail -4 gpurun_out/${T}_bandfit_probe.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${T}_bandfit_launches_3053.csv python tools/bandfit_probe.py 3053 > gpurun_out/${T}_ncu_probe.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/${T}_bandfit_launches_3053.csv > gpurun_out/${T}_bandfit_launches_3053_summary.txt 2>&1
head -12 gpurun_out/${T}_bandfit_launches_3053_summary.txt | cut -c1-160
