#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2f}
timeout 600 python tools/bandfit_probe.py 6138 > gpurun_out/${T}_bandfit_probe.log 2>&1; echo "probe rc=$?"
tail -3 gpurun_out/${T}_bandfit_probe.log | cut -c1-200
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"band_cholesky_solve|sbr_panel" -s 24 -c 24 -f -o gpurun_out/${T}_bandfit_full python tools/bandfit_probe.py 3053 > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/${T}_bandfit_full.ncu-rep
