#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2d}
M=${2:-3053}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/${T}_bandfit_launches_${M}.csv python tools/bandfit_probe.py ${M} ${M} > gpurun_out/${T}_ncu_probe.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/${T}_bandfit_launches_${M}.csv > gpurun_out/${T}_bandfit_launches_${M}_summary.txt 2>&1
head -40 gpurun_out/${T}_bandfit_launches_${M}_summary.txt | cut -c1-200
