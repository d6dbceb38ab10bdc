#!/bin/bash
# launch list of ONE Krylov root on the benchmark's landmark block + the same for one band fit at m = 15405; ncu full of the fast DGEMM
mkdir -p gpurun_out
T=${1:-r2w}
timeout 900 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/${T}_launches_root.csv python tools/root_probe.py 10 --once > gpurun_out/${T}_ncu_root.log 2>&1; echo "launch list root rc=$?"
python tools/summarize_launches.py gpurun_out/${T}_launches_root.csv "ncu --nvtx --nvtx-include timed/ --metrics gpu__time_duration.sum --clock-control none: python tools/root_probe.py 10 --once (ONE Krylov Nystrom root of the benchmark's 15404^2 landmark block; per-launch times are cold-cache and serialised: compare SHARES)" > gpurun_out/${T}_launches_root_summary.txt 2>&1
head -30 gpurun_out/${T}_launches_root_summary.txt | cut -c1-140
gzip -f gpurun_out/${T}_launches_root.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/${T}_launches_bandfit.csv python tools/bandfit_probe.py 20000 15405 > gpurun_out/${T}_ncu_bandfit.log 2>&1; echo "launch list bandfit rc=$?"
python tools/summarize_launches.py gpurun_out/${T}_launches_bandfit.csv "ncu --metrics gpu__time_duration.sum --clock-control none: python tools/bandfit_probe.py 20000 15405 (THREE band fits + two eigen routes at m = 15405; compare SHARES)" > gpurun_out/${T}_launches_bandfit_summary.txt 2>&1
head -16 gpurun_out/${T}_launches_bandfit_summary.txt | cut -c1-140
gzip -f gpurun_out/${T}_launches_bandfit.csv
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dgemm_fast_kernel -c 4 -f -o gpurun_out/${T}_dgemm_fast_full python tools/dgemm_probe.py 15404 1 > gpurun_out/${T}_ncu_dgemm.log 2>&1; echo "ncu dgemm rc=$?"
ls -la gpurun_out/${T}_dgemm_fast_full.ncu-rep
