#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2z}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:potrf_inv_block -s 3 -c 1 -f -o gpurun_out/${T}_potrf_block_full python tools/potrf_probe.py > gpurun_out/${T}_ncu_potrf.log 2>&1; echo "ncu potrf rc=$?"
ls -la gpurun_out/${T}_potrf_block_full.ncu-rep
