#!/bin/bash
# GPU call 30 of round 2 (the last 2.9 GPU-minutes): the short-row SpMM kernel with eight gathers in flight, the GPU suite
# with every flag on, one `ncu --set full` capture of the short-row kernel at the arxiv shape.
mkdir -p gpurun_out
T=${1:-r2Q}
timeout 55 python tools/spmm_policy_probe.py gpurun_out/${T}_spmm_policy_arxiv.json --shapes arxiv:9200 > gpurun_out/${T}_spmm_policy_arxiv.log 2>&1; echo "probe rc=$?"; tail -2 gpurun_out/${T}_spmm_policy_arxiv.log | cut -c1-300
BEST=$(cat gpurun_out/${T}_spmm_policy_arxiv.json.best 2>/dev/null || echo 7)
GNNGP_SPMM_FLAGS=15 timeout 60 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_gpu_tests_flags15.log 2>&1; echo "gpu tests (flags 15) rc=$?"; tail -1 gpurun_out/${T}_gpu_tests_flags15.log
GNNGP_SPMM_FLAGS=$BEST timeout 45 ncu --set full --import-source on --clock-control none -k regex:spmm_short_rows_kernel -c 1 -f -o gpurun_out/${T}_ncu_spmm_short_rows_arxiv python tools/prof_spmm_big.py 9200 arxiv > gpurun_out/${T}_ncu_spmm_short.log 2>&1; echo "ncu short rows (flags $BEST) rc=$?"
