#!/bin/bash
# band-fit validation, then compact ncu evidence: launch list of one fraction-50 step and full captures of the dominant
# kernels exported to TEXT on the box (the .ncu-rep files are deleted there: gpurun_out must stay under 64 MiB)
mkdir -p gpurun_out
T=${1:-r2q}
timeout 600 python -m pytest tests/test_gpu_ops.py -q -k "band_fit" > gpurun_out/${T}_bandfit_tests.log 2>&1; echo "bandfit tests rc=$?"; tail -2 gpurun_out/${T}_bandfit_tests.log
timeout 600 python tools/bandfit_probe.py > gpurun_out/${T}_bandfit_probe.log 2>&1; echo "probe rc=$?"; tail -3 gpurun_out/${T}_bandfit_probe.log | cut -c1-120
BENCH50="python bench.py --fraction 50 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-gpu-reference --no-sweep"
timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/${T}_launches_f50.csv $BENCH50 > /dev/null 2>&1; echo "launch list f50 rc=$?"
python tools/summarize_launches.py gpurun_out/${T}_launches_f50.csv "ncu --nvtx --nvtx-include timed/ --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv: $BENCH50 (one timed step; per-launch times are cold-cache and serialised: compare SHARES)" > gpurun_out/${T}_launches_f50_summary.txt 2>&1
gzip -f gpurun_out/${T}_launches_f50.csv
METRICS="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread"
for K in spmm_panel_kernel gemm_nt_tc3_kernel dgemm_kernel sbr_panel_cluster band_cholesky; do
  C=3; [ $K = dgemm_kernel ] && C=12
  timeout 420 ncu --nvtx --nvtx-include "timed/" --metrics $METRICS --clock-control none -k regex:$K -c $C --csv --log-file gpurun_out/${T}_ncu_f50_${K}.csv $BENCH50 > /dev/null 2>&1; echo "ncu $K rc=$?"
done
# the layer-2 sparse product at fraction 10 (k = 15 404), second spmm launch of the step
timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics $METRICS --clock-control none -k regex:spmm_panel_kernel -c 2 --csv --log-file gpurun_out/${T}_ncu_f10_spmm_panel_kernel.csv \
    python bench.py --fraction 10 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-gpu-reference --no-sweep > /dev/null 2>&1; echo "ncu spmm f10 rc=$?"
rm -f gpurun_out/*.ncu-rep; du -sh gpurun_out
