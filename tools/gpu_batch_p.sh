#!/bin/bash
# profiling batch: launch list of one timed step (fraction 10 and 50), then ncu --set full of the dominant kernels
mkdir -p gpurun_out
T=${1:-r2p}
for F in 50 10; do
  timeout 900 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/${T}_launches_f${F}.csv \
      python bench.py --fraction $F --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-gpu-reference --no-sweep > gpurun_out/${T}_ncu_bench_f${F}.log 2>&1; echo "launch list f$F rc=$?"
  python tools/summarize_launches.py gpurun_out/${T}_launches_f${F}.csv "ncu --nvtx --nvtx-include timed/ --metrics gpu__time_duration.sum --clock-control none: python bench.py --fraction $F --steps 1 --warmup 1 (one timed step; per-launch times are cold-cache and serialised: compare SHARES)" > gpurun_out/${T}_launches_f${F}_summary.txt 2>&1
  head -14 gpurun_out/${T}_launches_f${F}_summary.txt | cut -c1-150
done
# full captures at fraction 50 (each kernel once, inside the timed step)
timeout 1200 ncu --nvtx --nvtx-include "timed/" --set full --import-source on --clock-control none \
    -k regex:"spmm_panel_kernel|gemm_nt_tc3_kernel|dgemm_kernel|sbr_panel|band_cholesky" -c 40 -f -o gpurun_out/${T}_full_f50 \
    python bench.py --fraction 50 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-gpu-reference --no-sweep > gpurun_out/${T}_ncu_full_f50.log 2>&1; echo "ncu full f50 rc=$?"
# the layer-2 sparse product at fraction 10 (k = 15 404): DRAM traffic for the roofline record
timeout 1200 ncu --nvtx --nvtx-include "timed/" --set full --clock-control none -k regex:"spmm_panel_kernel" -c 2 -f -o gpurun_out/${T}_full_spmm_f10 \
    python bench.py --fraction 10 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-gpu-reference --no-sweep > gpurun_out/${T}_ncu_full_spmm_f10.log 2>&1; echo "ncu full spmm f10 rc=$?"
ls -la gpurun_out/${T}_full_*.ncu-rep
