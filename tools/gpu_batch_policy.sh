#!/bin/bash
# last GPU call of round 2: SpMM cache-flag probe, the GPU suite with every flag on, a bench line with the best flags,
# and `ncu --set full` captures of the two top kernels of the fraction-10 step (traffic for the bench's roofline record)
mkdir -p gpurun_out
T=${1:-r2N}
timeout 170 python tools/spmm_policy_probe.py gpurun_out/${T}_spmm_policy.json > gpurun_out/${T}_spmm_policy.log 2>&1; echo "probe rc=$?"; tail -4 gpurun_out/${T}_spmm_policy.log | cut -c1-400
BEST=$(cat gpurun_out/${T}_spmm_policy.json.best 2>/dev/null || echo 0)
echo "best flags: $BEST"
timeout 90 python -m pytest tests/test_gpu_ops.py -m gpu -q -x > gpurun_out/${T}_gpu_ops_default.log 2>&1; echo "operator tests (default flags) rc=$?"; tail -1 gpurun_out/${T}_gpu_ops_default.log
GNNGP_SPMM_FLAGS=7 timeout 170 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_gpu_tests_flags7.log 2>&1; echo "gpu tests (flags 7) rc=$?"; tail -2 gpurun_out/${T}_gpu_tests_flags7.log
GNNGP_SPMM_FLAGS=$BEST timeout 220 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-sweep --breakdown gpurun_out/${T}_bd_f10.json > gpurun_out/${T}_bench_f10.json 2> gpurun_out/${T}_bench_f10.err; echo "bench f10 (flags $BEST) rc=$?"; cut -c1-200 gpurun_out/${T}_bench_f10.json
GNNGP_SPMM_FLAGS=$BEST timeout 160 ncu --set full --import-source on --clock-control none -k regex:spmm_panel_kernel -c 1 -f -o gpurun_out/${T}_ncu_spmm_k15404 python tools/prof_spmm_big.py 15404 > gpurun_out/${T}_ncu_spmm.log 2>&1; echo "ncu spmm rc=$?"
timeout 110 ncu --set full --import-source on --clock-control none -k regex:gemm_nt_tc3_kernel -c 1 -f -o gpurun_out/${T}_ncu_tc3_backproject python tools/prof_tc_big.py > gpurun_out/${T}_ncu_tc3.log 2>&1; echo "ncu tc3 rc=$?"
ls -la gpurun_out/${T}_*
