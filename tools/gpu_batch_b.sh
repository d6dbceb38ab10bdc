#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2b}
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_cli.py -q > gpurun_out/${T}_tc_cli_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_tc_cli_tests.log
timeout 600 python tools/golden_fp64_on_gpu.py arxiv-GCN-nys10-L2 > gpurun_out/${T}_arxiv_fp64.log 2>&1; echo "arxiv fp64 rc=$?"
timeout 600 python tools/spectrum_probe.py > gpurun_out/${T}_spectrum.log 2>&1; echo "spectrum rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/${T}_bd_f10.json > gpurun_out/${T}_bench_f10.json 2> gpurun_out/${T}_bench_f10.err; echo "bench f10 rc=$?"
timeout 300 python tools/dense_probe.py > gpurun_out/${T}_dense_probe.log 2>&1; echo "dense probe rc=$?"
tail -3 gpurun_out/${T}_tc_cli_tests.log; tail -2 gpurun_out/${T}_arxiv_fp64.log; tail -3 gpurun_out/${T}_spectrum.log | cut -c1-600
