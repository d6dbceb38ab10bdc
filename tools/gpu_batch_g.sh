#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2h}
timeout 600 python tools/golden_fp64_on_gpu.py arxiv-GCN-nys10-L2 > gpurun_out/${T}_arxiv_fp64.log 2>&1; echo "arxiv fp64 rc=$?"
timeout 600 python tools/golden_fp64_on_gpu.py reddit-GCN-nys50-L2 > gpurun_out/${T}_reddit_fp64.log 2>&1; echo "reddit fp64 rc=$?"
timeout 900 python -m pytest tests/test_gpu_cli.py tests/test_gpu_fullsize.py -q > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${T}_tests.log | cut -c1-300
