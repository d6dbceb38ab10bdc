#!/bin/bash
# GPU call 29 of round 2 (5 GPU-minutes left): the short-row SpMM kernel against the plain kernels at the arxiv shape, the
# GPU suite at the library's default flags and with every flag on, the arxiv-shape bench line, the Reddit-shape line.
mkdir -p gpurun_out
T=${1:-r2P}
timeout 100 python tools/spmm_policy_probe.py gpurun_out/${T}_spmm_policy_arxiv.json --shapes arxiv:9200 > gpurun_out/${T}_spmm_policy_arxiv.log 2>&1; echo "probe rc=$?"; tail -3 gpurun_out/${T}_spmm_policy_arxiv.log | cut -c1-300
timeout 100 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_gpu_tests_default.log 2>&1; echo "gpu tests (default flags) rc=$?"; tail -1 gpurun_out/${T}_gpu_tests_default.log
GNNGP_SPMM_FLAGS=7 timeout 100 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_gpu_tests_flags7.log 2>&1; echo "gpu tests (flags 7) rc=$?"; tail -1 gpurun_out/${T}_gpu_tests_flags7.log
GNNGP_SPMM_FLAGS=7 timeout 90 python bench.py --shape arxiv --fraction 10 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-sweep > gpurun_out/${T}_bench_arxiv_f10.json 2> gpurun_out/${T}_bench_arxiv_f10.err; echo "bench arxiv rc=$?"; cut -c1-200 gpurun_out/${T}_bench_arxiv_f10.json
timeout 80 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-sweep --no-e2e > gpurun_out/${T}_bench_reddit_f10.json 2> gpurun_out/${T}_bench_reddit_f10.err; echo "bench reddit rc=$?"; cut -c1-200 gpurun_out/${T}_bench_reddit_f10.json
