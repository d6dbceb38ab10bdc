#!/bin/bash
# N = 2 sanity of the rank-parallel Krylov products: sharded-vs-single parity (small cases + arxiv), then the fraction-10 bench
mkdir -p gpurun_out
T=${1:-r2L}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
GNNGP_CHECK_TAG=_${T} timeout 600 $RUN tools/multigpu_check.py > gpurun_out/${T}_multigpu_check_n2.log 2>&1; echo "check rc=$?"
grep -E "^(small|pubmed|chameleon|arxiv|reddit)" gpurun_out/${T}_multigpu_check_n2.log | cut -c1-200
timeout 600 $RUN bench.py --gpus 2 --steps 3 --warmup 3 --no-sweep > gpurun_out/${T}_bench_f10_n2.json 2> gpurun_out/${T}_bench_f10_n2.err; echo "bench f10 rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${T}_bench_f10_n2.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["config"]["fit_solver"], d["config"]["root_solver"], d["config"]["best"], d["stages_ms_per_step"])
PY
tail -3 gpurun_out/${T}_bench_f10_n2.err
