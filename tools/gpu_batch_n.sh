#!/bin/bash
# multi-GPU batch under torchrun: sharded-vs-single parity check, then the bench at fraction 10 and 50
mkdir -p gpurun_out
N=${1:-2}
T=${2:-r2n}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
if [ "${3:-check}" = "check" ]; then
  GNNGP_CHECK_TAG=_${T} timeout 900 $RUN tools/multigpu_check.py --reddit > gpurun_out/${T}_multigpu_check_n${N}.log 2>&1; echo "check rc=$?"
  grep -E "^(small|pubmed|chameleon|arxiv|reddit)" gpurun_out/${T}_multigpu_check_n${N}.log | cut -c1-420
fi
timeout 900 $RUN bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${T}_bench_f10_n${N}.json 2> gpurun_out/${T}_bench_f10_n${N}.err; echo "bench f10 rc=$?"
timeout 600 $RUN bench.py --gpus $N --steps 5 --warmup 3 --fraction 50 > gpurun_out/${T}_bench_f50_n${N}.json 2> gpurun_out/${T}_bench_f50_n${N}.err; echo "bench f50 rc=$?"
tail -c 600 gpurun_out/${T}_bench_f10_n${N}.json | head -c 300; echo
python - <<PY
import json
for f in ("gpurun_out/${T}_bench_f10_n${N}.json", "gpurun_out/${T}_bench_f50_n${N}.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["config"]["fit_solver"], {k: v for k, v in list(d["stages_ms_per_step"].items())[:7]})
    except Exception as e:
        print(f, "ERR", e)
PY
