"""GPU-box probe: the batched substitution kernel against per-factor torch.cholesky_solve at fit sizes."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnngp_b200.ops import default_ops  # noqa: E402

ops = default_ops()
dev = torch.device("cuda:0")
out = {}


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for m, batch in ((3025, 13),):
    g = torch.Generator(device=dev).manual_seed(0)
    a = torch.randn(m, m + 64, device=dev, dtype=torch.float64, generator=g)
    spd = a @ a.T / m
    del a
    factor = torch.linalg.cholesky(spd + 0.1 * torch.eye(m, device=dev, dtype=torch.float64))
    stack = factor.unsqueeze(0).repeat(batch, 1, 1).contiguous()
    rhs = torch.randn(m, 41, device=dev, dtype=torch.float64, generator=g)
    want = torch.cholesky_solve(rhs, factor)
    got = ops.potrs_batched(stack, rhs)
    row = {"rel_err": float((got[batch - 1] - want).norm() / want.norm()),
           "kernel_ms_all_factors": timeit(lambda: ops.potrs_batched(stack, rhs)),
           "torch_cholesky_solve_ms_per_factor": timeit(lambda: torch.cholesky_solve(rhs, factor)),
           "potrf_ms_per_factor": timeit(lambda: torch.linalg.cholesky_ex(spd))}
    row["library_ms_all_factors"] = row["torch_cholesky_solve_ms_per_factor"] * batch
    out["m=%d batch=%d" % (m, batch)] = row
    print(m, batch, json.dumps(row), flush=True)
    del spd, factor, stack, rhs, want, got
    torch.cuda.empty_cache()
# the whole per-rank solve (potrf on side streams + batched substitution) against the number of streams
for m, count in ((3025, 13), (3025, 26), (6138, 13), (15356, 13)):
    g = torch.Generator(device=dev).manual_seed(1)
    q = torch.randn(m + 512, m, device=dev, generator=g)
    gram = torch.tril(q.T @ q)
    del q
    rhs_t = torch.randn(41, m, device=dev, generator=g)
    shifts = torch.logspace(-1, 2, count, device=dev, dtype=torch.float64)
    row = {}
    for lanes in (1, 2, 4, 8, 13):
        os.environ["GNNGP_SOLVER_STREAMS"] = str(lanes)
        row["streams=%d_ms" % lanes] = timeit(lambda: ops.shifted_solves(gram, rhs_t, shifts), reps=2)
    os.environ["GNNGP_SOLVER_STREAMS"] = "1"
    a, _ = ops.shifted_solves(gram, rhs_t, shifts)
    os.environ["GNNGP_SOLVER_STREAMS"] = "8"
    b, bad = ops.shifted_solves(gram, rhs_t, shifts)
    row["streams8_vs_1_rel"] = float((a - b).norm() / a.norm())
    row["bad"] = int(bad)
    out["shifted_solves m=%d count=%d" % (m, count)] = row
    print("shifted_solves", m, count, json.dumps(row), flush=True)
    del gram, rhs_t, a, b
    torch.cuda.empty_cache()
with open(os.path.join(ROOT, "gpurun_out", "potrs_probe.json"), "w") as fh:
    json.dump(out, fh, indent=1)
