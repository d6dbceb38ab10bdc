"""GPU-box probe: 13 / 26 shifted Cholesky factorisations (fp64, m = 3 025 / 6 138): loop of potrf calls vs one
batched torch call vs a blocked batched algorithm built from batched library calls."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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


def blocked(stack, nb):
    a = stack.clone()
    m = a.shape[1]
    bad = torch.zeros((), dtype=torch.int64, device=a.device)
    for k in range(0, m, nb):
        e = min(m, k + nb)
        lkk, info = torch.linalg.cholesky_ex(a[:, k:e, k:e])
        bad += (info != 0).sum()
        a[:, k:e, k:e] = lkk
        if e < m:
            x = torch.linalg.solve_triangular(lkk.mT, a[:, e:, k:e], upper=True, left=False)
            a[:, e:, k:e] = x
            a[:, e:, e:].baddbmm_(x, x.mT, alpha=-1.0)
    return torch.tril(a), bad


for m, count in ((3025, 13), (3025, 26), (6138, 13)):
    g = torch.Generator(device=dev).manual_seed(1)
    q = torch.randn(m + 512, m, device=dev, dtype=torch.float64, generator=g)
    sym = q.T @ q / m
    del q
    shifts = torch.logspace(-2, 1, count, device=dev, dtype=torch.float64)
    stack = sym.unsqueeze(0).repeat(count, 1, 1)
    stack.diagonal(dim1=1, dim2=2).add_(shifts.view(-1, 1))
    row = {"loop_ms": timeit(lambda: [torch.linalg.cholesky_ex(stack[i]) for i in range(count)]),
           "batched_call_ms": timeit(lambda: torch.linalg.cholesky_ex(stack))}
    want = torch.linalg.cholesky_ex(stack[count - 1])[0]
    for nb in (128, 256, 512):
        row["blocked_nb%d_ms" % nb] = timeit(lambda: blocked(stack, nb))
        got, bad = blocked(stack, nb)
        row["blocked_nb%d_rel" % nb] = float((got[count - 1] - want).norm() / want.norm())
    out["m=%d count=%d" % (m, count)] = row
    print(m, count, json.dumps(row), flush=True)
    del stack, sym
    torch.cuda.empty_cache()
with open(os.path.join(ROOT, "gpurun_out", "potrf_batch_probe.json"), "w") as fh:
    json.dump(out, fh, indent=1)
