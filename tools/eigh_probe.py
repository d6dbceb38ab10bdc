"""GPU-box probe: what the dense symmetric solvers cost at the landmark-block sizes of the Reddit shape.

Times torch.linalg.eigh (cuSOLVER and, when the build has it, MAGMA) in fp32 / fp64, and the fp64 Cholesky /
LU factorisations a nugget-parallel fit would use instead of the replicated eigendecomposition.
Writes gpurun_out/eigh_probe.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = torch.device("cuda:0")
out = {}


def timeit(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


sizes = [int(s) for s in (sys.argv[1:] or ["3025", "6138", "15344"])]
g = torch.Generator(device=dev).manual_seed(0)
for n in sizes:
    x = torch.randn(n, 700, device=dev, generator=g)
    x = x / x.norm(dim=1, keepdim=True)
    c = (x @ x.T).clamp(-1, 1)
    th = torch.acos(c)
    k32 = ((torch.sin(th) + (torch.pi - th) * c) / (2 * torch.pi)).contiguous()      # an arc-cosine Gram block
    k32 = 0.5 * (k32 + k32.T)
    k64 = k32.double()
    del x, c, th
    row = {}
    for lib in ("cusolver", "magma"):
        try:
            torch.backends.cuda.preferred_linalg_library(lib)
            row["eigh_f64_%s_ms" % lib] = timeit(lambda: torch.linalg.eigh(k64), reps=1 if n > 8000 else 2)
            row["eigh_f32_%s_ms" % lib] = timeit(lambda: torch.linalg.eigh(k32), reps=1 if n > 8000 else 2)
            row["eigvalsh_f64_%s_ms" % lib] = timeit(lambda: torch.linalg.eigvalsh(k64), reps=1 if n > 8000 else 2)
        except Exception as exc:
            row["eigh_%s_error" % lib] = repr(exc)[:200]
    torch.backends.cuda.preferred_linalg_library("default")
    shifted = k64 + 1e-3 * torch.eye(n, device=dev, dtype=torch.float64)
    rhs = torch.randn(n, 41, device=dev, dtype=torch.float64, generator=g)
    row["cholesky_f64_ms"] = timeit(lambda: torch.linalg.cholesky_ex(shifted))
    L = torch.linalg.cholesky_ex(shifted)[0]
    row["cholesky_solve_41rhs_f64_ms"] = timeit(lambda: torch.cholesky_solve(rhs, L))
    row["lu_factor_f64_ms"] = timeit(lambda: torch.linalg.lu_factor_ex(shifted))
    s32 = shifted.float()
    row["cholesky_f32_ms"] = timeit(lambda: torch.linalg.cholesky_ex(s32))
    row["gemm_f64_nxnxn_ms"] = timeit(lambda: k64 @ k64)
    row["gemm_f64_tflops"] = 2.0 * n ** 3 / (row["gemm_f64_nxnxn_ms"] * 1e-3) / 1e12
    # accuracy of the clamped root's projector V h(D) V^T against fp64 eigh, for the fp32 solver
    w64, v64 = torch.linalg.eigh(k64)
    w32, v32 = torch.linalg.eigh(k32)

    def filt(w, v):
        w = w.double(); v = v.double()
        tau = 1e-4 * w.max()
        s2 = w.clamp(min=0) / torch.maximum(w, tau) ** 2
        return (v * s2) @ v.T

    want = filt(w64, v64)
    row["root_filter_rel_err_f32_solver"] = float((filt(w32, v32) - want).norm() / want.norm())
    row["eig_below_tau"] = int((w64 < 1e-4 * w64.max()).sum())
    out[str(n)] = row
    print(n, json.dumps(row), flush=True)
    del k32, k64, shifted, L, s32, w64, v64, w32, v32, want, rhs
    torch.cuda.empty_cache()

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "eigh_probe.json"), "w") as fh:
    json.dump(out, fh, indent=1)
