"""GPU-box probe: the fp64 GEMM (csrc/dgemm.cuh) at the shapes of the band reduction and the Krylov root, against cuBLAS."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnngp_b200.ops import default_ops  # noqa: E402

dev = torch.device("cuda:0")
ops = default_ops()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 15404
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


gen = torch.Generator(device=dev).manual_seed(1)
m = torch.randn(n, n, device=dev, dtype=torch.float64, generator=gen)
v = torch.randn(n, 128, device=dev, dtype=torch.float64, generator=gen)
w = torch.randn(n, 128, device=dev, dtype=torch.float64, generator=gen)
vw = torch.cat([v, w], 1).contiguous()
wv = torch.cat([w, v], 1).contiguous()
c = torch.empty(n, 128, device=dev, dtype=torch.float64)
big = torch.zeros(n, n, device=dev, dtype=torch.float64)
out = {}
cases = {
    "M @ V  (n x 128 x n, Lanczos / Y = A22 V)": (lambda: ops.dgemm(m, v, out=c), lambda: torch.mm(m, v, out=c), 2.0 * n * n * 128),
    "V^T M  (128 x n x n)": (lambda: ops.dgemm(v.T, m, out=c.T if False else torch.empty(128, n, device=dev, dtype=torch.float64)),
                               lambda: torch.mm(v.T, m), 2.0 * n * n * 128),
    "A -= P P^T lower (n x n x 128, Cholesky update)": (lambda: ops.dgemm(v, v.T, out=big, alpha=-1.0, beta=1.0, lower=True),
                                                          lambda: torch.addmm(big, v, v.T, alpha=-1.0, out=big), 1.0 * n * n * 128),
    "A -= [V W][W V]^T (n x n x 256, band-reduction update)": (lambda: ops.dgemm(vw, wv.T, out=big, alpha=-1.0, beta=1.0),
                                                                 lambda: torch.addmm(big, vw, wv.T, alpha=-1.0, out=big), 2.0 * n * n * 256),
    "X (n x n) @ V (n x 128) transposed operand (L^T V)": (lambda: ops.dgemm(m.T, v, out=c), lambda: torch.mm(m.T, v, out=c), 2.0 * n * n * 128),
}
for name, (own, lib, flops) in cases.items():
    t_own, t_lib = timeit(own), timeit(lib)
    out[name] = {"own_ms": round(t_own, 3), "cublas_ms": round(t_lib, 3), "own_tflops": round(flops / t_own / 1e9, 1),
                 "cublas_tflops": round(flops / t_lib / 1e9, 1)}
    print(name, out[name], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "dgemm_probe.json"), "w") as fh:
    json.dump(out, fh, indent=1)
