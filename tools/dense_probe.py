"""GPU-box probe: dense contraction engines at Reddit-shape sizes (time + accuracy), eigh fp32 vs fp64."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnngp_b200.ops import default_ops  # noqa: E402

dev = torch.device("cuda:0")
ops = default_ops()
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


def rel(a, b):
    return float((a.double() - b).norm() / b.norm())


# ---- accuracy of the engines on same-sign and random-sign data, short and long K
g = torch.Generator(device=dev).manual_seed(0)
for label, M, N, K, positive in (("rand K=3107", 2048, 1024, 3107, False), ("pos K=3107", 2048, 1024, 3107, True),
                                 ("rand K=153431", 1024, 512, 153431, False), ("pos K=153431", 1024, 512, 153431, True)):
    a = torch.rand(M, K, device=dev, generator=g) if positive else torch.randn(M, K, device=dev, generator=g)
    b = torch.rand(N, K, device=dev, generator=g) if positive else torch.randn(N, K, device=dev, generator=g)
    want = a.double() @ b.double().T
    row = {"torch_fp32": rel(a @ b.T, want)}
    for engine, name in ((0, "simt"), (1, "tc3_chunked_128x256")):
        ops.set_gemm_engine(engine)
        row[name] = rel(ops.gemm_nt(a, b), want)
    out["accuracy " + label] = row
    print("accuracy", label, row, flush=True)
    del a, b, want

# ---- timing at Reddit-shape f=50 sizes
n, na, nt, G, C = 232965, 3106, 153431, 101, 41
m = na + 1
q = torch.randn(n, m, device=dev, generator=g)
l = q[:na].contiguous()
s, t = ops.row_norms(q), ops.row_norms(l)
wt = torch.randn(na, na, device=dev, generator=g)
bt = torch.randn(G * C, m, device=dev, generator=g)
qbt = torch.randn(m, nt, device=dev, generator=g)
e = ops.empty(n, na, dev)
fitted = torch.empty((G, n, C), device=dev)
for engine, name in ((1, "tc3_chunked_128x256"),):
    ops.set_gemm_engine(engine)
    row = {}
    row["gram_arccos %dx%dx%d" % (n, na, m)] = timeit(lambda: ops.gram_arccos(q, l, s, t, out=e))
    row["backproject %dx%dx%d" % (n, na, na)] = timeit(lambda: ops.gemm_nt(e, wt))
    row["train_gram %dx%dx%d" % (m, m, nt)] = timeit(lambda: ops.gemm_nt(qbt, qbt))
    row["train_syrk_lower_splitk %dx%dx%d" % (m, m, nt)] = timeit(lambda: ops.syrk(qbt))
    row["nugget_sweep %dx%dx%d" % (n, G * C, m)] = timeit(lambda: ops.nugget_sweep(q, bt, G, C, out=fitted))
    row["nugget_sweep_as_plain_gemm"] = timeit(lambda: ops.gemm_nt(q, bt))
    row["nugget_sweep C=40"] = timeit(lambda: ops.nugget_sweep(q, bt[:G * 40], G, 40, out=fitted.view(-1)[:G * n * 40].view(G, n, 40)))
    out["time_ms " + name] = row
    print(name, json.dumps(row), flush=True)
ops.set_gemm_engine(2)
fl = {"gram": 2.0 * n * na * m, "backproject": 2.0 * n * na * na, "train_gram": 2.0 * m * m * nt, "sweep": 2.0 * n * G * C * m}
out["flops"] = fl
del q, e, fitted, qbt
torch.cuda.empty_cache()

# ---- eigensolver
for size in ():
    x = torch.randn(size, size + 64, device=dev, generator=g)
    sym = (x @ x.T) / size
    row = {}
    row["fp32_ms"] = timeit(lambda: torch.linalg.eigh(sym), reps=1)
    s64 = sym.double()
    row["fp64_ms"] = timeit(lambda: torch.linalg.eigh(s64), reps=1)
    out["eigh n=%d" % size] = row
    print("eigh", size, row, flush=True)
    del x, sym, s64
    torch.cuda.empty_cache()

with open(os.path.join(ROOT, "gpurun_out", "dense_probe.json"), "w") as fh:
    json.dump(out, fh, indent=1)
