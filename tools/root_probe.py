"""GPU-box probe: the Nystrom root of the benchmark step (Reddit shape, fraction 10 by default) by the Krylov route
(gnngp_b200/krylov_root.py) against the eigen route — time (CUDA events), per-phase times, Krylov dimension, Ritz count,
agreement of Q Qᵀ on a sample of landmark rows, and cuBLAS DGEMM rates at the route's shapes for reference."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnngp_b200 import krylov_root, synth  # noqa: E402
from gnngp_b200.gnngp import GNNGP  # noqa: E402
from gnngp_b200.ops import default_ops  # noqa: E402
from gnngp_b200.pipeline import Pipeline  # noqa: E402

dev = torch.device("cuda:0")
ops = default_ops()
ONCE = "--once" in sys.argv          # one Krylov call inside an NVTX range "timed" (for an ncu launch list), nothing else
fractions = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [10]
out = {}


def timeit(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, r


data = synth.make_graph("reddit", seed=0).to(dev)
for fraction in fractions:
    seen = []
    real = Pipeline._root_factors

    def spy(self, kmm, _real=real):
        seen.append(kmm.clone())
        return _real(self, kmm)

    Pipeline._root_factors = spy
    os.environ["GNNGP_ROOT_SOLVER"] = "eigh"
    model = GNNGP(data, device=dev, Nystrom=True)
    model.mask["landmark"] = synth.draw_landmarks(data.train_mask.cpu(), fraction, seed=0).to(dev)
    model.set_hyper_param(2, 0.0, 1.0, method="GCN")
    model.get_kernel()
    Pipeline._root_factors = real
    del os.environ["GNNGP_ROOT_SOLVER"]
    kmm = max(seen, key=lambda t: t.shape[0])
    del model, seen
    ops.release_workspace()
    torch.cuda.empty_cache()
    n = kmm.shape[0]
    rec = {"n": n}

    if ONCE:
        krylov_root.root_factors(ops, kmm, {})
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("timed")
        krylov_root.root_factors(ops, kmm, {})
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        continue
    pipe = Pipeline(ops)
    os.environ["GNNGP_ROOT_SOLVER"] = "eigh"
    ms_eig, (w_e, l_e) = timeit(lambda: pipe._root_factors(kmm), reps=1)
    os.environ["GNNGP_ROOT_SOLVER"] = "krylov"
    stats = {}
    ms_kry, got = timeit(lambda: krylov_root.root_factors(ops, kmm, stats), reps=2)
    rec.update(eigh_route_ms=round(ms_eig, 1), krylov_route_ms=round(ms_kry, 1), converged=got is not None,
               krylov_dim=stats.get("krylov_dim"), ritz_above_clamp=stats.get("ritz_above_clamp"),
               residual_over_allowed=stats.get("residual_over_allowed"))
    tstats = {"time": True}
    krylov_root.root_factors(ops, kmm, tstats)
    rec["phase_ms"] = {k: round(v, 1) for k, v in tstats.get("phase_ms", {}).items()}
    if got is not None:
        w_k, l_k = got
        rows = kmm[:1024].double()
        qe, qk = rows @ w_e.double().T, rows @ w_k.double().T
        ref = qe @ qe.T
        rec["qqT_rel_rows"] = float((qk @ qk.T - ref).norm() / ref.norm())
        le, lk = l_e[:1024].double(), l_k[:1024].double()
        ref = le @ le.T
        rec["qqT_rel_landmark_rows"] = float((lk @ lk.T - ref).norm() / ref.norm())
        ref = qe @ le.T
        rec["qqT_rel_cross"] = float((qk @ lk.T - ref).norm() / ref.norm())
        del qe, qk, w_k, l_k
    del w_e, l_e, got
    # library DGEMM at the route's shapes (what the fp64 tensor path can give)
    m64 = ops.sym_f64(kmm)
    v = torch.randn(n, 128, device=dev, dtype=torch.float64)
    c = torch.empty(n, 128, device=dev, dtype=torch.float64)
    ms_own, _ = timeit(lambda: ops.dgemm(m64, v, out=c), reps=5)
    ms_lib, _ = timeit(lambda: torch.mm(m64, v, out=c), reps=5)
    rec["dgemm_n_x_128_x_n_ms"] = {"own": round(ms_own, 2), "cublas": round(ms_lib, 2), "own_tflops": round(2 * n * n * 128 / ms_own / 1e9, 1),
                                   "cublas_tflops": round(2 * n * n * 128 / ms_lib / 1e9, 1)}
    p = torch.randn(n, 128, device=dev, dtype=torch.float64)
    big = torch.zeros(n, n, device=dev, dtype=torch.float64)
    ms_own, _ = timeit(lambda: ops.dgemm(p, p.T, out=big, alpha=-1.0, beta=1.0, lower=True), reps=3)
    ms_lib, _ = timeit(lambda: torch.addmm(big, p, p.T, alpha=-1.0, out=big), reps=3)
    rec["rank128_update_ms"] = {"own_lower": round(ms_own, 2), "cublas_full": round(ms_lib, 2)}
    ms_own, _ = timeit(lambda: ops.potrf_f64(m64.clone(), torch.zeros(1, dtype=torch.int32, device=dev)), reps=1)
    ms_lib, _ = timeit(lambda: torch.linalg.cholesky(m64), reps=1)
    rec["cholesky_ms"] = {"own": round(ms_own, 1), "cusolver": round(ms_lib, 1)}
    out["fraction_%d" % fraction] = rec
    print(fraction, json.dumps(rec), flush=True)
    del m64, big, kmm
    ops.release_workspace()
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "root_probe.json"), "w") as fh:
    json.dump(out, fh, indent=1)
