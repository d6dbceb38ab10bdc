"""Reference arms of ``bench.py``: the UNMODIFIED reference (``baseline/_ref``: gnngp.py, compute.py, predict.py)
timed (a) on the GPU box's host cores and (b) on the same B200 through its own torch/ATen path.

Nothing here is imported by the product.  Two entry points:

``cpu_reference_sampled``  ``--impl reference`` and the ``cpu_baseline`` key.  The reference's CPU path on the
    Reddit shape takes minutes at fraction 50 (283 s on 8 cores, ``tests/golden/big``) and does not fit host
    memory / time budgets at fraction 10, so every stage is timed with the REFERENCE'S OWN calls on a bounded row
    sample of the SAME graph the GPU arm runs on, at the full operand widths, and scaled by the inverse sampling
    rate; the two eigendecompositions run at full size and are not scaled.  What is sampled, per stage:

      sparse products   ``sigma_w * A[rows] @ X`` (compute.py:142,145): torch sparse COO x dense, rows drawn uniformly
                        from the graph, scaled by nnz(A) / nnz(sample);
      kernel layer      ``compute._ExxT_ReLU_Nystrom(Q[S], mask[S])`` (compute.py:120-127, 6-17) with S = all landmark
                        rows + r other rows, Q zero-padded to N_a + 1 columns exactly as compute.py:141 leaves it;
                        the part that is not ``torch.linalg.eigh`` is scaled by N / |S|;
      fit               ``predict.fit_Nystrom(Q[T], y[T], train[T], nugget)`` (predict.py:36-63) on |S| uniformly drawn
                        rows T; the non-eigh part is scaled by N / |T|;
      metrics           ``predict.result`` (predict.py:66-90), scaled by N / |T|.

    The thread count is set explicitly (``torch.set_num_threads``; torchrun's OMP_NUM_THREADS=1 is overridden) and
    reported.  The figure is an ESTIMATE of the reference's CPU time, labelled as such; the measured full run of
    the golden generator (fraction 50, this container's 8 cores) is recorded next to it when available.

``gpu_reference_full``  the ``gpu_reference`` key: one full, unsampled run of the unmodified ``GNNGP.predict`` on
    CUDA tensors on the same B200 and the same inputs — what a user of the reference gets from ``--device 0``
    (src/main.py:68-69, 108-116) — with samples of its outputs for a live parity check of the product.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from .make_ref import import_reference


class _TimedEigh(object):
    """Context manager that times every ``torch.linalg.eigh`` call made inside it (instrumentation only: the
    reference's code is not touched; the wrapper forwards to the real function)."""

    def __init__(self, charge_instead: float = None):
        # charge_instead: seconds to charge for an eigendecomposition that was ALREADY timed at the same size in
        # this pass (the fit's m x m block after the landmark block's N_a x N_a, m = N_a + 1): the call is then
        # answered with a placeholder decomposition (identity vectors) instead of being run a second time, to keep
        # the bounded sample bounded (a CPU eigh of 15 404 rows costs about a minute).  Stated in the record.
        self.charge_instead = charge_instead

    def __enter__(self):
        self.seconds = 0.0
        self.sizes = []
        self.substituted = 0
        self._real = torch.linalg.eigh

        def timed(*args, **kwargs):
            n = int(args[0].shape[-1])
            self.sizes.append(n)
            if self.charge_instead is not None:
                self.seconds += self.charge_instead
                self.substituted += 1
                return torch.ones(n, dtype=args[0].dtype), torch.eye(n, dtype=args[0].dtype)
            t0 = time.perf_counter()
            out = self._real(*args, **kwargs)
            self.seconds += time.perf_counter() - t0
            return out

        torch.linalg.eigh = timed
        return self

    def __exit__(self, *exc):
        torch.linalg.eigh = self._real
        return False


def _row_block_coo(edge_index: torch.Tensor, edge_attr: torch.Tensor, n: int, rows: torch.Tensor) -> torch.Tensor:
    """Sparse COO [len(rows) x n] holding the given rows of the (row-sorted) adjacency, flagged uncoalesced like
    the tensor ``datasets.get_data`` builds (src/datasets.py:328)."""
    counts = torch.bincount(edge_index[0], minlength=n)
    rowptr = torch.zeros(n + 1, dtype=torch.int64)
    torch.cumsum(counts, 0, out=rowptr[1:])
    starts, lens = rowptr[rows], counts[rows]
    total = int(lens.sum())
    local = torch.repeat_interleave(torch.arange(rows.numel()), lens)
    offs = torch.arange(total) - torch.repeat_interleave(torch.cumsum(lens, 0) - lens, lens)
    pos = torch.repeat_interleave(starts, lens) + offs
    idx = torch.stack([local, edge_index[1][pos]])
    return torch.sparse_coo_tensor(idx, edge_attr[pos], (rows.numel(), n)), total


def cpu_reference_sampled(data, land: torch.Tensor, nugget: torch.Tensor, layers: int, sigma_b: float, sigma_w: float,
                          seed: int = 0, threads: int = 0, dense_flop_budget: float = 2.5e13,
                          sparse_budget: float = 3.5e10) -> dict:
    """See the module docstring.  ``data`` holds CPU tensors (``GraphData``); returns the estimate with its parts."""
    assert layers == 2, "the sampled reference arm covers the benchmark's L = 2"
    threads = int(threads) if threads else (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    ref_compute, ref_predict, _ = import_reference()
    t_start = time.perf_counter()
    x, y = data.x, data.y
    n, f = x.shape
    na = int(land.sum())
    m = na + 1
    g = int(nugget.numel())
    c = int(y.max()) + 1 if not y.dtype.is_floating_point else 1
    gen = torch.Generator().manual_seed(1234 + seed)
    land_idx = torch.nonzero(land).flatten()

    # ---- sample sizes
    flops_per_row = 2.0 * m * na + 2.0 * na * na + 0.66 * 2.0 * m * m + 2.0 * m * g * c
    others = torch.nonzero(~land).flatten()
    r = int(min(others.numel(), max(2048, dense_flop_budget / flops_per_row - na)))
    pick = others[torch.randperm(others.numel(), generator=gen)[:r]]
    s_rows = torch.sort(torch.cat([land_idx, pick])).values
    k_total = f + na
    nnz_total = int(data.edge_index.size(1))
    want_nnz = max(2.0e5, sparse_budget / k_total)
    sp_rows_n = int(min(n, max(256, want_nnz / (nnz_total / n))))
    sp_rows = torch.sort(torch.randperm(n, generator=gen)[:sp_rows_n]).values

    stages = {}
    # ---- sparse product 1: sigma_w * A[rows] @ X  (compute.py:142)
    a_sp, nnz_sp = _row_block_coo(data.edge_index, data.edge_attr, n, sp_rows)
    t0 = time.perf_counter()
    _ = sigma_w * a_sp @ x
    stages["spmm1_sample_s"] = time.perf_counter() - t0
    del _
    # rows S of the first-layer factor, exactly as compute.py:141-142 leaves them: N_a + 1 columns, first F filled
    a_s, _nnz = _row_block_coo(data.edge_index, data.edge_attr, n, s_rows)
    q_s = torch.zeros((s_rows.numel(), m))
    q_s[:, :f] = sigma_w * a_s @ x
    q_s[:, -1] = sigma_b
    del a_s
    mask_s = land[s_rows]

    # ---- kernel layer: _ExxT_ReLU_Nystrom on rows S (Gram, arc-cosine, eigh of the landmark block, back-projection)
    with _TimedEigh() as eig1:
        t0 = time.perf_counter()
        exxt_s = ref_compute._ExxT_ReLU_Nystrom(q_s, mask_s)
        stages["layer_sample_s"] = time.perf_counter() - t0
    stages["eigh_landmark_s"] = eig1.seconds
    del q_s

    # ---- sparse product 2: sigma_w * A[rows] @ ExxT with a full-height [N x N_a] operand (rows of ExxT[S] repeated)
    operand = exxt_s[torch.randint(0, exxt_s.shape[0], (n,), generator=gen)]
    t0 = time.perf_counter()
    _ = sigma_w * a_sp @ operand
    stages["spmm2_sample_s"] = time.perf_counter() - t0
    del _, operand, a_sp

    # ---- fit + metrics on |S| uniformly drawn rows T (train share as in the whole graph)
    t_rows = torch.sort(torch.randperm(n, generator=gen)[:s_rows.numel()]).values
    q_t = torch.zeros((t_rows.numel(), m))
    q_t[:, :-1] = exxt_s[torch.randint(0, exxt_s.shape[0], (t_rows.numel(),), generator=gen)]
    q_t[:, -1] = sigma_b
    del exxt_s
    y_t = y[t_rows]
    train_t = data.train_mask[t_rows]
    reuse = eig1.seconds * (m / float(na)) ** 3 if na >= 6000 else None
    with _TimedEigh(charge_instead=reuse) as eig2:
        t0 = time.perf_counter()
        fit_t = ref_predict.fit_Nystrom(q_t, y_t, train_t, nugget)
        stages["fit_sample_s"] = time.perf_counter() - t0 + (eig2.seconds if reuse is not None else 0.0)
    stages["eigh_fit_s"] = eig2.seconds
    masks_t = {"train": train_t, "val": data.val_mask[t_rows], "test": data.test_mask[t_rows], "landmark": land[t_rows]}
    t0 = time.perf_counter()
    ref_predict.result(fit_t, y_t, masks_t)
    stages["result_sample_s"] = time.perf_counter() - t0
    del fit_t, q_t

    up_sparse = nnz_total / float(nnz_sp)
    up_s = n / float(s_rows.numel())
    est = {
        "spmm_s": (stages["spmm1_sample_s"] + stages["spmm2_sample_s"]) * up_sparse,
        "layer_rows_s": (stages["layer_sample_s"] - stages["eigh_landmark_s"]) * up_s,
        "eigh_s": stages["eigh_landmark_s"] + stages["eigh_fit_s"],
        "fit_rows_s": (stages["fit_sample_s"] - stages["eigh_fit_s"]) * up_s,
        "result_s": stages["result_sample_s"] * up_s,
    }
    total = float(sum(est.values()))
    measured = float(sum(v for k, v in stages.items() if k.endswith("_sample_s")))
    return {
        "value": round(total, 3), "unit": "s", "cores": threads, "kind": "reference", "extrapolated": True,
        "sample": ("unmodified reference functions (baseline/_ref compute.py / predict.py, torch %s CPU, fp32) on row samples "
                   "of the GPU arm's graph at full operand widths: sparse products on %d of %d rows (%d of %d nonzeros); "
                   "kernel layer on all %d landmark rows + %d others; fit and metrics on %d rows; both eigendecompositions "
                   "(n = %d, %d) at full size, unscaled%s; row-proportional stage times scaled by the inverse sampling rate; "
                   "%.1f s of CPU work measured on %d threads" % (
                       torch.__version__, sp_rows_n, n, nnz_sp, nnz_total, na, r, t_rows.numel(), na, m,
                       " (the second one charged at the first one's measured time x (m/N_a)^3, not run again)" if reuse is not None else "",
                       measured, threads)),
        "estimated_stages_s": {k: round(v, 3) for k, v in est.items()},
        "measured_stages_s": {k: round(v, 3) for k, v in stages.items()},
        "wall_s": round(time.perf_counter() - t_start, 1),
    }


def golden_full_run_record(name: str):
    """The measured FULL run of the unmodified reference recorded by tests/golden/make_golden_big.py (build
    container CPU), if that fixture exists: {"seconds", "threads"}."""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "big", name + ".npz")
    if not os.path.exists(path):
        return None
    with np.load(path) as z:
        if "reference_cpu_seconds" not in z:
            return None
        return {"fixture": "tests/golden/big/%s.npz" % name, "kernel_s": round(float(z["reference_cpu_seconds"][0]), 1),
                "total_s": round(float(z["reference_cpu_seconds"][1]), 1), "threads": int(z["reference_cpu_threads"]),
                "where": "build container, full unsampled run"}


def gpu_reference_full(data, land: torch.Tensor, nugget: torch.Tensor, layers: int, sigma_b: float, sigma_w: float,
                       sample_idx: torch.Tensor, row_idx: torch.Tensor, steps: int = 1) -> dict:
    """One warm-up on a tiny graph (library handles), then ``steps`` full runs of the unmodified reference on CUDA
    tensors.  Returns seconds per run and output samples (Q Qᵀ block, posterior-mean rows, all argmax predictions,
    metric curves) for the live parity check.  Out of memory is reported, not raised."""
    _, _, ref_gnngp = import_reference()
    dev = data.x.device
    out = {"impl": "unmodified reference (baseline/_ref gnngp.py/compute.py/predict.py) on CUDA tensors, torch %s" % torch.__version__,
           "unit": "s"}
    try:
        # library warm-up (cuSOLVER / cuSPARSE / cuBLAS handles, allocator) outside the timed region
        from gnngp_b200 import synth
        small = synth.make_graph("small", seed=1).to(dev)
        warm = ref_gnngp.GNNGP(small, device=dev, Nystrom=True)
        warm.mask["landmark"] = synth.draw_landmarks(small.train_mask, 4, seed=1)
        warm.set_hyper_param(layers, sigma_b, sigma_w, method="GCN")
        warm.predict(nugget)
        del warm, small
        model = ref_gnngp.GNNGP(data, device=dev, Nystrom=True)
        model.mask["landmark"] = land
        times = []
        for _ in range(max(1, steps)):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            model.set_hyper_param(layers, sigma_b, sigma_w, method="GCN")
            fit = model.predict(nugget)
            summ = model.get_summary()
            torch.cuda.synchronize(dev)
            times.append(time.perf_counter() - t0)
        q = model.Q[sample_idx].double()
        best = int(torch.argmax(model.result["val"]))
        out.update({
            "value": round(min(times), 4), "runs_s": [round(t, 4) for t in times], "steps_run": len(times),
            "best": {k: float(summ[k]) for k in ("train", "val", "test", "nugget")},
            "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 1e9, 1),
        })
        samples = {"ksamp": (q @ q.T).cpu().numpy(), "best": best, "fit_best": fit[best][row_idx].cpu().numpy(),
                   "fit_mid": fit[50][row_idx].cpu().numpy(), "pred_best": torch.argmax(fit[best], 1).cpu().numpy(),
                   "curves": {k: v.numpy() for k, v in model.result.items()}}
        top2 = torch.topk(fit[best], 2, dim=1).values
        samples["margin_best"] = (top2[:, 0] - top2[:, 1]).cpu().numpy()
        del model, fit, q
        torch.cuda.empty_cache()
        return out, samples
    except torch.cuda.OutOfMemoryError as exc:
        torch.cuda.empty_cache()
        out["unavailable"] = "CUDA out of memory in the reference's torch path: %s" % (str(exc).split("\n")[0][:160],)
        return out, None
