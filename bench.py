#!/usr/bin/env python
"""bench.py — GCNGP-X fit+predict seconds on a synthetic Reddit-shape graph (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--shape reddit] [--fraction 10] [--impl reference]

One "step" = what the reference times per run (``src/main.py:109-116``, landmark draw injected):
``set_hyper_param`` + ``predict`` (kernel recursion, fit for 101 nuggets, metrics) + ``get_summary``.
Default workload = the north-star point: Reddit shape, GCN, L = 2, fraction 10 (15 404 landmarks); the
fraction-50 point of the sweep (figure4.sh:2) is measured too and reported under ``sweep``.

``value``   seconds per step with the graph resident in HBM (CSR built once, like the model object the
            reference keeps across runs), CUDA events, max over ranks, K steps after W warm-ups.
``e2e``     seconds per step through the public ``GNNGP`` API from pinned HOST buffers: host→device copy
            of features / labels / edges / masks, model construction (COO→CSR), predict, summary on host.
``roofline``the dominant kernel of this library (layer-2 CSR SpMM): algorithmic work ÷ its
            CUDA-event time, against the binding peak of ``max(bytes / HBM, flops / fp32-FMA)``.
``config.parity_vs_reference``  the timed run's outputs against the UNMODIFIED reference: live against the
            reference's own torch path on this GPU (``gpu_reference``) at the headline fraction, and against the
            golden fixture ``tests/golden/big/reddit-GCN-nys50-L2.npz`` (reference CPU run) at fraction 50.
``gpu_reference``  the unmodified reference (``baseline/_ref``) on CUDA tensors on the same B200, same inputs.
``cpu_baseline``   the unmodified reference's functions on the host cores on a bounded row sample (estimate).
``--impl reference`` prints the CPU arm alone (rank 0 only under torchrun).

The graph is generated on the HOST (seeded, deterministic) so that it is exactly the graph the golden fixture
was made from; under torchrun rank 0 generates it and broadcasts it.

N > 1 (torchrun, one rank per GPU): adjacency and factor are row-sharded, each layer all-gathers the
factor rows over NCCL and the fit all-reduces ``Q_bᵀQ_b`` (gnngp_b200/shard.py); total work is fixed →
``"scaling": "strong"``.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "gcngp_x_fit_predict_seconds"
UNIT = "s"
GOLDEN_F50 = "reddit-GCN-nys50-L2"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="reddit")
    ap.add_argument("--fraction", type=int, default=10)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--sigma_b", type=float, default=0.0)
    ap.add_argument("--sigma_w", type=float, default=1.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the fraction-50 point and its golden parity check")
    ap.add_argument("--stream-upload", default="0", choices=["0", "1"],
                    help="e2e leg: GNNGP_STREAM_UPLOAD (opt-in piecewise upload of the pinned edge list); default = the "
                         "library's default configuration (off)")
    ap.add_argument("--cpu-threads", type=int, default=0, help="host threads of the CPU arm (0 = all cores)")
    ap.add_argument("--breakdown", default="", help="write the per-operator timing table to this JSON file")
    return ap.parse_args()


def workload_name(args, fraction=None) -> str:
    return "GCNGP-X (GCN, L=%d, fraction=%d, sigma_b=%g, sigma_w=%g, 101 nuggets) on synthetic %s-shape graph" % (
        args.layers, fraction or args.fraction, args.sigma_b, args.sigma_w, args.shape)


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler(object):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2]))
            except ValueError:
                continue
            for label, flag in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(label)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------------------------
def host_graph(args):
    """The seeded synthetic graph, generated on the host: the same tensors on every machine (this is the graph
    the golden fixtures of tests/golden/big were computed from; checked through the input checksum)."""
    from gnngp_b200 import synth
    return synth.make_graph(args.shape, seed=args.seed, device=None, normalization="sym")


def device_graph(args, dev, world: int, rank: int):
    """The graph on this rank's GPU.  One rank generates (host, ~25 s at the Reddit shape) and broadcasts."""
    import torch.distributed as dist
    from gnngp_b200.datasets import GraphData
    fields = ("x", "y", "edge_index", "edge_attr", "train_mask", "val_mask", "test_mask")
    if world == 1:
        host = host_graph(args)
        return host.to(dev), host
    meta = [None]
    host = None
    if rank == 0:
        host = host_graph(args)
        meta = [{k: (tuple(getattr(host, k).shape), getattr(host, k).dtype) for k in fields} | {"c": host.num_classes}]
    dist.broadcast_object_list(meta, src=0)
    tensors = {}
    for k in fields:
        shape, dtype = meta[0][k]
        t = getattr(host, k).to(dev) if rank == 0 else torch.empty(shape, dtype=torch.uint8 if dtype == torch.bool else dtype, device=dev)
        if dtype == torch.bool:
            t = t.to(torch.uint8)
        dist.broadcast(t, src=0)
        tensors[k] = t.to(torch.bool) if dtype == torch.bool else t
    return GraphData(tensors["x"], tensors["y"], tensors["edge_index"], tensors["edge_attr"], tensors["train_mask"],
                     tensors["val_mask"], tensors["test_mask"], num_classes=meta[0]["c"]), host


# --------------------------------------------------------------------------------------------
# parity of a finished run against reference outputs (golden fixture or live reference run)
# --------------------------------------------------------------------------------------------
def rel_fro(a, b) -> float:
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def collect_samples(model, sample_idx, row_idx, best=None, fit=None, ksamp=None, curves=None) -> dict:
    """Small samples of a finished single-GPU run: Q Qᵀ block, posterior-mean rows, all argmax predictions, curves.
    ``best`` = the nugget index to sample at (the REFERENCE's selected nugget when comparing with a reference run)."""
    if ksamp is None:
        q = model.Q[sample_idx].double()
        ksamp = (q @ q.T).cpu().numpy()
    if curves is None:
        curves = {k: v.numpy() for k, v in model.result.items()}
    if best is None:
        best = int(np.argmax(curves["val"]))
    fit = model.fit if fit is None else fit
    return {"ksamp": ksamp, "best": best, "fit_best": fit[best][row_idx].cpu().numpy(),
            "fit_mid": fit[50][row_idx].cpu().numpy(), "pred_best": torch.argmax(fit[best], 1).cpu().numpy(), "curves": curves}


def parity_record(ours: dict, ref: dict, source: str) -> dict:
    """qqT_rel / fit_rel / argmax_flips of ``ours`` against reference samples (same sample indices, same nugget).

    Posterior means: the reference's own fp32 run differs from its fp64 run by 7e-3 (nugget 0.1) to 6e-2 (nugget
    1e-3) on the Reddit shape (tests/golden/big, ``fit_*64``) — two fp32 evaluations are two draws of rounding noise
    times the conditioning of ``Q_b^T Q_b + eps d I`` — so ``fit_rel`` is taken against the fp64 run when the source
    has one and bounded by max(1e-4, 6 x that self-error); against an fp32-only source it is reported, not bounded.
    A flip counts as "clear" when the reference's own top-2 margin at that node exceeds its fp32 noise (fp64 companion)
    or 1e-3 of its largest posterior mean."""
    flips = ours["pred_best"].astype(np.int64) != np.asarray(ref["pred_best"]).astype(np.int64)
    rec = {"source": source, "nugget_index": int(ref["best"]), "qqT_rel": rel_fro(ours["ksamp"], ref["ksamp"]),
           "argmax_flips": int(flips.sum()), "argmax_nodes": int(flips.size)}
    fit_ok = True
    if "fit_best64" in ref:
        rec["qqT_rel_fp64"] = rel_fro(ours["ksamp"], ref["ksamp64"])
        for key, tag in (("fit_best", "fit_rel"), ("fit_mid", "fit_rel_nugget_0.1")):
            floor = rel_fro(ref[key], ref[key + "64"])
            rec[tag] = rel_fro(ours[key], ref[key + "64"])
            rec[tag + "_reference_fp32_self_error"] = float("%.3e" % floor)
            rec[tag + "_vs_fp32_run"] = float("%.3e" % rel_fro(ours[key], ref[key]))
            fit_ok = fit_ok and rec[tag] <= max(1e-4, 6.0 * floor)
        rec["fit_rel_against"] = "the reference's fp64 run"
        noise = float(np.max(np.abs(ref["fit_best"].astype(np.float64) - ref["fit_best64"])))
    else:
        rec["fit_rel"] = rel_fro(ours["fit_best"], ref["fit_best"])
        rec["fit_rel_nugget_0.1"] = rel_fro(ours["fit_mid"], ref["fit_mid"])
        rec["fit_rel_against"] = "the reference's fp32 run (no fp64 companion: reported, not bounded)"
        noise = None
    if "margin_best" in ref:
        scale = float(ref.get("fit_scale", np.abs(ref["fit_best"]).max()))
        tol = (2.0 * noise + 1e-5 * scale) if noise is not None else 1e-3 * scale
        rec["argmax_flips_clear_margin"] = int((flips & (np.asarray(ref["margin_best"]) > tol)).sum())
    curves = ref.get("curves") or {k: ref["result_" + k] for k in ("train", "val", "test")}
    diffs = {k: np.abs(ours["curves"][k] - np.asarray(curves[k])) for k in ("train", "val", "test")}
    rec["curve_max_abs_diff"] = float(max(d.max() for d in diffs.values()))
    best = int(ref["best"])
    rec["metrics_at_selected_nugget_abs_diff"] = float(max(d[best] for d in diffs.values()))
    # accuracies are compared nugget by nugget; a nugget where either side's posterior solve sits next to a pole of
    # 1 / (w + eps d) (a rounding-noise eigenvalue of Q_b^T Q_b close to -eps d) is noise in BOTH implementations
    rec["curve_nuggets_within_2e-3"] = int(sum(bool(all(d[i] <= 2e-3 for d in diffs.values())) for i in range(len(diffs["val"]))))
    rec["curve_val_ours_first_last"] = [round(float(v), 5) for v in list(ours["curves"]["val"][:6]) + list(ours["curves"]["val"][-2:])]
    rec["curve_val_reference_first_last"] = [round(float(v), 5) for v in list(np.asarray(curves["val"])[:6]) + list(np.asarray(curves["val"])[-2:])]
    if "fit_best64" in ref:
        rec["ok"] = bool(rec["qqT_rel"] <= 1e-4 and fit_ok and rec.get("argmax_flips_clear_margin", 0) == 0
                         and rec["metrics_at_selected_nugget_abs_diff"] <= 2e-3 and rec["curve_nuggets_within_2e-3"] >= 0.9 * len(diffs["val"]))
    else:
        # fp32-only reference (its fp64 run does not fit the GPU at this size).  Its own accuracy curve collapses at the
        # smallest nuggets (rounding-noise eigenvalues of its fp32 Gram / fp32 eigh next to -eps d; the fraction-50 golden
        # shows the same run differing from its fp64 companion by 6e-2 there), so curves are compared on the reference's
        # PLATEAU (nuggets whose validation accuracy is within 0.01 of its best) and a flip rate of 1e-5 is tolerated.
        ref_val = np.asarray(curves["val"])
        plateau = ref_val >= ref_val.max() - 0.01
        on_plateau = int(sum(bool(all(d[i] <= 2e-3 for d in diffs.values())) for i in range(len(ref_val)) if plateau[i]))
        rec["reference_plateau_nuggets"] = int(plateau.sum())
        rec["curve_plateau_nuggets_within_2e-3"] = on_plateau
        # argmax at the nugget the REFERENCE selected.  When that nugget lies at the edge of the reference's noisy range
        # (its validation curve is still climbing there: curve_val_reference_first_last) its posterior means carry the
        # fp32 noise this record measures as fit_rel, and nodes whose top-2 margin is below that noise flip.  Beyond the
        # 1e-5 rate, every flip must be EXPLAINED by the measured discrepancy: reference margin <= 6 x the rms difference
        # of the two runs' posterior means at that nugget (sampled rows); the count and the largest margin are reported.
        few = rec["argmax_flips"] <= max(1, int(1e-5 * flips.size))
        explained = few
        if not few and "margin_best" in ref:
            rms = float(np.sqrt(np.mean((ours["fit_best"].astype(np.float64) - np.asarray(ref["fit_best"], dtype=np.float64)) ** 2)))
            margins = np.asarray(ref["margin_best"])[flips]
            rec["fit_abs_rms_diff_at_selected_nugget"] = float("%.3e" % rms)
            rec["argmax_flips_max_reference_margin"] = float("%.3e" % float(margins.max()))
            rec["argmax_flips_rate"] = float("%.3e" % (float(flips.sum()) / flips.size))
            explained = bool(margins.max() <= 6.0 * rms and flips.sum() <= 1e-3 * flips.size)
        rec["argmax_flips_explained_by_reference_noise"] = bool(explained)
        rec["ok"] = bool(rec["qqT_rel"] <= 1e-4 and explained and rec["fit_rel_nugget_0.1"] <= 5e-2
                         and rec["metrics_at_selected_nugget_abs_diff"] <= 2e-3 and on_plateau >= 0.95 * plateau.sum())
    for k in ("qqT_rel", "qqT_rel_fp64", "fit_rel", "fit_rel_nugget_0.1", "curve_max_abs_diff"):
        if k in rec:
            rec[k] = float("%.3e" % rec[k])
    return rec


def fit_self_check(model, data, nugget, row_idx, indices=(0, 4, 50)) -> dict:
    """The timed run's posterior means against an fp64 DENSE solve on its own kernel factor (torch fp64 on the GPU: the
    Gram matrix of the train rows accumulated in fp64, one Cholesky solve per checked nugget): isolates the fit route
    (3xTF32 Gram + block Lanczos / band reduction) from the noise of an fp32-only reference.  Checker code, after the
    timed region."""
    q = model.Q
    dev = q.device
    m = q.shape[1]
    train = data.train_mask.to(dev).nonzero().squeeze(1)
    classes = int(model.fit.shape[-1])
    gram = torch.zeros((m, m), dtype=torch.float64, device=dev)
    rhs = torch.zeros((m, classes), dtype=torch.float64, device=dev)
    for r0 in range(0, train.numel(), 8192):
        rows = train[r0:r0 + 8192]
        blk = q[rows].double()
        gram.addmm_(blk.T, blk)
        rhs.addmm_(blk.T, torch.nn.functional.one_hot(data.y.to(dev)[rows], classes).double())
    d = 0.0
    for r0 in range(0, q.shape[0], 16384):
        d += float((q[r0:r0 + 16384].double() ** 2).sum())
    d /= q.shape[0]
    qs = q[row_idx].double()
    eye = torch.eye(m, dtype=torch.float64, device=dev)
    out = {}
    for gi in indices:
        factor, bad = torch.linalg.cholesky_ex(gram + float(nugget[gi]) * d * eye)
        if int(bad) != 0:
            out["nugget_%d" % gi] = None
            continue
        want = qs @ torch.cholesky_solve(rhs, factor)
        got = model.fit[gi][row_idx].double()
        out["nugget_%d" % gi] = float("%.3e" % float((got - want).norm() / want.norm()))
    del gram, eye, qs
    return out


def load_big_golden(name: str):
    path = os.path.join(ROOT, "tests", "golden", "big", name + ".npz")
    return dict(np.load(path)) if os.path.exists(path) else None


# --------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference's functions on a bounded sample (baseline/ref_arm.py)
# --------------------------------------------------------------------------------------------
def cpu_arm(args, host, land_host) -> dict:
    from baseline import ref_arm
    from gnngp_b200 import synth
    rec = ref_arm.cpu_reference_sampled(host, land_host, synth.nugget_grid(), args.layers, args.sigma_b, args.sigma_w,
                                        seed=args.seed, threads=args.cpu_threads)
    full = ref_arm.golden_full_run_record("reddit-GCN-nys%d-L%d" % (args.fraction, args.layers)) if args.shape == "reddit" else None
    if full is not None:
        rec["measured_full_run"] = full
    return rec


def run_reference_arm(args):
    """``--impl reference``: the CPU arm alone, ONE measured pass whatever --steps says (stated in ``steps_run``)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gnngp_b200 import synth
    t0 = time.perf_counter()
    try:
        host = host_graph(args)
        land = synth.draw_landmarks(host.train_mask, args.fraction, seed=args.seed)
        base = cpu_arm(args, host, land)
    except Exception as exc:
        print(json.dumps({"impl": "reference", "unavailable": "CPU reference arm failed: %r" % (exc,)}), flush=True)
        return
    value = base["value"]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "steps_run": 1, "warmup_run": 0, "ms_per_step": value * 1e3, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "device": "cpu", "threads": base["cores"],
                       "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"), "wall_s": round(time.perf_counter() - t0, 1),
                       "note": "one bounded-sample pass of the unmodified reference functions on the same seeded graph as the "
                               "GPU arm; an estimate of the full CPU run, see cpu_baseline.sample"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def ncu_traffic(kernel: str, info: dict, k: int):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel from the
    committed `ncu --set full` capture (profiles/ncu_traffic.json) — only when that capture was taken at the
    SAME launch shape as the one timed here, else null."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            table = json.load(fh)
        for rec in table.get(kernel, []) if isinstance(table.get(kernel), list) else [table[kernel]]:
            if int(rec["nnz"]) == int(info["nnz"]) and int(rec["n"]) == int(info["nodes"]) and int(rec["k"]) == int(k):
                return float(rec["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def load_peaks() -> dict:
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0, "source": "fallback (B200_PROFILING.md)"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks.update(json.load(fh)); peaks["source"] = "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return peaks


def spmm_roofline(spmm_calls, info, peaks, kernel_names) -> dict:
    """Roofline of the sparse product with the most nnz x k (the layer-2 propagation): regime from
    max(bytes / HBM peak, flops / fp32 FMA peak), SURVEY.md §8d."""
    work = max(i[1] * i[2] for i, _ in spmm_calls)
    big = [(i, ms) for i, ms in spmm_calls if i[1] * i[2] == work]
    (rows, nnz, k), _ = big[0]
    ms = sum(m for _, m in big) / len(big)
    n_cols = info["nodes"]
    alg_bytes = 8.0 * nnz + 4.0 * (rows + 1) + 4.0 * n_cols * k + 4.0 * rows * k
    flops = 2.0 * nnz * k
    fma_peak = 148 * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12          # fp32 CUDA-core TFLOP/s at max clock
    t_hbm = alg_bytes / (peaks["hbm_gbs"] * 1e9)
    t_fma = flops / (fma_peak * 1e12)
    hbm_view = {"achieved_gbs": round(alg_bytes / (ms * 1e-3) / 1e9, 2), "peak_gbs": peaks["hbm_gbs"],
                "frac": round(alg_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 5)}
    kernel = kernel_names.get(k, "spmm kernel")
    common = {"kernel": "%s (layer-2 propagation A @ Ny, rows=%d nnz=%d k=%d)" % (kernel, rows, nnz, k), "ms_per_launch": round(ms, 3),
              "algorithmic_bytes": alg_bytes, "algorithmic_flops": flops, "t_roof_ms": round(max(t_hbm, t_fma) * 1e3, 3),
              "t_hbm_ms": round(t_hbm * 1e3, 3), "t_fma_ms": round(t_fma * 1e3, 3), "peak_source": peaks["source"],
              "traffic": ncu_traffic(kernel, info, k), "l2_gather_bytes": 4.0 * nnz * k,
              "l2_gather_tbs": round(4.0 * nnz * k / (ms * 1e-3) / 1e12, 3), "hbm_view": hbm_view}
    if t_hbm >= t_fma:
        common.update({"bound": "hbm", "achieved": hbm_view["achieved_gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                       "frac": hbm_view["frac"]})
    else:
        tf = flops / (ms * 1e-3) / 1e12
        common.update({"bound": "fp32-fma", "achieved": round(tf, 2), "peak": round(fma_peak, 1), "unit": "TFLOP/s",
                       "frac": round(tf / fma_peak, 4),
                       "note": "average degree %.0f puts this launch above the HBM/FMA crossover (SURVEY §8d): bound = fp32 FMA issue; "
                               "the resource it actually saturates is the L2->SM gather path (l2_gather_tbs), 4 bytes per FMA" % (nnz / max(rows, 1))})
    return common


def run_ours(args):
    import torch.distributed as dist
    from gnngp_b200 import synth
    from gnngp_b200.gnngp import GNNGP
    from gnngp_b200.ops import default_ops
    from gnngp_b200.profiling import OpTimer
    import cases

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the GNNGP hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops = default_ops()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_setup = time.perf_counter()
    data, host = device_graph(args, dev, world, rank)
    land = synth.draw_landmarks(data.train_mask.cpu(), args.fraction, seed=args.seed).to(dev)      # host RNG: same set everywhere
    nugget = synth.nugget_grid(dev)
    info = synth.describe(data)
    info["landmarks"] = int(land.sum())
    setup_s = time.perf_counter() - t_setup

    def make_model():
        if world == 1:
            return GNNGP(data, device=dev, Nystrom=True)
        from gnngp_b200.shard import ShardedGNNGP
        return ShardedGNNGP(data, device=dev, Nystrom=True)

    model = make_model()
    model.mask["landmark"] = land

    def step():
        model.set_hyper_param(args.layers, args.sigma_b, args.sigma_w, method="GCN")
        model.predict(nugget)
        return model.get_summary()

    timer = OpTimer(ops).install()
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = ops.launch_count()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    timer.reset()
    timer.enabled = True
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.nvtx.range_push("timed")          # ncu --nvtx --nvtx-include "timed/" selects exactly this region
    ev0.record()
    for _ in range(args.steps):
        summary = step()
    ev1.record()
    barrier()
    torch.cuda.nvtx.range_pop()
    elapsed = torch.tensor([ev0.elapsed_time(ev1) / 1e3], device=dev)
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    seconds = float(elapsed) / args.steps
    clock_info = clocks.stop() if rank == 0 else None
    launches_total = ops.launch_count() - launches0                  # this library's kernels inside the timed region
    launches = launches_total // max(args.steps, 1)
    timer.enabled = False
    table = timer.summary()
    spmm_calls = timer.calls("spmm") + timer.calls("spmm_panels")
    fit_solver = getattr(ops, "last_fit_solver", None) or "eigh"
    fit_stats = dict(getattr(ops, "last_fit_stats", None) or {}) if fit_solver == "krylov" or getattr(ops, "last_fit_stats", None) else None
    root_solver = {"solver": getattr(ops, "last_root_solver", None) or "eigh"}
    root_solver.update({k: v for k, v in (getattr(ops, "last_root_stats", None) or {}).items() if k != "phase_ms"})
    best = {k: float(summary[k]) for k in ("train", "val", "test", "nugget")}

    peaks = load_peaks()
    roof = spmm_roofline(spmm_calls, info, peaks, getattr(ops, "spmm_kernel_by_k", {})) if spmm_calls else None

    # ---- tensor-core side: the largest tcgen05 launch of the step, plus the other two fused contractions; op times
    # include the operand split pre-pass
    roof_tensor = None
    tf32_peak = peaks.get("bf16_tflops", 1590.0) / 2.0
    tf32_sustained = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)) / 2.0   # under the power cap: a long step

    def tensor_entry(op_name, kernel_name, share=1.0):
        calls = [(i, ms) for i, ms in timer.calls(op_name) if i is not None]
        if not calls:
            return None
        work = max(i[0] * i[1] * i[2] for i, _ in calls)
        big = [(i, ms) for i, ms in calls if i[0] * i[1] * i[2] == work]
        (gm, gn, gk), _ = big[0]
        gms = sum(m for _, m in big) / len(big)
        useful = share * 2.0 * gm * gn * gk / (gms * 1e-3) / 1e12
        return {"bound": "tensor", "kernel": "%s (%d x %d x %d, 3xTF32)" % (kernel_name, gm, gn, gk),
                "achieved": round(useful, 1), "issued_tflops": round(3 * useful, 1), "peak": round(tf32_peak, 1),
                "peak_source": "%s bf16 burst / 2 (TF32)" % peaks["source"], "unit": "TFLOP/s",
                "frac": round(3 * useful / tf32_peak, 4), "peak_sustained": round(tf32_sustained, 1),
                "frac_sustained": round(3 * useful / tf32_sustained, 4), "ms_per_launch": round(gms, 3), "flops": share * 2.0 * gm * gn * gk,
                "note": "useful flops x 3 MMAs per product (fp32-equivalent split) against the TF32 rate"}

    entries = [e for e in (tensor_entry("nugget_sweep", "gemm_nt_tc3_kernel<EpiSweep>"),
                           tensor_entry("gemm_nt", "gemm_nt_tc3_kernel<EpiPlain>"),
                           # the symmetric product computes the tiles that touch the lower triangle: (n^2 / 2 + n * 128) * K
                           tensor_entry("syrk", "gemm_nt_tc3_kernel<lower>", share=0.5 + 128.0 / max(info["landmarks"], 1)),
                           tensor_entry("gram_arccos", "gemm_nt_tc3_kernel<EpiArccos>")) if e is not None]
    if entries:
        entries.sort(key=lambda e: -e["flops"])
        roof_tensor = entries[0]
        roof_tensor["others"] = [{k: e[k] for k in ("kernel", "achieved", "frac", "ms_per_launch")} for e in entries[1:]]

    # ---- fp64 tensor side: the Lanczos product M V (n x 128 x n) of the Krylov Nystrom root and the rank-256 symmetric
    # update of the band reduction, timed here in isolation (CUDA events) at the step's landmark count
    roof_fp64 = None
    if world == 1 and hasattr(ops, "dgemm") and info["landmarks"] >= 1024:
        try:
            nl = int(info["landmarks"])
            m64 = torch.randn((nl, nl), dtype=torch.float64, device=dev)
            v64 = torch.randn((nl, 256), dtype=torch.float64, device=dev)
            o64 = torch.empty((nl, 128), dtype=torch.float64, device=dev)

            def timed(fn, reps=5):
                fn()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    fn()
                b.record()
                torch.cuda.synchronize()
                return a.elapsed_time(b) / reps
            t_mv = timed(lambda: ops.dgemm(m64, v64[:, :128], out=o64))
            t_up = timed(lambda: ops.dgemm(v64, v64.T, out=m64, alpha=-1.0, beta=1.0, lower=True))
            fp64_peak = 40.0                                             # B200 fp64 tensor, nominal (no measured figure in MEASURED_PEAKS.json)
            mv = 2.0 * nl * nl * 128 / (t_mv * 1e-3) / 1e12
            up = 2.0 * (nl * nl / 2.0 + nl * 128.0) * 256 / (t_up * 1e-3) / 1e12
            roof_fp64 = {"bound": "fp64-tensor", "kernel": "dgemm_fast_kernel (Lanczos product %d x 128 x %d)" % (nl, nl),
                         "achieved": round(mv, 1), "peak": fp64_peak, "peak_source": "nominal B200 fp64 tensor rate (cuBLAS reaches 34 on this product)",
                         "unit": "TFLOP/s", "frac": round(mv / fp64_peak, 4), "ms_per_launch": round(t_mv, 3),
                         "others": [{"kernel": "dgemm_fast_kernel (lower rank-256 update %d x %d)" % (nl, nl), "achieved": round(up, 1),
                                     "frac": round(up / fp64_peak, 4), "ms_per_launch": round(t_up, 3)}]}
            del m64, v64, o64
        except RuntimeError:
            roof_fp64 = None
        torch.cuda.empty_cache()

    # ---- samples of the timed run's outputs for the parity record (single GPU: Q and fit hold all rows)
    n = info["nodes"]
    sample_idx = torch.from_numpy(cases.sample_index(n, args.seed)).to(dev)
    row_idx = torch.from_numpy(cases.big_row_sample(n, args.seed)).to(dev)
    # kept for the live comparison with the reference's own run further down (sampled at ITS selected nugget):
    # the posterior means [G, N, C] (3.9 GB), the sampled Q Q^T block and the metric curves
    ours_headline = None
    fit_check = None
    if world == 1 and info["classes"] > 1 and 1024 <= info["landmarks"] <= 20000:
        try:
            fit_check = fit_self_check(model, data, nugget, row_idx)
        except RuntimeError as exc:
            fit_check = {"error": str(exc)[:120]}
        torch.cuda.empty_cache()
    if world == 1:
        q_s = model.Q[sample_idx].double()
        ours_headline = {"fit": model.fit, "ksamp": (q_s @ q_s.T).cpu().numpy(), "curves": {k: v.numpy() for k, v in model.result.items()},
                         "best": int(np.argmax(model.result["val"].numpy()))}
        del q_s

    # ---- fraction-50 point of the landmark sweep (figure4.sh:2) + parity against the reference's golden fixture
    sweep = None
    parity_f50 = None
    if not args.no_sweep and args.shape == "reddit" and args.fraction != 50:
        land50 = synth.draw_landmarks(data.train_mask.cpu(), 50, seed=args.seed).to(dev)
        model.mask["landmark"] = land50
        for _ in range(2):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            s50 = step()
        e1.record()
        barrier()
        t50 = torch.tensor([e0.elapsed_time(e1) / 1e3 / reps], device=dev)
        if world > 1:
            dist.all_reduce(t50, op=dist.ReduceOp.MAX)
        sweep = [{"fraction": 50, "landmarks": int(land50.sum()), "value": round(float(t50), 5), "unit": UNIT, "steps": reps, "warmup": 2,
                  "best": {k: float(s50[k]) for k in ("train", "val", "test", "nugget")}}]
        golden = load_big_golden(GOLDEN_F50)
        if world == 1 and golden is not None and args.seed == 0:
            ours50 = collect_samples(model, sample_idx, row_idx, best=int(golden["best"]))
            parity_f50 = parity_record(ours50, golden, "golden tests/golden/big/%s.npz (unmodified reference, CPU fp32 run)" % GOLDEN_F50)
        model.mask["landmark"] = land
    elif args.shape == "reddit" and args.fraction == 50 and world == 1 and args.seed == 0:
        golden = load_big_golden(GOLDEN_F50)
        if golden is not None:
            ours50 = collect_samples(model, sample_idx, row_idx, best=int(golden["best"]))
            parity_f50 = parity_record(ours50, golden, "golden tests/golden/big/%s.npz (unmodified reference, CPU fp32 run)" % GOLDEN_F50)

    # ---- end to end from pinned host buffers through the public API (N = 1 path; sharded: each rank its copy)
    e2e = None
    if not args.no_e2e:
        src = {k: getattr(data, k) for k in ("x", "y", "edge_index", "edge_attr", "train_mask", "val_mask", "test_mask")}
        if world > 1:
            # a sharded job reads a sharded input: each rank's host buffers hold ITS rows of the adjacency
            # (partitioned once, like a dataset loader would, outside the timed region) plus features / labels
            from gnngp_b200.pipeline import RowLayout
            lay = RowLayout(info["nodes"], world, rank)
            mine = (src["edge_index"][0] >= lay.begin) & (src["edge_index"][0] < lay.end)
            src["edge_index"] = src["edge_index"][:, mine].contiguous()
            src["edge_attr"] = src["edge_attr"][mine].contiguous()
        host_bufs = {k: v.cpu().pin_memory() for k, v in src.items()}
        del src
        host_bufs["landmark"] = land.cpu().pin_memory()
        host_nugget = nugget.cpu().pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in host_bufs.values()) + host_nugget.numel() * 4
        from gnngp_b200.datasets import GraphData
        os.environ["GNNGP_STREAM_UPLOAD"] = args.stream_upload

        def e2e_step():
            ops.clear_cache()
            # the small inputs first: host->device copies execute in issue order
            land_dev = host_bufs["landmark"].to(dev, non_blocking=True)
            nugget_dev = host_nugget.to(dev, non_blocking=True)
            # the call a user of the reference makes: host data + a device; the model moves it (src/gnngp.py:69)
            g = GraphData(host_bufs["x"], host_bufs["y"], host_bufs["edge_index"], host_bufs["edge_attr"], host_bufs["train_mask"],
                          host_bufs["val_mask"], host_bufs["test_mask"])
            if world == 1:
                m = GNNGP(g, device=dev, Nystrom=True)
            else:
                from gnngp_b200.shard import ShardedGNNGP
                m = ShardedGNNGP(g, device=dev, Nystrom=True)
            m.mask["landmark"] = land_dev
            m.set_hyper_param(args.layers, args.sigma_b, args.sigma_w, method="GCN")
            m.predict(nugget_dev)
            s = m.get_summary()
            out = [float(s["train"]), float(s["val"]), float(s["test"]), float(s["nugget"])]
            d2h = sum(v.numel() * 4 for v in m.result.values())
            return out, d2h

        del model
        ops.clear_cache()
        torch.cuda.empty_cache()
        e2e_step()
        barrier()
        reps = max(1, min(args.steps, 2))
        t0 = time.perf_counter()
        for _ in range(reps):
            e2e_out, d2h = e2e_step()
        barrier()
        e2e_s = torch.tensor([(time.perf_counter() - t0) / reps], device=dev)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        moved = torch.tensor([float(h2d), float(d2h)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(moved)                       # whole-job bytes: sum over ranks
        e2e = {"value": round(float(e2e_s), 4), "unit": UNIT, "h2d_bytes_per_step": int(moved[0]), "d2h_bytes_per_step": int(moved[1]),
               "steps_run": reps, "stream_upload": args.stream_upload == "1",
               # the end-to-end run must reproduce the metrics of the HBM-resident run (same inputs, same landmarks)
               "matches_resident_run": bool(all(abs(a - best[k]) <= 2e-3 for a, k in zip(e2e_out[:3], ("train", "val", "test"))))}
        del host_bufs
    else:
        del model
    ops.clear_cache()
    ops.release_workspace()
    torch.cuda.empty_cache()

    # ---- the unmodified reference on this GPU, same inputs (single GPU only) + live parity of the timed run
    gpu_ref = None
    parity_live = None
    if world == 1 and not args.no_gpu_reference:
        try:
            from baseline import ref_arm
            gpu_ref, ref_samples = ref_arm.gpu_reference_full(data, land, nugget, args.layers, args.sigma_b, args.sigma_w,
                                                              sample_idx, row_idx, steps=1)
            if ref_samples is not None:
                # compare at the REFERENCE's selected nugget (validation accuracy is flat over the small nuggets, so
                # the two argmax-over-ties can land on different indices of the same plateau)
                gpu_ref["selected_nugget_index"] = {"ours": ours_headline["best"], "reference": ref_samples["best"]}
                ref_samples["fit_scale"] = float(np.abs(ref_samples["fit_best"]).max())
                ours_at = collect_samples(None, sample_idx, row_idx, best=ref_samples["best"], fit=ours_headline["fit"],
                                          ksamp=ours_headline["ksamp"], curves=ours_headline["curves"])
                parity_live = parity_record(ours_at, ref_samples, "live: unmodified reference (baseline/_ref) on this GPU, "
                                                                  "same inputs, fraction %d" % args.fraction)
                gpu_ref["speedup_value"] = round(gpu_ref["value"] / seconds, 2)
            ours_headline = None
        except Exception as exc:
            gpu_ref = {"unavailable": "gpu reference failed: %r" % (exc,)}
        torch.cuda.empty_cache()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu = cpu_arm(args, host, land.cpu())
            except Exception as exc:  # the CPU leg must never take the GPU number down with it
                cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % (exc,)}
        parity = {}
        if parity_live is not None:
            parity["fraction_%d" % args.fraction] = parity_live
        if parity_f50 is not None:
            parity["fraction_50"] = parity_f50
        headline_parity = parity_live or parity_f50
        config = {"workload": workload_name(args), "graph": info,
                  "l2": "inputs larger than L2 (adjacency %.2f GB, factor %.2f GB)" % (info["nnz"] * 8 / 1e9, info["nodes"] * (info["landmarks"] + 1) * 4 / 1e9),
                  "parallelism": "row-sharded x%d" % world if world > 1 else "single GPU",
                  "fit_solver": fit_solver, "fit_solver_stats": fit_stats, "fit_rel_vs_fp64_solve_on_own_factor": fit_check,
                  "root_solver": root_solver, "best": best, "setup_s": round(setup_s, 1),
                  "parity_vs_reference": ({k: headline_parity[k] for k in ("qqT_rel", "fit_rel", "argmax_flips", "ok", "source")}
                                          | {"cases": parity}) if headline_parity else None}
        line = {"metric": METRIC, "value": round(seconds, 5), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(seconds * 1e3, 3), "higher_is_better": False,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "clocks": clock_info, "e2e": e2e, "gpu_launches": int(launches_total), "gpu_launches_per_step": int(launches),
                "roofline": roof, "roofline_tensor": roof_tensor, "roofline_fp64": roof_fp64, "cpu_baseline": cpu, "gpu_reference": gpu_ref, "sweep": sweep,
                "stages_ms_per_step": {k: round(v["ms"] / args.steps, 3) for k, v in sorted(table.items(), key=lambda kv: -kv[1]["ms"])}}
        if args.breakdown:
            with open(args.breakdown, "w") as fh:
                json.dump({"table": table, "spmm": [[list(i), ms] for i, ms in spmm_calls], "line": line}, fh, indent=1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)
