#!/usr/bin/env python
"""bench.py — GCNGP-X fit+predict seconds on a synthetic Reddit-shape graph (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--shape reddit] [--fraction 50] [--impl reference]

One "step" = what the reference times per run (``src/main.py:109-116``, landmark draw injected):
``set_hyper_param`` + ``predict`` (kernel recursion, fit for 101 nuggets, metrics) + ``get_summary``.

``value``   seconds per step with the graph resident in HBM (CSR built once, like the model object the
            reference keeps across runs), CUDA events, max over ranks, K steps after W warm-ups.
``e2e``     seconds per step through the public ``GNNGP`` API from pinned HOST buffers: host→device copy
            of features / labels / edges / masks, model construction (COO→CSR), predict, summary on host.
``roofline``the dominant kernel (layer-2 CSR SpMM): algorithmic bytes ÷ its CUDA-event time.
``cpu_baseline`` the oracle port of the reference algorithm on the host cores, on a bounded row sample.
``--impl reference`` prints the CPU arm alone (rank 0 only under torchrun).

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

import torch  # noqa: E402

METRIC = "gcngp_x_fit_predict_seconds"
UNIT = "s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="reddit")
    ap.add_argument("--fraction", type=int, default=50)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--sigma_b", type=float, default=0.0)
    ap.add_argument("--sigma_w", type=float, default=1.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--stream-upload", default="auto", choices=["auto", "0", "1"],
                    help="e2e leg: let the model upload the pinned edge list in pieces and start on them as they arrive "
                         "(GNNGP_STREAM_UPLOAD); auto = the measured best for this N")
    ap.add_argument("--cpu-sample-rows", type=int, default=0, help="adjacency rows in the CPU sample (0 = auto)")
    ap.add_argument("--breakdown", default="", help="write the per-operator timing table to this JSON file")
    return ap.parse_args()


# e2e leg: streamed upload per world size ("1" only where an A/B on the B200 box showed a gain)
# N = 1: 0.3200 s vs 0.3325 s (profiles/r1/call30_bench_reddit_f50_n1_*_upload.json); N > 1 not measured yet
STREAM_UPLOAD_AUTO = {1: "1"}


def workload_name(args) -> str:
    return "GCNGP-X (GCN, L=%d, fraction=%d, sigma_b=%g, sigma_w=%g, 101 nuggets) on synthetic %s-shape graph" % (
        args.layers, args.fraction, args.sigma_b, args.sigma_w, args.shape)


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
# CPU arm (oracle port on a bounded sample)
# --------------------------------------------------------------------------------------------
def cpu_sampled_baseline(args, data=None, land_mask=None) -> dict:
    """Time the reference algorithm's stages on host cores on a row sample and scale to full size."""
    import numpy as np
    from oracle import sampled
    from oracle import gnngp_oracle as orc
    from gnngp_b200 import synth
    spec = synth.SHAPES[args.shape]
    n, f, c = spec["n"], spec["f"], max(spec["c"], 1)
    n_train = spec["split"][0]
    if data is None:
        # no GPU data at hand (--impl reference): generate a graph of the same shape on the host.  For the
        # Reddit shape that is ~115 M edges; the sample only needs `rows` adjacency rows, so a graph with the
        # same degree law restricted to the sampled rows is drawn instead (same nnz per row in expectation).
        data, land_mask = host_sample_graph(args, spec)
    rows_wanted = args.cpu_sample_rows or max(256, min(n, int(n * 3.2e7 / max(1, data["nnz_total"]))))
    rng = np.random.default_rng(args.seed)
    pick = np.sort(rng.choice(n, size=min(rows_wanted, n), replace=False))
    indptr, indices, vals = data["rows_csr"](pick)
    sample = sampled.RowSample(indptr, indices, vals, n)
    x = data["x"]
    land_idx = np.nonzero(land_mask)[0]
    n_land = int(land_idx.shape[0])
    dense_rows = int(min(n, max(4096, min(65536, n // 4))))
    threads = orc._native().oracle_num_threads() if orc._native() is not None else 1
    est = sampled.estimate_gcn_nystrom_seconds(sample, n, x, x[land_idx] if n_land >= f else np.zeros((n_land, f), np.float32),
                                               n_land, n_train, c, 101, args.layers, args.sigma_b, args.sigma_w,
                                               dense_rows, seed=args.seed)
    return {"value": est["total_estimated_s"], "unit": UNIT, "cores": int(max(threads, os.cpu_count() or 1)),
            "kind": "port",
            "sample": "%d of %d adjacency rows for the two sparse products, %d rows for the dense row-separable stages, "
                      "both eigendecompositions (n=%d, %d) at full size; stage times scaled by the inverse sampling rate; "
                      "measured %.1f s of CPU work" % (sample.n_rows, n, dense_rows, n_land, n_land + 1, est["sample_measured_s"]),
            "stages_s": {k: round(v, 4) for k, v in est.items()}}


def host_sample_graph(args, spec):
    """Host-only stand-in used by --impl reference: features and a landmark mask of the right shape, and
    adjacency rows drawn with the generator's degree law for the sampled nodes only."""
    import numpy as np
    n, f = spec["n"], spec["f"]
    rng = np.random.default_rng(args.seed + 1)
    x = rng.standard_normal((n, f), dtype=np.float32)
    train = np.zeros(n, dtype=bool)
    train[rng.choice(n, spec["split"][0], replace=False)] = True
    land = train & (rng.random(n) < 1.0 / args.fraction)
    nnz_total = 2 * spec["und"] + n
    weights = np.arange(1, n + 1, dtype=np.float64) ** -0.7
    perm = rng.permutation(n)
    share = np.empty(n)
    share[perm] = weights / weights.sum()           # expected share of edge endpoints per node

    def rows_csr(pick):
        deg = np.maximum(1, rng.poisson(share[pick] * 2 * spec["und"]) + 1)
        indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
        cdf = np.cumsum(share)
        indices = np.searchsorted(cdf, rng.random(int(indptr[-1]))).clip(max=n - 1).astype(np.int32)
        for i in range(len(pick)):                   # column order inside a row, as in a coalesced operand
            indices[indptr[i]:indptr[i + 1]].sort()
        vals = (rng.random(int(indptr[-1]), dtype=np.float32) + 0.5) / np.float32(500.0)
        return indptr, indices, vals

    return {"x": x, "rows_csr": rows_csr, "nnz_total": nnz_total}, land


def device_graph_for_cpu(model_data, csr):
    """Adapter: hand the CPU arm the SAME graph the GPU arm ran on (rows sampled from the device CSR)."""
    import numpy as np
    rowptr = csr.rowptr.cpu().numpy()
    x = model_data.x.detach().cpu().numpy()

    def rows_csr(pick):
        starts, ends = rowptr[pick], rowptr[pick + 1]
        deg = ends - starts
        indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
        pos = torch.from_numpy(np.concatenate([np.arange(s, e) for s, e in zip(starts, ends)]).astype(np.int64)).to(csr.colidx.device)
        return indptr, csr.colidx[pos].cpu().numpy(), csr.vals[pos].cpu().numpy()

    return {"x": x, "rows_csr": rows_csr, "nnz_total": int(csr.nnz)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    samples = []
    base = None
    for i in range(max(1, min(args.steps, 2))):
        args.seed = i
        base = cpu_sampled_baseline(args)
        samples.append(base["value"])
    value = sum(samples) / len(samples)
    base["value"] = value
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": value * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "device": "cpu", "wall_s": round(time.perf_counter() - t0, 1)},
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
            rec = json.load(fh)[kernel]
        if int(rec["nnz"]) == int(info["nnz"]) and int(rec["n"]) == int(info["nodes"]) and int(rec["k"]) == int(k):
            return float(rec["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def run_ours(args):
    import torch.distributed as dist
    from gnngp_b200 import synth
    from gnngp_b200.gnngp import GNNGP
    from gnngp_b200.ops import default_ops
    from gnngp_b200.profiling import OpTimer

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

    # ---- synthetic inputs, generated on the device (identical on every rank: same seed, same generator)
    data = synth.make_graph(args.shape, seed=args.seed, device=dev, normalization="sym")
    land = synth.draw_landmarks(data.train_mask, args.fraction, seed=args.seed)
    nugget = synth.nugget_grid(dev)
    info = synth.describe(data)
    info["landmarks"] = int(land.sum())

    if world == 1:
        model = GNNGP(data, device=dev, Nystrom=True)
    else:
        from gnngp_b200.shard import ShardedGNNGP
        model = ShardedGNNGP(data, device=dev, Nystrom=True)
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
    fit_solver = getattr(getattr(model, "pipe", None), "last_fit_solver", None) or "eigh"

    # ---- roofline of the dominant kernel: the layer-2 SpMM (largest k)
    peaks = {"hbm_gbs": 6650.0, "source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh); peaks["source"] = "measured"
    except Exception:
        pass
    roof = None
    if spmm_calls:
        work = max(i[1] * i[2] for i, _ in spmm_calls)            # the launch with the most nnz x k: layer-2 A @ E
        big = [(i, ms) for i, ms in spmm_calls if i[1] * i[2] == work]
        (rows, nnz, k), _ = big[0]
        ms = sum(m for _, m in big) / len(big)
        n_cols = info["nodes"]
        alg_bytes = 8.0 * nnz + 4.0 * (rows + 1) + 4.0 * n_cols * k + 4.0 * rows * k
        achieved = alg_bytes / (ms * 1e-3) / 1e9
        gather_bytes = 4.0 * nnz * k
        roof = {"bound": "hbm", "kernel": "spmm_panel_kernel (layer-2 propagation A @ E, k=%d)" % k, "achieved": round(achieved, 2),
                "peak": peaks["hbm_gbs"], "peak_source": peaks["source"], "unit": "GB/s", "frac": round(achieved / peaks["hbm_gbs"], 5),
                "traffic": ncu_traffic("spmm_panel_kernel", info, k), "ms_per_launch": round(ms, 3), "algorithmic_bytes": alg_bytes,
                "l2_gather_bytes": gather_bytes, "l2_gather_tbs": round(gather_bytes / (ms * 1e-3) / 1e12, 3),
                "fp32_tflops": round(2.0 * nnz * k / (ms * 1e-3) / 1e12, 2)}

    # ---- tensor-core side: the largest tcgen05 launch of the step (the all-nugget posterior sweep at the default
    # workload), plus the fused Gram + arc-cosine contraction; op times include the operand split pre-pass
    roof_tensor = None
    tf32_peak = peaks.get("bf16_tflops", 1590.0) / 2.0

    def tensor_entry(op_name, kernel_name):
        calls = [(i, ms) for i, ms in timer.calls(op_name) if i is not None]
        if not calls:
            return None
        work = max(i[0] * i[1] * i[2] for i, _ in calls)
        big = [(i, ms) for i, ms in calls if i[0] * i[1] * i[2] == work]
        (gm, gn, gk), _ = big[0]
        gms = sum(m for _, m in big) / len(big)
        useful = 2.0 * gm * gn * gk / (gms * 1e-3) / 1e12
        return {"bound": "tensor", "kernel": "%s (%d x %d x %d, 3xTF32)" % (kernel_name, gm, gn, gk),
                "achieved": round(useful, 1), "issued_tflops": round(3 * useful, 1), "peak": round(tf32_peak, 1),
                "peak_source": "%s bf16 burst / 2 (TF32)" % peaks["source"], "unit": "TFLOP/s",
                "frac": round(3 * useful / tf32_peak, 4), "ms_per_launch": round(gms, 3), "flops": 2.0 * gm * gn * gk,
                "note": "useful flops x 3 MMAs per product (fp32-equivalent split) against the TF32 rate"}

    entries = [e for e in (tensor_entry("nugget_sweep", "gemm_nt_tc3_kernel<EpiSweep>"),
                           tensor_entry("gemm_nt", "gemm_nt_tc3_kernel<EpiPlain>"),
                           tensor_entry("gram_arccos", "gemm_nt_tc3_kernel<EpiArccos>")) if e is not None]
    if entries:
        entries.sort(key=lambda e: -e["flops"])
        roof_tensor = entries[0]
        roof_tensor["others"] = [{k: e[k] for k in ("kernel", "achieved", "frac", "ms_per_launch")} for e in entries[1:]]

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
        host = {k: v.cpu().pin_memory() for k, v in src.items()}
        del src
        host["landmark"] = land.cpu().pin_memory()
        host_nugget = nugget.cpu().pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in host.values()) + host_nugget.numel() * 4
        from gnngp_b200.datasets import GraphData

        stream_upload = STREAM_UPLOAD_AUTO.get(world, "0") if args.stream_upload == "auto" else args.stream_upload
        os.environ["GNNGP_STREAM_UPLOAD"] = stream_upload

        def e2e_step():
            ops.clear_cache()
            # the small inputs first: host->device copies execute in issue order, so a copy issued after the edge
            # list would wait for all of it (and hold back the stream that waits for that copy)
            land_dev = host["landmark"].to(dev, non_blocking=True)
            nugget_dev = host_nugget.to(dev, non_blocking=True)
            # the call a user of the reference makes: host data + a device; the model moves it (src/gnngp.py:69)
            g = GraphData(host["x"], host["y"], host["edge_index"], host["edge_attr"], host["train_mask"], host["val_mask"],
                          host["test_mask"])
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
               "stream_upload": stream_upload == "1",
               # the end-to-end run must reproduce the metrics of the HBM-resident run (same inputs, same landmarks)
               "matches_resident_run": bool(all(abs(a - float(summary[k])) <= 2e-3
                                                for a, k in zip(e2e_out[:3], ("train", "val", "test"))))}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                csr = ops.csr_of(torch.sparse_coo_tensor(data.edge_index, data.edge_attr, (info["nodes"],) * 2))
                cpu = cpu_sampled_baseline(args, device_graph_for_cpu(data, csr), land.cpu().numpy())
            except Exception as exc:  # the CPU leg must never take the GPU number down with it
                cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (exc,)}
        line = {"metric": METRIC, "value": round(seconds, 5), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(seconds * 1e3, 3), "higher_is_better": False,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(args), "graph": info, "l2": "inputs larger than L2 (adjacency %.2f GB, factor %.2f GB)" % (
                    info["nnz"] * 8 / 1e9, info["nodes"] * (info["landmarks"] + 1) * 4 / 1e9),
                    "parallelism": "row-sharded x%d" % world if world > 1 else "single GPU",
                    "fit_solver": fit_solver,
                    "best": {k: float(summary[k]) for k in ("train", "val", "test", "nugget")}},
                "clocks": clock_info, "e2e": e2e, "gpu_launches": int(launches_total), "gpu_launches_per_step": int(launches), "roofline": roof, "roofline_tensor": roof_tensor,
                "cpu_baseline": cpu,
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
