"""GPU-box probe: the SpMM of the layer-2 propagation under every tuning-flag combination (``gnngp_spmm_set_flags``)
and both gather layouts, at the arxiv shape (short rows: the HBM-bound regime of SURVEY §8d) and the Reddit shape
(dense rows: the L2-gather-bound regime).  Every variant is checked against the plain kernels' result (flags 0) before it is
timed.  (The first run of this probe, profiles/r2/call28_*, still had an L2 evict_last flag on the gathers — flag 2 there, no
effect at either shape, removed since — and numbered the zero-fill flag 4; the third, call30_*, had a flag 8 that gave the
short-row kernel eight gathers in flight instead of four: slower, removed.)

    python tools/spmm_policy_probe.py [out.json] [--shapes arxiv:9200,reddit:3052,reddit:15404]

Writes one JSON record (all variants, CUDA-event times of the whole ``ops.spmm`` call: zero-fill + panel copy + kernel) and
``<out>.best`` = the fastest flag set at the LAST shape of the list (the batch script hands it to bench.py).
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnngp_b200 import synth            # noqa: E402
from gnngp_b200.ops import default_ops  # noqa: E402


def timed(fn, reps):
    times = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    return times


def rel(a, b):
    num = torch.zeros((), dtype=torch.float64, device=a.device)
    den = torch.zeros((), dtype=torch.float64, device=a.device)
    for r0 in range(0, a.shape[0], 16384):       # blockwise: no second full-size temporary
        x, y = a[r0:r0 + 16384].double(), b[r0:r0 + 16384].double()
        num += (x - y).pow(2).sum()
        den += y.pow(2).sum()
    return float((num / den.clamp_min(1e-300)).sqrt())


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else os.path.join(ROOT, "gpurun_out", "spmm_policy_probe.json")
    shapes = "arxiv:9200,reddit:3052,reddit:15404"
    if "--shapes" in sys.argv:
        shapes = sys.argv[sys.argv.index("--shapes") + 1]
    dev = torch.device("cuda:0")
    ops = default_ops()
    record = {"device": torch.cuda.get_device_name(0), "cases": []}
    graphs = {}
    best_last = 0
    for item in shapes.split(","):
        shape, k = item.split(":")
        k = int(k)
        if shape not in graphs:
            graphs.clear()
            torch.cuda.empty_cache()
            data = synth.make_graph(shape, seed=0, device=dev)
            n = data.num_nodes
            graphs[shape] = (ops.csr_from_coo(data.edge_index[0], data.edge_index[1], data.edge_attr, n, n), n)
            del data
        csr, n = graphs[shape]
        b = ops.empty(n, k, dev)
        b.normal_()
        c = ops.empty(n, k + 1, dev)[:, :k]         # the pipeline's output view (bias column behind it)
        want = ops.empty(n, k + 1, dev)[:, :k]
        library_default = ops.set_spmm_flags(0)
        ops.spmm_panels_mode = "auto"
        ops.spmm(csr, b, 1.0, out=want)
        auto_kernel = ops.spmm_kernel_by_k[k]
        reps = 2 if csr.nnz * k > 1e12 else 4
        alg_bytes = 8.0 * csr.nnz + 4.0 * (n + 1) + 8.0 * n * k
        case = {"shape": shape, "n": n, "nnz": csr.nnz, "k": k, "default_kernel": auto_kernel, "algorithmic_bytes": alg_bytes,
                "variants": []}
        layouts = ["auto"] if csr.nnz >= 64 * n else ["0", "1"]           # dense rows: the row-major gather is known slower
        tiles = [0] if csr.nnz >= 64 * n else [0, 32, 128]
        for layout in layouts:
            for tile in tiles:
                if tile and layout == "1":
                    continue
                for flags in range(8):
                    if (tile or layout == "1") and flags not in (0, 3):
                        continue
                    ops.spmm_panels_mode, ops.spmm_tile_cols = layout, tile
                    ops.set_spmm_flags(flags)
                    c.fill_(float("nan"))
                    ops.spmm(csr, b, 1.0, out=c)
                    err = rel(c, want)
                    ok = bool(err < 1e-6) and not bool(torch.isnan(c).any())
                    ms = timed(lambda: ops.spmm(csr, b, 1.0, out=c), reps)
                    best = min(ms)
                    case["variants"].append({"layout": layout, "tile": tile, "flags": flags, "kernel": ops.spmm_kernel_by_k[k],
                                             "ms": [round(t, 3) for t in ms], "ms_best": round(best, 3), "rel_vs_default": err,
                                             "ok": ok, "algorithmic_gbs": round(alg_bytes / best / 1e6, 1),
                                             "gather_tbs": round(4.0 * csr.nnz * k / best / 1e9, 2)})
                    print(json.dumps(case["variants"][-1]), flush=True)
        ops.spmm_panels_mode, ops.spmm_tile_cols = "auto", 0
        ops.set_spmm_flags(library_default)
        case["library_default_flags"] = library_default
        good = [v for v in case["variants"] if v["ok"]]
        base = [v for v in good if v["flags"] == 0 and v["tile"] == 0 and v["layout"] == layouts[0]][0]
        top = min(good, key=lambda v: v["ms_best"])
        case["default_ms"], case["best"] = base["ms_best"], top
        same_layout = [v for v in good if v["tile"] == 0 and v["layout"] == layouts[0]]
        best_last = min(same_layout, key=lambda v: v["ms_best"])["flags"]
        case["best_flags_default_layout"] = best_last
        record["cases"].append(case)
        print("%s k=%d: default %.2f ms, best %.2f ms (%s)" % (shape, k, base["ms_best"], top["ms_best"], json.dumps(top)), flush=True)
        del b, c, want
        torch.cuda.empty_cache()
    record["all_ok"] = all(v["ok"] for case in record["cases"] for v in case["variants"])
    with open(out_path, "w") as fh:
        json.dump(record, fh, indent=1)
    with open(out_path + ".best", "w") as fh:
        fh.write("%d\n" % best_last)
    print("all variants agree with the plain kernels: %s; best flags at the last shape: %d" % (record["all_ok"], best_last))


if __name__ == "__main__":
    main()
