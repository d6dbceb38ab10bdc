"""Run a parity case through the product's host pipeline with a given operator set, and check the
outputs against the golden fixture (shared by the CPU host-logic tests and the GPU parity tests)."""
from __future__ import annotations

import numpy as np
import torch

import cases
import parity
from gnngp_b200.pipeline import Landmarks, Pipeline, RowLayout


def run_pipeline(ops, case, data, masks, nugget, device="cpu"):
    """Single-rank path: initial factor/kernel → recursion → fit → metrics (what GNNGP.predict does)."""
    pipe = Pipeline(ops)
    x = data.x.to(device)
    y = data.y.to(device)
    masks = {k: v.to(device) for k, v in masks.items()}
    nugget = nugget.to(device)
    adj = torch.sparse_coo_tensor(data.edge_index.to(device), data.edge_attr.to(device), (data.num_nodes,) * 2)
    csr = ops.csr_of(adj)
    rows = torch.nonzero(masks["train"], as_tuple=False).flatten()
    classes = None if y.dtype.is_floating_point else int(y[rows].max()) + 1
    out = {}
    if case["fraction"] is not None:
        land = Landmarks(masks["landmark"], RowLayout(data.num_nodes))
        q0 = pipe.initial_factor(x, land, case["initial"], case["params"])
        q = pipe.kernel_nystrom(q0, csr, case["L"], case["sigma_b"], case["sigma_w"], land, case["method"], case["params"])
        out["Q"] = q
        fit = pipe.fit_nystrom(q, y, rows, nugget, None, classes)
    else:
        c0 = pipe.initial_kernel(x, case["initial"], case["params"])
        k = pipe.kernel_exact(c0, csr, case["L"], case["sigma_b"], case["sigma_w"], case["method"], case["params"])
        out["K"] = k
        fit = pipe.fit_exact(k, y, rows, nugget, classes)
    out["fit"] = fit
    out["result"] = pipe.metrics(fit, y, masks)
    return out


def kernel_block(out, idx):
    tidx = torch.as_tensor(idx, device=(out.get("Q", out.get("K"))).device)
    if "Q" in out:
        q = out["Q"][tidx].double()
        return (q @ q.T).cpu().numpy()
    return out["K"][tidx][:, tidx].cpu().numpy()


def check_against_golden(name, out, masks, report=None):
    golden = parity.load_golden(name)
    assert golden is not None, "fixture %s.npz missing" % name
    e32, e64 = parity.check_kernel_block(golden, kernel_block(out, golden["sample_index"]))
    best = int(golden["best"])
    fit = out["fit"]
    fits = parity.check_fit(golden, fit[best].cpu().numpy(), fit[50].cpu().numpy())
    agree = parity.check_predictions(golden, fit[best].cpu().numpy())
    sizes = {k: int(masks[k].sum()) for k in ("train", "val", "test")}
    parity.check_curves(golden, {k: v.numpy() for k, v in out["result"].items()}, sizes)
    if report is not None:
        report[name] = dict(k_rel_fp32=e32, k_rel_fp64=e64, fit_best=fits["fit_best"], fit_mid=fits["fit_mid"], argmax_agree=agree)
    return e32, e64
