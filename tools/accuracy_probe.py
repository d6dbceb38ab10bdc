"""GPU-box probe (not a test): where does the CUDA path lose accuracy against the reference?

Runs parity cases with (1) the CUDA operator set, fp32 CUDA-core contractions, (2) the same with the
tcgen05 3xTF32 engine, (3) either of them with eigh moved to the host LAPACK, (4) plain torch ops on the
GPU (tests/stub_ops.py on CUDA tensors), and on larger synthetic shapes against a float64 torch pipeline
computed on the GPU.  Prints relative errors of sampled K / QQ^T blocks and of the posterior means."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases      # noqa: E402
import parity     # noqa: E402
import runner     # noqa: E402
from stub_ops import StubOps            # noqa: E402
from gnngp_b200 import synth            # noqa: E402
from gnngp_b200.ops import default_ops  # noqa: E402

dev = torch.device("cuda:0")
ops = default_ops()
report = {}


class HostEigh(object):
    """ops proxy: everything on the GPU except the eigensolver (LAPACK ssyevd on the host)."""
    def __init__(self, inner, dtype=torch.float32, on_gpu=False):
        self.inner, self.dtype, self.on_gpu = inner, dtype, on_gpu

    def __getattr__(self, name):
        return getattr(self.inner, name)

    def eigh(self, m):
        src = m.contiguous().to(self.dtype)
        if not self.on_gpu:
            src = src.cpu()
        w, v = torch.linalg.eigh(src)
        return w.to(device=m.device, dtype=torch.float32), v.to(device=m.device, dtype=torch.float32).contiguous()


class CudaStub(StubOps):
    def empty(self, n, m, device=None):
        return torch.zeros((n, (max(m, 1) + 3) // 4 * 4), device=dev)[:, :m]
    zeros = empty

    def argmax_count(self, fitted, y, membership):
        return super().argmax_count(fitted.cpu(), y.cpu(), membership.cpu()).to(dev)

    def sqerr_sum(self, fitted, y, membership):
        return super().sqerr_sum(fitted.cpu(), y.cpu(), membership.cpu()).to(dev)


def golden_cases():
    names = ["tinyreg-GIN-nys1-L3", "tinyreg-GCN2-nys1-L3", "tiny-GIN-nys2-L3", "small-GCN-nys4-L2", "pubmed-GCN-nys4-L2",
             "chameleon-GIN-nys1-L2", "chameleon-SAGE-exact-L10", "cora-GCN-exact-L2"]
    for name in names:
        case = cases.CASES[name]
        data, masks, nugget = cases.build_inputs(case)
        g = parity.load_golden(name)
        row = {"ref32_vs_64": parity.rel_fro(g["ksamp"], g["ksamp64"]), "fit_floor": parity.rel_fro(g["fit_best"], g["fit_best64"])}
        variants = {"cuda_simt": (ops, 0), "cuda_tc": (ops, 1), "cuda_simt_host_eigh": (HostEigh(ops), 0),
                    "cuda_simt_gpu_eigh64": (HostEigh(ops, torch.float64, True), 0),
                    "cuda_tc_gpu_eigh64": (HostEigh(ops, torch.float64, True), 1), "torch_gpu": (CudaStub(), 0)}
        for label, (o, engine) in variants.items():
            ops.set_gemm_engine(engine)
            try:
                out = runner.run_pipeline(o, case, data, masks, nugget, device=dev)
                blk = runner.kernel_block(out, g["sample_index"])
                best = int(g["best"])
                row[label] = {"k64": parity.rel_fro(blk, g["ksamp64"]),
                              "fit_best64": parity.rel_fro(out["fit"][best].cpu().numpy(), g["fit_best64"]),
                              "fit_mid64": parity.rel_fro(out["fit"][50].cpu().numpy(), g["fit_mid64"])}
            except Exception as exc:
                row[label] = "failed: %r" % (exc,)
        ops.set_gemm_engine(2)
        report[name] = row
        print(name, json.dumps(row), flush=True)


def big_shape(shape, fraction, rows=1500):
    """fp64 torch pipeline on the GPU as ground truth for a big synthetic shape."""
    data = synth.make_graph(shape, seed=0, device=dev)
    land = synth.draw_landmarks(data.train_mask, fraction, seed=0)
    nugget = synth.nugget_grid(dev)
    masks = {"train": data.train_mask, "val": data.val_mask, "test": data.test_mask, "landmark": land}
    case = dict(fraction=fraction, initial="linear", params={}, L=2, sigma_b=0.0, sigma_w=1.0, method="GCN")
    idx = torch.randperm(data.num_nodes, device=dev)[:rows].sort().values

    class Stub64(CudaStub):
        def empty(self, n, m, device=None):
            return torch.zeros((n, m), device=dev, dtype=torch.float64)
        zeros = empty

        def csr_from_coo(self, row, col, val, n_rows, n_cols, row_offset=0):
            return super().csr_from_coo(row, col, val.double(), n_rows, n_cols, row_offset)

        def onehot_t(self, y, idx_, c):
            return super().onehot_t(y, idx_, c).double()

    d64 = type(data)(data.x.double(), data.y, data.edge_index, data.edge_attr.double(), data.train_mask, data.val_mask, data.test_mask)

    # the stub casts sparse values to float; build the fp64 operator by hand
    class Csr64(object):
        pass
    s64 = Stub64()
    s64.csr_from_coo = lambda row, col, val, n_rows, n_cols, row_offset=0: type("C", (), dict(
        mat=torch.sparse_coo_tensor(torch.stack([row, col]), val.double(), (n_rows, n_cols)).coalesce(), n_rows=n_rows, n_cols=n_cols,
        row_offset=row_offset, nnz=row.numel()))()
    ref = runner.run_pipeline(s64, case, d64, masks, nugget.double(), device=dev)
    qr = ref["Q"][idx]
    kr = (qr @ qr.T).cpu().numpy()
    out_row = {"landmarks": int(land.sum())}
    for label, engine in (("cuda_simt", 0), ("cuda_tc", 1)):
        ops.set_gemm_engine(engine)
        out = runner.run_pipeline(ops, case, data, masks, nugget, device=dev)
        q = out["Q"][idx].double()
        kk = (q @ q.T).cpu().numpy()
        fit, fr = out["fit"], ref["fit"]
        best = int(torch.argmax(ref["result"]["val"]))
        agree = float((fit[best].argmax(1) == fr[best].argmax(1)).float().mean())
        agree_mid = float((fit[50].argmax(1) == fr[50].argmax(1)).float().mean())
        out_row[label] = {"k64": parity.rel_fro(kk, kr), "best_nugget_index": best,
                          "fit_best64": parity.rel_fro(fit[best].cpu().numpy(), fr[best].cpu().numpy()),
                          "fit_mid64": parity.rel_fro(fit[50].cpu().numpy(), fr[50].cpu().numpy()),
                          "argmax_agree_best": agree, "argmax_agree_mid": agree_mid,
                          "val_curve_maxdiff": float((out["result"]["val"] - ref["result"]["val"]).abs().max())}
    ops.set_gemm_engine(2)
    report["%s-f%d" % (shape, fraction)] = out_row
    print(shape, fraction, json.dumps(out_row), flush=True)
    del ref, out
    torch.cuda.empty_cache()


if __name__ == "__main__":
    golden_cases()
    big_shape("arxiv", 50)
    if "--reddit" in sys.argv:
        big_shape("reddit", 50)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "accuracy_probe.json"), "w") as fh:
        json.dump(report, fh, indent=1)
