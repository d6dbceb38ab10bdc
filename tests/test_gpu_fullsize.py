"""Parity at BASELINE.json's full sizes (B200 only), where no CPU oracle finishes in seconds:

* against a float64 torch pipeline computed ON THE GPU (tests/stub_ops.py in fp64 — plain torch ops, the
  same restatement the CPU suite pins to the reference's goldens) on the ogbn-arxiv shape, and
* through size-independent properties on the Reddit shape: the sparse product against cuSPARSE, its
  linearity and row sums; the Nystrom root reproducing the landmark block; linearity of the posterior
  mean in the targets; run-to-run agreement."""
import math

import pytest
import torch

import runner
from stub_ops import StubOps

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class Fp64GpuOps(StubOps):
    """The torch restatement of every operator, in float64 on the GPU (reference for big shapes)."""

    def empty(self, n, m, device=None):
        return torch.zeros((n, m), device=DEV, dtype=torch.float64)
    zeros = empty

    def csr_from_coo(self, row, col, val, n_rows, n_cols, row_offset=0):
        from stub_ops import _Csr
        mat = torch.sparse_coo_tensor(torch.stack([row, col]), val.double(), (n_rows, n_cols)).coalesce()
        return _Csr(mat, n_rows, n_cols, row_offset)

    def onehot_t(self, y, idx, c):
        return super().onehot_t(y, idx, c).double()

    def argmax_count(self, fitted, y, membership):
        return super().argmax_count(fitted.cpu(), y.cpu(), membership.cpu()).to(DEV)


@pytest.fixture(scope="module")
def ops():
    from gnngp_b200.ops import default_ops
    return default_ops()


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def test_arxiv_shape_gcngp_x_matches_fp64_pipeline(ops):
    from gnngp_b200 import synth
    data = synth.make_graph("arxiv", seed=0, device=DEV)
    land = synth.draw_landmarks(data.train_mask, 50, seed=0)
    nugget = synth.nugget_grid(DEV)
    masks = {"train": data.train_mask, "val": data.val_mask, "test": data.test_mask, "landmark": land}
    case = dict(fraction=50, initial="linear", params={}, L=2, sigma_b=0.0, sigma_w=1.0, method="GCN")
    d64 = type(data)(data.x.double(), data.y, data.edge_index, data.edge_attr.double(), data.train_mask, data.val_mask, data.test_mask)
    ref = runner.run_pipeline(Fp64GpuOps(), case, d64, masks, nugget.double(), device=DEV)
    out = runner.run_pipeline(ops, case, data, masks, nugget, device=DEV)
    idx = torch.randperm(data.num_nodes, device=DEV)[:2000]
    q, qr = out["Q"][idx].double(), ref["Q"][idx]
    assert rel(q @ q.T, qr @ qr.T) <= 1e-4
    best = int(torch.argmax(ref["result"]["val"]))
    for g in (best, 50, 75):
        assert rel(out["fit"][g], ref["fit"][g]) <= 1e-3, g
        agree = float((out["fit"][g].argmax(1) == ref["fit"][g].argmax(1)).float().mean())
        assert agree >= 0.9995, (g, agree)
    for key in ("train", "val", "test"):
        assert float((out["result"][key] - ref["result"][key]).abs().max()) <= 2e-3


@pytest.fixture(scope="module")
def reddit(ops):
    from gnngp_b200 import synth
    data = synth.make_graph("reddit", seed=0, device=DEV)
    n = data.num_nodes
    csr = ops.csr_from_coo(data.edge_index[0], data.edge_index[1], data.edge_attr, n, n)
    return data, csr


def test_reddit_shape_spmm_properties(ops, reddit):
    data, csr = reddit
    n = data.num_nodes
    assert csr.nnz == data.edge_index.shape[1] and int(csr.rowptr[-1]) == csr.nnz
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(n, 320, device=DEV, generator=g)
    y = torch.randn(n, 320, device=DEV, generator=g)
    ax, ay = ops.spmm(csr, x), ops.spmm(csr, y)
    # against cuSPARSE on the same CSR arrays
    a = torch.sparse_csr_tensor(csr.rowptr, csr.colidx.long(), csr.vals, (n, n))
    assert rel(ax, a @ x) <= 2e-6
    # linearity and scaling
    assert rel(ops.spmm(csr, x + y), ax + ay) <= 2e-6
    assert rel(ops.spmm(csr, x, alpha=0.37), 0.37 * ax) <= 1e-6
    # A 1 = row sums of the stored values
    ones = torch.ones(n, 64, device=DEV)
    rows = torch.repeat_interleave(torch.arange(n, device=DEV), csr.rowptr[1:] - csr.rowptr[:-1])
    want = torch.zeros(n, device=DEV, dtype=torch.float64).index_add_(0, rows, csr.vals.double())
    got = ops.spmm(csr, ones)
    assert rel(got[:, 0], want) <= 1e-6 and rel(got[:, 63], want) <= 1e-6
    # the panel path and the row-major path agree
    ops.spmm_tile_cols = 64
    try:
        assert rel(ops.spmm(csr, x), ax) <= 2e-6
    finally:
        ops.spmm_tile_cols = 0


def test_reddit_shape_nystrom_root_and_fit_properties(ops, reddit):
    from gnngp_b200 import compute, predict, synth
    data, csr = reddit
    n = data.num_nodes
    land = synth.draw_landmarks(data.train_mask, 100, seed=3)
    na = int(land.sum())
    g = torch.Generator(device=DEV).manual_seed(1)
    q = torch.randn(n, 200, device=DEV, generator=g) + 0.5
    root = compute._ExxT_ReLU_Nystrom(q, land)                    # [n, na]
    assert root.shape == (n, na)
    # landmark rows of the root reproduce the landmark block of g(QQ^T) (Nystrom is exact on landmarks)
    lq = q[land]
    s = ops.row_norms(lq)
    emm = ops.gram_arccos(lq, lq, s, s).double()
    rl = root[land].double()
    assert rel(rl @ rl.T, emm) <= 1e-4
    # and for non-landmark rows: R R_l^T = E[:, landmarks] wherever the spectrum is not clamped
    rows = torch.randperm(n, device=DEV)[:1024]
    e_rows = ops.gram_arccos(q[rows], lq, ops.row_norms(q[rows]), s).double()
    assert rel(root[rows].double() @ rl.T, e_rows) <= 2e-3
    # posterior mean is linear in float targets; two runs agree
    qf = root[:, :256].contiguous()
    y1 = torch.randn(n, device=DEV, generator=g)
    y2 = torch.randn(n, device=DEV, generator=g)
    nug = torch.logspace(-2, 1, 7, device=DEV)
    f1, f2 = predict.fit_Nystrom(qf, y1, data.train_mask, nug), predict.fit_Nystrom(qf, y2, data.train_mask, nug)
    f12 = predict.fit_Nystrom(qf, y1 + y2, data.train_mask, nug)
    assert f1.shape == (7, n)
    assert rel(f12, f1 + f2) <= 1e-4
    assert rel(predict.fit_Nystrom(qf, y1, data.train_mask, nug), f1) <= 1e-5
    res = predict.result(f1, y1, {"train": data.train_mask, "test": data.test_mask})
    r2 = 1 - ((f1[3][data.test_mask] - y1[data.test_mask]) ** 2).sum() / ((y1[data.test_mask] - y1[data.test_mask].mean()) ** 2).sum()
    assert abs(float(res["test"][3]) - float(r2)) <= 1e-4
