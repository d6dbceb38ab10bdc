"""Host logic of the Krylov route of the Nystrom root (gnngp_b200/krylov_root.py) through the CPU stand-in of the fp64
operators: the factor it returns must reproduce the reference's ``Q Qᵀ`` (src/compute.py:6-17) — the factor itself
differs by an orthogonal matrix — and indefinite / unconverged blocks must be handed back to the eigen route."""
import numpy as np
import pytest
import torch

from gnngp_b200 import krylov_root
from gnngp_b200.pipeline import Pipeline
from stub_ops import StubOps


def _block(n, above, seed=0, floor=8e-6, gap=0.5):
    """SPD matrix with ``above`` eigenvalues spread over four decades above the clamp and the rest packed below it
    (the shape of the benchmark's landmark block: 647 of 15 404 above 1e-4 * max)."""
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.concatenate([np.exp(rng.uniform(np.log(floor), np.log(gap * 1e-4), n - above)),
                          np.exp(rng.uniform(np.log(1.2e-4), 0.0, above - 1)), [1.0]])
    m = (q * lam) @ q.T
    return torch.from_numpy((m + m.T) / 2).float()


def _reference_root(k_cols, kmm):
    """src/compute.py:11-16 in fp64 (rows of ``k_cols`` are non-landmark rows)."""
    d, v = torch.linalg.eigh(kmm.double())
    d_sqrt = torch.sqrt(torch.clamp(d, 0))
    inv = d_sqrt / torch.clamp(d, d[-1] * 1e-4)
    return v * d_sqrt.view(1, -1), k_cols.double() @ v * inv.view(1, -1)


@pytest.mark.parametrize("n,above", [(900, 40), (1300, 90)])
def test_krylov_root_reproduces_the_reference_factor_products(n, above, monkeypatch):
    monkeypatch.setenv("GNNGP_ROOT_SOLVER", "krylov")
    kmm = _block(n, above)
    gen = torch.Generator().manual_seed(3)
    k_cols = (torch.randn(64, n, generator=gen).double() @ kmm.double()).float()     # rows "like" landmark rows
    stats = {}
    got = krylov_root.root_factors(StubOps(), kmm, stats, block=32)
    assert got is not None and stats["converged"], stats
    w_t, land = got
    assert stats["ritz_above_clamp"][-1] == above
    ref_land, ref_rows = _reference_root(k_cols, kmm)
    rows = k_cols.double() @ w_t.double().T
    for mine, ref in ((land.double() @ land.double().T, ref_land @ ref_land.T),
                      (rows @ rows.T, ref_rows @ ref_rows.T),
                      (rows @ land.double().T, ref_rows @ ref_land.T)):
        assert float(torch.linalg.norm(mine - ref) / torch.linalg.norm(ref)) < 2e-6


def test_krylov_root_hands_back_indefinite_and_unconverged_blocks(monkeypatch):
    monkeypatch.setenv("GNNGP_ROOT_SOLVER", "krylov")
    kmm = _block(900, 40)
    bad = kmm.clone()
    bad[5, 5] = -1.0                                                   # indefinite: the Cholesky pivot check fires
    assert krylov_root.root_factors(StubOps(), bad, block=32) is None
    dense = _block(900, 700, gap=0.9)                                  # most of the spectrum above the clamp
    stats = {}
    assert krylov_root.root_factors(StubOps(), dense, stats, block=32) is None and not stats["converged"]


def test_pipeline_falls_back_to_eigh(monkeypatch):
    monkeypatch.setenv("GNNGP_ROOT_SOLVER", "krylov")
    monkeypatch.setenv("GNNGP_ROOT_BLOCK", "32")
    pipe = Pipeline(StubOps())
    dense = _block(900, 700, gap=0.9)
    w_t, v_root = pipe._root_factors(dense)
    assert pipe.last_root_solver == "eigh" and w_t.shape == (900, 900)
    w_t, v_root = pipe._root_factors(_block(900, 40))
    assert pipe.last_root_solver == "krylov"


def test_route_policy(monkeypatch):
    ops = StubOps()
    monkeypatch.delenv("GNNGP_ROOT_SOLVER", raising=False)
    assert not krylov_root.wanted(3052, ops) and krylov_root.wanted(15404, ops)
    monkeypatch.setenv("GNNGP_ROOT_SOLVER", "eigh")
    assert not krylov_root.wanted(15404, ops)
