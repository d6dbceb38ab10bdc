"""Host logic of the block-Lanczos posterior solve (gnngp_b200/krylov_fit.py) through the CPU stand-in of the fp64
operators: the coefficient blocks of every nugget against dense fp64 solves, and the hand-back when the step budget is
not enough."""
import numpy as np
import torch

from gnngp_b200 import krylov_fit
from stub_ops import StubOps


def _gram(m, outliers, seed=0, bulk=(2e-8, 2e-7)):
    """PSD matrix with the spectrum shape of Q_bᵀQ_b: a few hundred significant eigenvalues and a bulk far below."""
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((m, m)))
    lam = np.concatenate([rng.uniform(bulk[0], bulk[1], m - outliers), np.exp(rng.uniform(np.log(1e-5), 0.0, outliers))])
    g = (q * lam) @ q.T
    return torch.from_numpy((g + g.T) / 2 * 5e4).float()


def test_krylov_fit_matches_dense_solves(monkeypatch):
    m, c = 700, 6
    gram = _gram(m, 30)
    gen = torch.Generator().manual_seed(1)
    y = torch.randn(2000, c, generator=gen)
    qb = torch.randn(2000, m, generator=gen) * 0.01
    qty_t = ((gram.double() @ torch.randn(m, c, generator=gen).double()).float()).T.contiguous()    # in the range of G
    d = float(torch.diagonal(gram).mean())
    shifts = torch.logspace(-3, 1, 11, dtype=torch.float64) * d
    stats = {}
    got = krylov_fit.coefficients(StubOps(), gram, qty_t, shifts, stats)
    assert got is not None and stats["converged"], stats
    y_t, basis_t = got
    got = y_t.double() @ basis_t.double()                              # [(G*C) x m]
    sym = torch.tril(gram).double()
    sym = sym + torch.tril(sym, -1).T
    for g in range(shifts.numel()):
        want = torch.linalg.solve(sym + shifts[g] * torch.eye(m, dtype=torch.float64), qty_t.double().T).T
        block = got[g * c:(g + 1) * c].double()
        assert float((block - want).norm() / want.norm()) < 1e-5, (g, stats)


def test_krylov_fit_gives_up_on_a_flat_spectrum(monkeypatch):
    m, c = 600, 5
    rng = np.random.default_rng(3)
    q, _ = np.linalg.qr(rng.standard_normal((m, m)))
    lam = np.exp(rng.uniform(np.log(1e-6), 0.0, m))                  # no bulk: every eigenvalue matters at the small shift
    gram = torch.from_numpy((q * lam) @ q.T).float()
    qty_t = torch.randn(c, m)
    shifts = torch.tensor([1e-9, 1e-3], dtype=torch.float64)
    stats = {}
    assert krylov_fit.coefficients(StubOps(), gram, qty_t, shifts, stats) is None and not stats["converged"]


def test_policy(monkeypatch):
    ops = StubOps()
    monkeypatch.delenv("GNNGP_FIT_SOLVER", raising=False)
    assert krylov_fit.wanted(15405, 41, ops) and not krylov_fit.wanted(1500, 41, ops) and not krylov_fit.wanted(15405, 1, ops)
    monkeypatch.setenv("GNNGP_FIT_SOLVER", "band")
    assert not krylov_fit.wanted(15405, 41, ops)
