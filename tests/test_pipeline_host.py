"""Host logic of the product (gnngp_b200/pipeline.py: layer-loop assembly, column layouts of
GCN/GCN2/GIN/SAGE factors, streamed Gram→J1→back-projection, all-nugget sweep, fused metrics) driven
through a CPU stand-in for the CUDA operator set and checked against the reference's golden outputs.
The CUDA kernels themselves are covered by the ``-m gpu`` tests."""
import pytest
import torch

import cases
import runner
from stub_ops import StubOps


@pytest.mark.parametrize("name", cases.FAST)
def test_host_pipeline_matches_reference(name):
    case = cases.CASES[name]
    data, masks, nugget = cases.build_inputs(case)
    out = runner.run_pipeline(StubOps(), case, data, masks, nugget)
    runner.check_against_golden(name, out, masks)


def test_row_block_streaming_is_exact(monkeypatch):
    """relu_nystrom streams row blocks; forcing tiny blocks must not change the result."""
    import gnngp_b200.pipeline as pl
    case = cases.CASES["tiny-GCN-nys2-L3"]
    data, masks, nugget = cases.build_inputs(case)
    ref = runner.run_pipeline(StubOps(), case, data, masks, nugget)["Q"]
    monkeypatch.setattr(pl, "_row_block", lambda na_ld, n_local: 37)
    got = runner.run_pipeline(StubOps(), case, data, masks, nugget)["Q"]
    assert torch.allclose(ref @ ref.T, got @ got.T, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["tiny-GCN-nys2-L3", "tiny-GCN2-nys2-L3", "tinyreg-GCN2-nys1-L3", "small-GCN-nys4-L2"])
def test_reassociated_layer_schedule_matches_reference(name, monkeypatch):
    """A Ny(E) = (A E) W + A[:, landmarks] (R - E_mm W): the schedule used when the eigensolver is overlapped."""
    import gnngp_b200.pipeline as pl
    monkeypatch.setattr(pl.Pipeline, "force_reassociated", True, raising=False)
    case = cases.CASES[name]
    data, masks, nugget = cases.build_inputs(case)
    out = runner.run_pipeline(StubOps(), case, data, masks, nugget)
    runner.check_against_golden(name, out, masks)


@pytest.mark.parametrize("name", ["tiny-GCN-nys2-L3", "tinyreg-GIN-nys1-L3", "small-SAGE-nys4-L4"])
def test_nugget_parallel_cholesky_fit_matches_reference(name, monkeypatch):
    """The sharded mode's nugget-parallel fit (shifted Cholesky solves instead of the replicated
    eigendecomposition, pipeline._fit_by_shifted_solves) is the same posterior mean: golden check."""
    monkeypatch.setenv("GNNGP_FIT_SOLVER", "cholesky")
    case = cases.CASES[name]
    data, masks, nugget = cases.build_inputs(case)
    out = runner.run_pipeline(StubOps(), case, data, masks, nugget)
    runner.check_against_golden(name, out, masks)


@pytest.mark.parametrize("name", ["tiny-GCN-nys2-L3", "tinyreg-GIN-nys1-L3", "small-SAGE-nys4-L4"])
def test_band_reduction_fit_route_matches_reference(name, monkeypatch):
    """The single-GPU fit route without an eigendecomposition (pipeline._fit_by_band_reduction; the CUDA operator is
    csrc/bandfit.cu, here its dense stand-in) gives the reference's posterior means: golden check."""
    monkeypatch.setenv("GNNGP_FIT_SOLVER", "band")
    case = cases.CASES[name]
    data, masks, nugget = cases.build_inputs(case)
    out = runner.run_pipeline(StubOps(), case, data, masks, nugget)
    runner.check_against_golden(name, out, masks)


@pytest.mark.parametrize("name", ["small-GCN-nys4-L2", "small-SAGE-nys4-L4"])
def test_krylov_fit_route_matches_reference(name, monkeypatch):
    """The block-Lanczos posterior solve (pipeline._fit_by_krylov / krylov_fit.py; factored coefficients, sweep over the
    Krylov dimension) forced on the 176-landmark cases: golden check, whether it converges or hands back."""
    monkeypatch.setenv("GNNGP_FIT_SOLVER", "krylov")
    case = cases.CASES[name]
    data, masks, nugget = cases.build_inputs(case)
    ops = StubOps()
    out = runner.run_pipeline(ops, case, data, masks, nugget)
    runner.check_against_golden(name, out, masks)
    assert getattr(ops, "last_fit_solver", None) in ("krylov", "eigh", "band")


def test_fit_solver_policy(monkeypatch):
    from gnngp_b200.pipeline import Pipeline

    class Comm(object):
        rank = 0

        def __init__(self, world):
            self.world = world

    monkeypatch.delenv("GNNGP_FIT_SOLVER", raising=False)
    assert Pipeline(StubOps(), Comm(1))._fit_solver(15356, 101, 41) == "band"    # one rank: band reduction replaces the eigh
    assert Pipeline(StubOps(), Comm(1))._fit_solver(3053, 101, 41) == "band"
    assert Pipeline(StubOps(), Comm(1))._fit_solver(600, 101, 41) == "eigh"      # small block: eigh is cheap
    assert Pipeline(StubOps(), Comm(1))._fit_solver(15356, 101, 100) == "eigh"   # more targets than the solve kernel holds
    assert Pipeline(StubOps(), Comm(1))._fit_solver(31000, 101, 41) == "eigh"    # beyond the shared-memory panel
    assert Pipeline(StubOps(), Comm(2))._fit_solver(3025, 101, 41) == "band"     # 51 dense solves per rank cost more
    assert Pipeline(StubOps(), Comm(2))._fit_solver(15356, 101, 41) == "band"
    assert Pipeline(StubOps(), Comm(4))._fit_solver(15356, 101, 41) == "band"    # 26 x 47 ms > 0.49 s
    assert Pipeline(StubOps(), Comm(8))._fit_solver(3025, 101, 41) == "cholesky" # 13 x 1.7 ms < 41 ms
    assert Pipeline(StubOps(), Comm(8))._fit_solver(15356, 101, 41) == "band"    # 13 x 47 ms > 0.49 s
    assert Pipeline(StubOps(), Comm(4))._fit_solver(3025, 101, 100) == "cholesky"  # too many targets for the band kernel
    assert Pipeline(StubOps(), Comm(8))._fit_solver(604, 101) == "eigh"          # small block: eigh is cheap
    monkeypatch.setenv("GNNGP_FIT_SOLVER", "cholesky")
    assert Pipeline(StubOps(), Comm(1))._fit_solver(100, 101) == "cholesky"


def test_indefinite_shifted_gram_falls_back_to_eigh(monkeypatch):
    """A Gram block whose rounding noise exceeds the smallest shift is not positive definite: the
    nugget-parallel fit must report it (None) so that every rank takes the eigen path."""
    from gnngp_b200.pipeline import Pipeline
    monkeypatch.setenv("GNNGP_FIT_SOLVER", "cholesky")
    ops = StubOps()
    pipe = Pipeline(ops)
    gram = torch.diag(torch.tensor([4.0, 1.0, -0.5]))
    qty_t = torch.ones(2, 3)
    q = torch.randn(5, 3)
    got = pipe._fit_by_shifted_solves(q, gram, qty_t, torch.ones(1), torch.tensor([0.1, 1.0]), (2,))
    assert got is None
    ok = pipe._fit_by_shifted_solves(q, gram.abs(), qty_t, torch.ones(1), torch.tensor([0.1, 1.0]), (2,))
    want = q @ torch.linalg.solve(gram.abs() + 0.1 * torch.eye(3), qty_t.T)
    assert torch.allclose(ok[0], want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("method", ["GCN", "GIN", "SAGE"])
def test_first_product_reuse_across_hyper_parameters(method):
    """with a product cache attached, a sigma_w / L sweep computes A @ Q0 once and the
    results equal the uncached ones."""
    from gnngp_b200.pipeline import Landmarks, Pipeline, RowLayout

    class Counting(StubOps):
        def __init__(self):
            self.spmm_widths = []

        def spmm(self, csr, B, alpha=1.0, out=None):
            self.spmm_widths.append(int(B.shape[1]))
            return StubOps.spmm(self, csr, B, alpha, out)

    case = cases.CASES["tiny-GCN-nys2-L3"]
    data, masks, _ = cases.build_inputs(case)
    land = Landmarks(masks["landmark"], RowLayout(data.num_nodes))
    ni = data.x.shape[1]
    plain_ops, cached_ops = Counting(), Counting()
    plain, cached = Pipeline(plain_ops), Pipeline(cached_ops)
    cached.product_cache = {}
    adj = torch.sparse_coo_tensor(data.edge_index, data.edge_attr, (data.num_nodes,) * 2)
    csr_a, csr_b = plain_ops.csr_of(adj), cached_ops.csr_of(adj)
    for layers, sigma_b, sigma_w in ((2, 0.1, 1.0), (3, 0.2, 1.3), (2, 0.0, 0.7)):
        want = plain.kernel_nystrom(data.x, csr_a, layers, sigma_b, sigma_w, land, method, {})
        got = cached.kernel_nystrom(data.x, csr_b, layers, sigma_b, sigma_w, land, method, {})
        assert torch.allclose(want @ want.T, got @ got.T, rtol=2e-4, atol=1e-5)
    assert plain_ops.spmm_widths.count(ni) == 3           # the reference's behaviour: once per call
    assert cached_ops.spmm_widths.count(ni) == 1          # reused twice
    # a different initial factor must not hit the cache
    other = data.x * 2.0
    got = cached.kernel_nystrom(other, csr_b, 2, 0.1, 1.0, land, method, {})
    want = plain.kernel_nystrom(other, csr_a, 2, 0.1, 1.0, land, method, {})
    assert torch.allclose(want @ want.T, got @ got.T, rtol=2e-4, atol=1e-5)
    assert cached_ops.spmm_widths.count(ni) == 2
