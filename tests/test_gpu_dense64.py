"""The fp64 dense operators of the Krylov Nystrom root (csrc/dense64.cu) against torch fp64 on the B200, and the
route itself (gnngp_b200/krylov_root.py through the C ABI) against the reference's eigen route (src/compute.py:6-17)."""
import pytest
import torch

from gnngp_b200 import krylov_root

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from gnngp_b200.ops import default_ops
    return default_ops()


def rel(a, b):
    return float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-300))


@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (257, 130, 1000), (1000, 64, 3), (129, 1, 77), (1, 200, 513), (700, 700, 128),
                                   (2000, 128, 20000), (128, 128, 30000)])
def test_dgemm_views_against_torch(ops, m, n, k):
    gen = torch.Generator(device=DEV).manual_seed(m * 7 + n * 3 + k)
    big_a = torch.randn(max(m, k) + 5, max(m, k) + 9, device=DEV, dtype=torch.float64, generator=gen)
    big_b = torch.randn(max(n, k) + 3, max(n, k) + 7, device=DEV, dtype=torch.float64, generator=gen)
    c0 = torch.randn(m, n + 4, device=DEV, dtype=torch.float64, generator=gen)
    for a in (big_a[2:2 + m, 3:3 + k], big_a[1:1 + k, 4:4 + m].T):            # plain and transposed views
        for b in (big_b[1:1 + k, 2:2 + n], big_b[3:3 + n, 1:1 + k].T):
            for alpha, beta in ((1.0, 0.0), (-1.0, 1.0), (0.5, -2.0)):
                out = c0.clone()
                ops.dgemm(a, b, out=out[:, :n], alpha=alpha, beta=beta)
                want = alpha * (a @ b) + beta * c0[:, :n]
                assert rel(out[:, :n], want) < 1e-13, (m, n, k, alpha, beta)
                assert torch.equal(out[:, n:], c0[:, n:])                  # columns outside the view untouched


@pytest.mark.parametrize("m,n,k", [(257, 130, 1000), (129, 66, 77), (1000, 64, 3), (127, 63, 17), (1, 1, 1), (2, 3, 16), (300, 129, 32),
                                   (4000, 128, 4001), (128, 128, 40001)])
def test_dgemm_aligned_views_take_the_fast_path(ops, m, n, k):
    """16-byte aligned operands with even strides (the layouts the band reduction and the Krylov root use): all four
    layout combinations, ragged M / N / K edges, split-K and read-modify-write epilogues."""
    gen = torch.Generator(device=DEV).manual_seed(m * 5 + n * 11 + k)
    e = lambda x: x + (x & 1)                                                  # even
    big_a = torch.randn(e(max(m, k)) + 6, e(max(m, k)) + 8, device=DEV, dtype=torch.float64, generator=gen)
    big_b = torch.randn(e(max(n, k)) + 4, e(max(n, k)) + 10, device=DEV, dtype=torch.float64, generator=gen)
    c0 = torch.randn(m, e(n) + 4, device=DEV, dtype=torch.float64, generator=gen)
    for a in (big_a[2:2 + m, 4:4 + k], big_a[2:2 + k, 4:4 + m].T):
        for b in (big_b[4:4 + k, 2:2 + n], big_b[2:2 + n, 6:6 + k].T):
            for alpha, beta in ((1.0, 0.0), (-1.0, 1.0), (0.5, -2.0)):
                out = c0.clone()
                ops.dgemm(a, b, out=out[:, :n], alpha=alpha, beta=beta)
                want = alpha * (a @ b) + beta * c0[:, :n]
                assert rel(out[:, :n], want) < 1e-13, (m, n, k, alpha, beta, a.stride(), b.stride())
                assert torch.equal(out[:, n:], c0[:, n:])


def test_dgemm_lower_updates_the_lower_triangle(ops):
    gen = torch.Generator(device=DEV).manual_seed(11)
    n, k = 777, 128
    p = torch.randn(n, k, device=DEV, dtype=torch.float64, generator=gen)
    c0 = torch.randn(n, n, device=DEV, dtype=torch.float64, generator=gen)
    out = c0.clone()
    ops.dgemm(p, p.T, out=out, alpha=-1.0, beta=1.0, lower=True)
    want = c0 - p @ p.T
    assert rel(torch.tril(out), torch.tril(want)) < 1e-13
    far = torch.triu(torch.ones_like(out, dtype=torch.bool), 128)            # tiles wholly above the diagonal are skipped
    assert torch.equal(out[far], c0[far])


@pytest.mark.parametrize("n", [1, 37, 128, 129, 500, 1153, 3000])
def test_potrf_against_torch(ops, n):
    gen = torch.Generator(device=DEV).manual_seed(n)
    x = torch.randn(n, n + 8, device=DEV, dtype=torch.float64, generator=gen)
    spd = x @ x.T + 0.1 * torch.eye(n, device=DEV, dtype=torch.float64)
    a = spd.clone()
    info = torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.potrf_f64(a, info)
    want = torch.linalg.cholesky(spd)
    assert int(info) == 0
    assert rel(a, want) < 1e-12
    assert float(torch.triu(a, 1).abs().max()) == 0.0 if n > 1 else True


def test_potrf_flags_an_indefinite_matrix(ops):
    a = torch.eye(300, device=DEV, dtype=torch.float64)
    a[200, 200] = -1.0
    info = torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.potrf_f64(a, info)
    assert int(info) >= 1 and bool(torch.isfinite(a).all())


@pytest.mark.parametrize("nb", [1, 5, 32, 100, 128])
def test_potrf_inv_block(ops, nb):
    gen = torch.Generator(device=DEV).manual_seed(nb)
    x = torch.randn(nb, nb + 4, device=DEV, dtype=torch.float64, generator=gen)
    spd = x @ x.T + 0.05 * torch.eye(nb, device=DEV, dtype=torch.float64)
    g = spd.clone()
    info = torch.zeros(1, dtype=torch.int32, device=DEV)
    linv = ops.potrf_inv_block(g, info)
    want = torch.linalg.cholesky(spd)
    assert int(info) == 0
    assert rel(torch.tril(g), want) < 1e-12
    assert rel(linv @ want, torch.eye(nb, device=DEV, dtype=torch.float64)) < 1e-11
    assert float(torch.triu(linv, 1).abs().max()) == 0.0 if nb > 1 else True


def test_small_helpers(ops):
    gen = torch.Generator(device=DEV).manual_seed(5)
    g32 = torch.randn(333, 333, device=DEV, generator=gen)
    sym = ops.sym_f64(g32)
    low = torch.tril(g32).double()
    assert torch.equal(sym, low + torch.tril(low, -1).T)
    a = torch.randn(301, 77, device=DEV, dtype=torch.float64, generator=gen)
    assert torch.equal(ops.transpose_scale_f64(a, 0.25), 0.25 * a.T)
    d = torch.randn(77, device=DEV, dtype=torch.float64, generator=gen)
    assert torch.equal(ops.scale_cols_f64(a, d), a * d)
    assert torch.equal(ops.to_f32(a), a.float())
    v = torch.randn(301, 77, device=DEV, dtype=torch.float64, generator=gen)
    got = ops.residual_sumsq(a, v, d)
    assert rel(got, ((a - v * d) ** 2).sum(0)) < 1e-13


@pytest.mark.parametrize("rows,n,k", [(300, 1000, 128), (2500, 2500, 300), (700, 4000, 647)])
def test_qr_rotate(ops, rows, n, k):
    """X Q_H: an orthogonal transformation whose leading columns are +-(X z_j) for orthonormal z_j."""
    gen = torch.Generator(device=DEV).manual_seed(rows + n + k)
    zq, _ = torch.linalg.qr(torch.randn(n, k, device=DEV, dtype=torch.float64, generator=gen))
    kp = -(-k // 128) * 128
    z = torch.zeros(n, kp, device=DEV, dtype=torch.float64)
    z[:, :k] = zq
    x0 = torch.randn(rows, n, device=DEV, dtype=torch.float64, generator=gen)
    x = x0.clone()
    ops.qr_rotate(x, z)
    assert rel(x @ x.T, x0 @ x0.T) < 1e-12                                     # orthogonal
    lead = x0 @ zq
    assert rel(x[:, :k].abs(), lead.abs()) < 1e-11
    assert rel(x[:, :k] * torch.sign((x[:, :k] * lead).sum(0)), lead) < 1e-11
    rest = x[:, k:]                                                            # = X Z_perp: orthogonal to the lead block's range
    assert rel(rest @ rest.T + lead @ lead.T, x0 @ x0.T) < 1e-11


def _block(n, above, seed=0, gap=0.5):
    gen = torch.Generator(device=DEV).manual_seed(seed)
    q, _ = torch.linalg.qr(torch.randn(n, n, device=DEV, dtype=torch.float64, generator=gen))
    low = torch.exp(torch.empty(n - above, device=DEV, dtype=torch.float64).uniform_(float(torch.log(torch.tensor(8e-6))),
                                                                                     float(torch.log(torch.tensor(gap * 1e-4))), generator=gen))
    high = torch.exp(torch.empty(above - 1, device=DEV, dtype=torch.float64).uniform_(float(torch.log(torch.tensor(1.2e-4))), 0.0, generator=gen))
    lam = torch.cat([low, high, torch.ones(1, device=DEV, dtype=torch.float64)])
    m = (q * lam) @ q.T
    return ((m + m.T) / 2).float()


@pytest.mark.parametrize("n,above,block", [(1500, 60, 32), (4300, 150, 128)])
def test_krylov_root_against_the_eigen_route(ops, n, above, block):
    """Q Qᵀ of the factor from the Krylov route vs the reference formula in fp64 (landmark rows, other rows, cross block)."""
    kmm = _block(n, above)
    gen = torch.Generator(device=DEV).manual_seed(3)
    k_cols = (torch.randn(96, n, device=DEV, dtype=torch.float64, generator=gen) @ kmm.double()).float()
    stats = {}
    got = krylov_root.root_factors(ops, kmm, stats, block=block)
    assert got is not None and stats["converged"], stats
    w_t, land = got
    assert stats["ritz_above_clamp"][-1] == above
    sym = torch.tril(kmm).double()
    sym = sym + torch.tril(sym, -1).T
    d, v = torch.linalg.eigh(sym)
    d_sqrt = torch.sqrt(torch.clamp(d, 0))
    ref_land = v * d_sqrt
    ref_rows = k_cols.double() @ v * (d_sqrt / torch.clamp(d, d[-1] * 1e-4))
    rows = k_cols.double() @ w_t.double().T
    land = land.double()
    assert rel(land @ land.T, ref_land @ ref_land.T) < 2e-6
    assert rel(rows @ rows.T, ref_rows @ ref_rows.T) < 2e-6
    assert rel(rows @ land.T, ref_rows @ ref_land.T) < 2e-6


def test_krylov_root_hands_back_what_it_cannot_do(ops):
    kmm = _block(1500, 60)
    bad = kmm.clone()
    bad[7, 7] = -1.0
    assert krylov_root.root_factors(ops, bad, block=32) is None
    stats = {}
    assert krylov_root.root_factors(ops, _block(1500, 1200, gap=0.9), stats, block=32) is None and not stats["converged"]


@pytest.mark.parametrize("dim,half,c,g", [(700, 81, 41, 5), (129, 128, 3, 2), (2400, 95, 48, 101), (50, 20, 7, 3)])
def test_band_solve_against_dense_solves(ops, dim, half, c, g):
    """(band(H) + shift I) x = rhs for every shift: block-tridiagonal-like SPD matrices, entries beyond the band ignored."""
    gen = torch.Generator(device=DEV).manual_seed(dim + half)
    a = torch.randn(dim, dim, device=DEV, dtype=torch.float64, generator=gen)
    i = torch.arange(dim, device=DEV)
    band = (i.view(-1, 1) - i.view(1, -1)).abs() <= half
    h = torch.where(band, (a + a.T) / 2, torch.zeros_like(a)) + (2.0 * half) * torch.eye(dim, device=DEV, dtype=torch.float64)
    noisy = h + torch.where(band, torch.zeros_like(a), torch.full_like(a, 3.0))       # beyond |i-j| <= 128 nothing is read
    far = (i.view(-1, 1) - i.view(1, -1)).abs() > 128
    noisy = torch.where(far, noisy, h)
    rhs_t = torch.randn(c, dim, device=DEV, dtype=torch.float64, generator=gen)
    shifts = torch.logspace(-2, 1, g, device=DEV, dtype=torch.float64)
    sol, info = ops.band_solve(noisy, rhs_t, shifts)
    assert int(info) == 0
    eye = torch.eye(dim, device=DEV, dtype=torch.float64)
    for k in range(g):
        want = torch.linalg.solve(h + shifts[k] * eye, rhs_t.T).T
        assert rel(sol[k * c:(k + 1) * c], want) < 1e-11, (dim, k)


def test_krylov_fit_against_dense_solves(ops):
    """The block-Lanczos posterior solve (krylov_fit.py) through the C ABI on a Gram matrix with the benchmark's spectrum
    shape: every nugget's coefficient block against a dense fp64 solve."""
    from gnngp_b200 import krylov_fit
    m, c, outliers = 4200, 41, 160
    gen = torch.Generator(device=DEV).manual_seed(9)
    q, _ = torch.linalg.qr(torch.randn(m, m, device=DEV, dtype=torch.float64, generator=gen))
    lam = torch.cat([torch.empty(m - outliers, device=DEV, dtype=torch.float64).uniform_(2e-8, 2e-7, generator=gen),
                     torch.exp(torch.empty(outliers, device=DEV, dtype=torch.float64).uniform_(-11.5, 0.0, generator=gen))]) * 5e4
    g64 = (q * lam) @ q.T
    gram = ((g64 + g64.T) / 2).float()
    qty_t = (g64 @ torch.randn(m, c, device=DEV, dtype=torch.float64, generator=gen)).float().T.contiguous()
    d = float(torch.diagonal(gram).double().mean())
    shifts = torch.logspace(-3, 1, 101, device=DEV, dtype=torch.float64) * d
    stats = {}
    got = krylov_fit.coefficients(ops, gram, qty_t, shifts, stats)
    assert got is not None and stats["converged"], stats
    y_t, basis_t = got
    got = y_t.double() @ basis_t.double()                              # [(G*C) x m]
    sym = torch.tril(gram).double()
    sym = sym + torch.tril(sym, -1).T
    eye = torch.eye(m, device=DEV, dtype=torch.float64)
    for k in (0, 17, 50, 100):
        want = torch.linalg.solve(sym + shifts[k] * eye, qty_t.double().T).T
        assert rel(got[k * c:(k + 1) * c], want) < 2e-6, (k, stats)
