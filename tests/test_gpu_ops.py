"""Operator-level parity on the B200: every C-ABI kernel against the same operator restated in plain
torch on the CPU (tests/stub_ops.py) on seeded inputs, including ragged / empty / misaligned shapes."""
import math

import numpy as np
import pytest
import torch

from stub_ops import StubOps

pytestmark = pytest.mark.gpu

TOL = 2e-5   # fp32 kernels vs fp32 torch: relative Frobenius


@pytest.fixture(scope="module")
def ops():
    from gnngp_b200.ops import default_ops
    o = default_ops()
    o.set_gemm_engine(0)      # fp32 CUDA-core tiles; the tcgen05 engine has its own test file
    yield o
    o.set_gemm_engine(2)


@pytest.fixture(scope="module")
def ref():
    return StubOps()


def rel(a, b):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def rand_graph(n, nnz, seed, hub=None, empty_rows=()):
    g = torch.Generator().manual_seed(seed)
    row = torch.randint(0, n, (nnz,), generator=g)
    col = torch.randint(0, n, (nnz,), generator=g)
    if hub is not None:            # one very long row, like the power-law hubs of the Reddit shape
        row[: nnz // 3] = hub
    for r in empty_rows:
        row[row == r] = (r + 1) % n
    val = torch.randn(nnz, generator=g)
    return row, col, val


@pytest.mark.parametrize("n,nnz,hub", [(64, 400, None), (1000, 30000, 7), (5000, 200000, 123), (300, 0, None)])
def test_csr_from_coo_and_spmm(ops, ref, n, nnz, hub):
    row, col, val = rand_graph(n, nnz, seed=n + nnz, hub=hub, empty_rows=(0, n - 1))
    csr = ops.csr_from_coo(row.cuda(), col.cuda(), val.cuda(), n, n)
    rp = csr.rowptr.cpu()
    assert rp[0] == 0 and rp[-1] == nnz and bool((rp[1:] >= rp[:-1]).all())
    counts = torch.bincount(row, minlength=n) if nnz else torch.zeros(n, dtype=torch.long)
    assert torch.equal(rp[1:] - rp[:-1], counts)
    if nnz:
        # sorted by (row, col), values carried along
        rows_sorted = torch.repeat_interleave(torch.arange(n), counts)
        key = rows_sorted * n + csr.colidx.cpu().long()
        assert bool((key[1:] >= key[:-1]).all())
        dense = torch.zeros(n, n, dtype=torch.float64)
        dense.index_put_((rows_sorted, csr.colidx.cpu().long()), csr.vals.cpu().double(), accumulate=True)
        want = torch.zeros(n, n, dtype=torch.float64)
        want.index_put_((row, col), val.double(), accumulate=True)
        assert torch.allclose(dense, want, atol=1e-6)
    cref = ref.csr_from_coo(row, col, val, n, n)
    for k in (1, 3, 32, 37, 64, 130, 257):
        g = torch.Generator().manual_seed(k)
        b = torch.randn(n, k, generator=g)
        want = ref.spmm(cref, b, 0.7)
        got = ops.spmm(csr, b.cuda(), 0.7)
        assert got.shape == want.shape
        assert rel(got, want) < TOL or float(want.norm()) == 0.0, (n, nnz, k, rel(got, want))


@pytest.mark.parametrize("tile,chunk", [(32, 32), (64, 64), (128, 256), (0, 0)])
def test_spmm_tiles_alignment_and_strided_output(ops, ref, tile, chunk):
    n = 3000
    row, col, val = rand_graph(n, 150000, seed=5, hub=11)
    csr = ops.csr_from_coo(row.cuda(), col.cuda(), val.cuda(), n, n)
    cref = ref.csr_from_coo(row, col, val, n, n)
    ops.spmm_tile_cols, ops.spmm_nnz_per_warp = tile, chunk
    try:
        for k, pad in ((602, 0), (602, 2), (301, 3), (129, 3), (200, 56)):
            g = torch.Generator().manual_seed(k + pad)
            store = torch.randn(n, k + pad, generator=g)
            b = store[:, :k]                        # leading dimension k + pad: 16 B / 8 B / 4 B aligned rows
            want = ref.spmm(cref, b.contiguous(), 1.3)
            big = torch.full((n, k + 9), 7.0).cuda()
            got = ops.spmm(csr, b.cuda() if pad == 0 else store.cuda()[:, :k], 1.3, out=big[:, 5:5 + k])
            assert rel(got, want) < TOL, (tile, chunk, k, pad, rel(got, want))
            assert bool((big[:, :5] == 7.0).all()) and bool((big[:, 5 + k:] == 7.0).all()), "wrote outside the view"
    finally:
        ops.spmm_tile_cols, ops.spmm_nnz_per_warp = 0, 0


@pytest.mark.parametrize("flags", [1, 2, 3, 4, 6])
def test_spmm_tuning_flags_do_not_change_results(ops, ref, flags):
    """gnngp_spmm_set_flags only changes store width, which rows are zero-filled and which kernel walks short rows:
    every flag combination must reproduce the plain kernels (flags 0) on a graph with a hub row, cut rows and empty
    rows and on a short-row graph, on both gather paths, aligned and misaligned outputs, and must not leave a row of
    a garbage-filled output untouched."""
    n = 4000
    row, col, val = rand_graph(n, 300000, seed=17, hub=29, empty_rows=(0, 1, 2, 777, n - 1))
    csr = ops.csr_from_coo(row.cuda(), col.cuda(), val.cuda(), n, n)
    cref = ref.csr_from_coo(row, col, val, n, n)
    short_row, short_col, short_val = rand_graph(n, 9 * n, seed=18, empty_rows=(5, 6, n - 2))
    short = ops.csr_from_coo(short_row.cuda(), short_col.cuda(), short_val.cuda(), n, n)
    short_ref = ref.csr_from_coo(short_row, short_col, short_val, n, n)
    previous = ops.set_spmm_flags(0)
    try:
        for matrix, mref in ((csr, cref), (short, short_ref)):
            for k, lead, shift, tile, chunk in ((320, 320, 0, 0, 0), (321, 324, 4, 0, 64), (256, 256, 0, 64, 0), (130, 136, 4, 128, 32),
                                                 (320, 323, 1, 0, 0), (77, 80, 0, 32, 0), (256, 260, 0, 128, 512), (384, 384, 0, 0, 2048)):
                g = torch.Generator().manual_seed(k + lead)
                b = torch.randn(n, k, generator=g)
                ops.spmm_tile_cols, ops.spmm_nnz_per_warp = tile, chunk
                results = []
                for f in (0, flags):
                    ops.set_spmm_flags(f)
                    big = torch.full((n, lead + 8), float("nan")).cuda()
                    results.append(ops.spmm(matrix, b.cuda(), 0.9, out=big[:, shift:shift + k]).clone())
                    edge = torch.cat([big[:, :shift], big[:, shift + k:]], 1)
                    assert bool(torch.isnan(edge).all()), "wrote outside the view"
                assert not bool(torch.isnan(results[1]).any()), (flags, k, lead, shift, tile, chunk)
                assert rel(results[1], results[0]) < 1e-6, (flags, k, lead, shift, tile, chunk)
                assert rel(results[1], ref.spmm(mref, b, 0.9)) < TOL
    finally:
        ops.set_spmm_flags(previous)
        ops.spmm_tile_cols, ops.spmm_nnz_per_warp = 0, 0


def test_spmm_duplicates_are_summed(ops):
    row = torch.tensor([0, 0, 0, 2, 2]); col = torch.tensor([1, 1, 3, 0, 0]); val = torch.tensor([1., 2., 4., 8., 16.])
    csr = ops.csr_from_coo(row.cuda(), col.cuda(), val.cuda(), 4, 4)
    b = torch.eye(4).cuda()
    got = ops.spmm(csr, b).cpu()
    want = torch.zeros(4, 4); want[0, 1] = 3; want[0, 3] = 4; want[2, 0] = 24
    assert torch.equal(got, want)


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (5, 7, 3), (128, 128, 16), (130, 257, 70), (513, 300, 1031), (64, 2000, 257)])
def test_gemm_nt_simt(ops, ref, M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    got = ops.gemm_nt(a.cuda(), b.cuda(), alpha=0.5)
    assert rel(got, 0.5 * a @ b.T) < TOL
    # strided operands / output view, beta accumulate
    big = torch.randn(M, N + 6, generator=g)
    out = big.cuda()
    ops.gemm_nt(a.cuda(), b.cuda(), alpha=2.0, out=out[:, 3:3 + N], beta=1.0)
    want = big.clone(); want[:, 3:3 + N] += 2.0 * a @ b.T
    assert rel(out, want) < TOL


def test_gram_arccos_and_dense_arccos(ops, ref):
    g = torch.Generator().manual_seed(1)
    q = torch.randn(700, 45, generator=g)
    idx = torch.randperm(700, generator=g)[:130].sort().values
    l = q[idx]
    s, t = ref.row_norms(q), ref.row_norms(l)
    got = ops.gram_arccos(q.cuda(), l.cuda(), ops.row_norms(q.cuda()), ops.row_norms(l.cuda()))
    want = ref.gram_arccos(q, l, s, t)
    assert rel(got, want) < TOL
    # the landmark block must be symmetric to rounding and have s_i^2 / 2 on its diagonal (J1(0) = 1/2)
    blk = got.cpu()[idx]
    assert rel(blk, blk.T) < 1e-6
    assert torch.allclose(torch.diag(blk), 0.5 * t * t, rtol=2e-5)
    k = (q[:300] @ q[:300].T)
    got = ops.arccos_dense(k.cuda())
    assert rel(got, ref.arccos_dense(k)) < TOL
    # in place
    kk = k.cuda()
    ops.arccos_dense(kk, out=kk)
    assert rel(kk, ref.arccos_dense(k)) < TOL


def test_arccos_nan_semantics_match_reference(ops):
    """A zero-norm row gives 0/0 = NaN in the reference (src/compute.py:125): propagated, not masked."""
    q = torch.randn(40, 6); q[3] = 0
    l = q[[1, 3, 5]]
    got = ops.gram_arccos(q.cuda(), l.cuda(), ops.row_norms(q.cuda()), ops.row_norms(l.cuda())).cpu()
    assert bool(torch.isnan(got[3]).all()) and bool(torch.isnan(got[:, 1]).all())
    assert not bool(torch.isnan(got[[0, 1, 2, 4]][:, [0, 2]]).any())


@pytest.mark.parametrize("kind,params", [("linear", {}), ("rbf", dict(gamma=0.01)), ("laplacian", dict(gamma=0.02)),
                                         ("sigmoid", dict(gamma=0.01, c=0.1)), ("polynomial", dict(c=1.0, d=2.0)),
                                         ("polynomial", dict(c=5.0, d=3.0))])
def test_initial_kernels(ops, ref, kind, params):
    g = torch.Generator().manual_seed(2)
    x, r = torch.randn(333, 41, generator=g), torch.randn(77, 41, generator=g)
    got = ops.initial_kernel(kind, x.cuda(), r.cuda(), **params)
    assert rel(got, ref.initial_kernel(kind, x, r, **params)) < 5e-5
    xn, rn = x / x.norm(dim=1, keepdim=True), r / r.norm(dim=1, keepdim=True)
    assert rel(ops.initial_kernel("arccos", xn.cuda(), rn.cuda()), ref.initial_kernel("arccos", xn, rn)) < 5e-5


def test_row_and_column_utilities(ops, ref):
    g = torch.Generator().manual_seed(3)
    m = torch.randn(501, 67, generator=g)
    idx = torch.randperm(501, generator=g)[:120].sort().values
    cidx = torch.randperm(67, generator=g)[:20].sort().values
    mc = m.cuda()
    assert rel(ops.row_norms(mc), ref.row_norms(m)) < 1e-6
    assert rel(ops.row_norms(mc, sqrt=False), ref.row_norms(m, sqrt=False)) < 1e-6
    assert abs(float(ops.mean(ops.row_norms(mc))) - float(ref.row_norms(m).mean())) < 1e-5
    sq = torch.randn(90, 90, generator=g)
    assert abs(float(ops.diag_mean(sq.cuda())) - float(torch.diag(sq).mean())) < 1e-6
    assert torch.equal(ops.gather_rows(mc, idx.cuda()).cpu(), m[idx])
    assert torch.equal(ops.gather_rows_t(mc, idx.cuda()).cpu(), m[idx].T)
    assert torch.equal(ops.transpose(mc).cpu(), m.T)
    assert torch.equal(ops.gather_cols(mc, cidx.cuda()).cpu(), m[:, cidx])
    dst = torch.zeros(501, 67).cuda()
    ops.scatter_rows(dst, idx.cuda(), ops.gather_rows(mc, idx.cuda()))
    want = torch.zeros(501, 67); want[idx] = m[idx]
    assert torch.equal(dst.cpu(), want)
    q = torch.zeros(501, 68).cuda()[:, :67]
    ops.fill_col(q, 66, 0.25)
    assert bool((q[:, 66] == 0.25).all()) and float(q[:, :66].abs().sum()) == 0.0
    y = torch.randn(501, 67, generator=g)
    out = ops.empty(501, 67, mc.device)
    ops.affine(out, mc, y.cuda(), 0.3, -1.5, 2.0)
    assert rel(out, 0.3 * m - 1.5 * y + 2.0) < 1e-6
    ops.affine(out, mc, None, 2.0, 0.0, 0.0)
    assert rel(out, 2.0 * m) < 1e-7
    d = torch.randn(67, generator=g)
    assert rel(ops.scale_cols(mc, d.cuda()), m * d) < 1e-7
    assert rel(ops.scale_cols(mc, d.cuda(), transpose=True), (m * d).T) < 1e-7
    ev = torch.sort(torch.randn(64, generator=g)).values
    ev[:5] = torch.tensor([-1e-3, -1e-6, 0.0, 1e-7, 1e-5])
    ev = torch.sort(ev).values
    r1, i1 = ops.nystrom_spectrum(ev.cuda())
    r2, i2 = ref.nystrom_spectrum(ev)
    assert rel(r1, r2) < 1e-7 and rel(i1, i2) < 1e-6


def test_targets_sweep_and_metrics(ops, ref):
    g = torch.Generator().manual_seed(4)
    n, m, C, G = 400, 33, 5, 7
    y = torch.randint(0, C, (n,), generator=g)
    idx = torch.randperm(n, generator=g)[:150].sort().values
    assert torch.equal(ops.onehot_t(y.cuda(), idx.cuda(), C).cpu(), ref.onehot_t(y, idx, C))
    temp = torch.randn(m, C, generator=g)
    w = torch.rand(m, generator=g) + 0.1
    eps = torch.logspace(-3, 1, G)
    d = torch.tensor([1.7])
    s1 = ops.nugget_coeff(temp.cuda(), w.cuda(), eps.cuda(), d.cuda())
    assert rel(s1, ref.nugget_coeff(temp, w, eps, d)) < 1e-6
    q = torch.randn(n, m, generator=g)
    bt = torch.randn(G * C, m, generator=g)
    f1 = ops.nugget_sweep(q.cuda(), bt.cuda(), G, C)
    f2 = ref.nugget_sweep(q, bt, G, C)
    assert f1.shape == (G, n, C) and rel(f1, f2) < TOL
    member = torch.randint(0, 8, (n,), generator=g).to(torch.uint8)
    f2[1, 5, :] = f2[1, 5, 0]                    # exact tie -> first index, like torch.argmax
    f2[2, 9, 3] = float("nan")                   # NaN is maximal for torch.argmax
    c1 = ops.argmax_count(f2.cuda(), y.cuda(), member.cuda()).cpu()
    assert torch.equal(c1, ref.argmax_count(f2, y, member))
    yr = torch.randn(n, generator=g)
    fr = torch.randn(G, n, generator=g)
    e1 = ops.sqerr_sum(fr.cuda(), yr.cuda(), member.cuda()).cpu()
    assert torch.allclose(e1, ref.sqerr_sum(fr, yr, member), rtol=1e-6)


def test_cpu_tensors_are_rejected(ops):
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.row_norms(torch.randn(4, 4))
    with pytest.raises(RuntimeError):
        ops.gemm_nt(torch.randn(4, 4).cuda(), torch.randn(4, 5).cuda())


def test_gemm_nt_split_k_for_skinny_long_k(ops):
    """[41 x 3000 x 150k]-like products (Q_b^T y_b of the fit) go through the split-K path."""
    g = torch.Generator().manual_seed(77)
    a, b = torch.randn(41, 60000, generator=g), torch.randn(515, 60000, generator=g)
    got = ops.gemm_nt(a.cuda(), b.cuda(), alpha=1.5)
    assert rel(got, 1.5 * (a.double() @ b.double().T)) < 2e-6
    out = torch.full((41, 520), 3.0).cuda()
    ops.gemm_nt(a.cuda(), b.cuda(), out=out[:, 2:517])
    assert rel(out[:, 2:517], a.double() @ b.double().T) < 2e-6
    assert bool((out[:, :2] == 3.0).all()) and bool((out[:, 517:] == 3.0).all())


@pytest.mark.parametrize("n,C,G", [(1000, 41, 5), (257, 3, 2), (300, 120, 3)])
def test_argmax_count_variants(ops, ref, n, C, G):
    g = torch.Generator().manual_seed(n + C)
    f = torch.randn(G, n, C, generator=g)
    f[0, 3, :] = 1.25                      # all tied -> index 0
    f[G - 1, n - 1, C - 1] = float("nan")  # NaN wins
    y = torch.randint(0, C, (n,), generator=g)
    y[3] = 0
    y[n - 1] = C - 1
    member = torch.randint(0, 16, (n,), generator=g).to(torch.uint8)
    got = ops.argmax_count(f.cuda(), y.cuda(), member.cuda()).cpu()
    assert torch.equal(got, ref.argmax_count(f, y, member))


def test_panel_entry_points_single_rank(ops, ref):
    """gnngp_panelize_peers_f32 + gnngp_spmm_panels_f32 with a 'world' of two blocks living on this GPU."""
    n, k = 3001, 300
    row, col, val = rand_graph(n, 400000, seed=9, hub=5)
    csr = ops.csr_from_coo(row.cuda(), col.cuda(), val.cuda(), n, n)
    cref = ref.csr_from_coo(row, col, val, n, n)
    g = torch.Generator().manual_seed(1)
    b = torch.randn(n, k, generator=g)
    rpr = (n + 1) // 2
    ld = 300
    blocks = [torch.zeros(rpr, ld).cuda() for _ in range(2)]
    blocks[0][:rpr] = b[:rpr].cuda()
    blocks[1][: n - rpr] = b[rpr:].cuda()
    ptrs = torch.tensor([blk.data_ptr() for blk in blocks], dtype=torch.int64).cuda()
    ws, panels = ops.panels_from_peers(ptrs.data_ptr(), 2, rpr, ld, n, k, b.cuda().device)
    out = torch.full((n, k + 3), 5.0).cuda()
    ops.spmm_panels(csr, panels, k, 0.5, out[:, 1:1 + k])
    want = ref.spmm(cref, b, 0.5)
    assert rel(out[:, 1:1 + k], want) < TOL
    assert bool((out[:, 0] == 5.0).all()) and bool((out[:, 1 + k:] == 5.0).all())


@pytest.mark.parametrize("m,nrhs,batch", [(1, 1, 1), (31, 7, 2), (32, 8, 1), (33, 9, 3), (100, 41, 2), (257, 3, 1),
                                          (1000, 41, 4)])
def test_potrs_batched_matches_cholesky_solve(ops, m, nrhs, batch):
    """fp64 batched substitution kernel vs torch.cholesky_solve (cuSOLVER potrs); tolerance 1e-11 relative."""
    g = torch.Generator().manual_seed(m * 100 + nrhs)
    a = torch.randn(m, m + 5, generator=g, dtype=torch.float64)
    spd = a @ a.T / (m + 5)
    shifts = torch.logspace(-3, 0, batch, dtype=torch.float64)
    factors = torch.stack([torch.linalg.cholesky(spd + s * torch.eye(m, dtype=torch.float64)) for s in shifts]).cuda()
    # the kernel must not read the strict upper triangle: poison it
    poison = torch.triu(torch.full((m, m), float("nan"), dtype=torch.float64), 1).cuda()
    factors_poisoned = torch.where(torch.isnan(poison), poison, factors)
    rhs = torch.randn(m, nrhs, generator=g, dtype=torch.float64).cuda()
    got = ops.potrs_batched(factors_poisoned, rhs)
    assert got.shape == (batch, m, nrhs)
    for b in range(batch):
        want = torch.cholesky_solve(rhs, factors[b])
        assert rel(got[b], want) < 1e-11
        resid = (spd.cuda() + shifts[b].cuda() * torch.eye(m, dtype=torch.float64, device="cuda")) @ got[b] - rhs
        assert float(resid.norm() / rhs.norm()) < 1e-9


def test_shifted_solves_match_stub(ops, ref):
    """ops.shifted_solves (potrf through torch + the batched substitution kernel) vs the torch restatement."""
    g = torch.Generator().manual_seed(5)
    q = torch.randn(400, 150, generator=g)
    gram = torch.tril(q.T @ q)                        # lower triangle only, like syrk's output
    rhs_t = torch.randn(7, 150, generator=g)
    shifts = torch.tensor([0.01, 0.3, 5.0], dtype=torch.float64)
    want, bad_ref = ref.shifted_solves(gram, rhs_t, shifts)
    got, bad = ops.shifted_solves(gram.cuda(), rhs_t.cuda(), shifts.cuda())
    assert int(bad) == 0 and int(bad_ref) == 0
    assert got.shape == want.shape and rel(got, want) < 1e-5
    # an indefinite shifted matrix is reported, not hidden
    _, bad = ops.shifted_solves(torch.diag(torch.tensor([1.0, -2.0, 3.0])).cuda(), torch.ones(2, 3).cuda(),
                                torch.tensor([0.5], dtype=torch.float64).cuda())
    assert int(bad) == 1


# ---------------------------------------------------------------------------------------------------------
# K9: the eigen-free posterior solve (csrc/bandfit.cu)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,c,g,rank", [(700, 5, 7, 300), (129, 3, 4, 129), (1, 2, 3, 1), (128, 4, 2, 60), (1500, 41, 11, 400),
                                        (257, 48, 3, 257), (3053, 41, 101, 700)])
def test_band_fit_matches_dense_fp64_solves(ops, m, c, g, rank):
    """(G + shift I)^-1 rhs for every shift against torch's fp64 dense solve, on PSD Gram matrices that are
    rank-deficient like Q_b^T Q_b (most eigenvalues far below the shifts), ragged sizes around the band width."""
    gen = torch.Generator(device="cuda").manual_seed(m + c)
    x = torch.randn(m, rank, device="cuda", generator=gen, dtype=torch.float64)
    x = x * torch.logspace(0, -4, rank, device="cuda", dtype=torch.float64)
    gram64 = x @ x.T
    gram = torch.tril(gram64).float()                               # only the lower triangle is meaningful
    gram = gram + torch.triu(torch.full_like(gram, 7.0), 1)         # garbage above the diagonal must be ignored
    rhs_t = torch.randn(c, m, device="cuda", generator=gen)
    scale = float(gram64.diagonal().mean())
    shifts = torch.logspace(-3, 1, g, device="cuda", dtype=torch.float64) * scale
    got, info = ops.band_fit(gram, rhs_t, shifts)
    torch.cuda.synchronize()
    assert int(info) == 0
    sym = torch.tril(gram).double()
    sym = sym + torch.tril(sym, -1).T
    eye = torch.eye(m, device="cuda", dtype=torch.float64)
    for i in range(g):
        want = torch.linalg.solve(sym + shifts[i] * eye, rhs_t.double().T).T          # [c x m]
        block = got[i * c:(i + 1) * c].double()
        err = float((block - want).norm() / want.norm())
        assert err < 5e-7, (m, i, err)                                                 # fp32 output rounding only


def test_band_fit_reports_an_indefinite_shifted_matrix(ops):
    gram = torch.diag(torch.tensor([4.0, 1.0, -0.5, 2.0] * 50, device="cuda"))
    rhs_t = torch.ones(2, 200, device="cuda")
    _, info = ops.band_fit(gram, rhs_t, torch.tensor([0.1, 1.0], device="cuda", dtype=torch.float64))
    assert int(info) == 1                                             # the shift 0.1 leaves -0.4 on the diagonal
