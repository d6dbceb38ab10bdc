"""The tcgen05 / TMEM / TMA 3xTF32 contraction engine against the fp32 CUDA-core engine and fp64 torch.
Kept in its own file so a protocol bug (trap) in the tensor-core kernel cannot take the other GPU
tests' process down with it."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from gnngp_b200.ops import default_ops
    o = default_ops()
    prev = o.set_gemm_engine(1)
    yield o
    o.set_gemm_engine(prev)


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (128, 256, 64), (256, 512, 96), (100, 200, 50), (1, 1, 1),
                                   (1000, 700, 333), (4096, 3069, 3070), (20000, 520, 1027)])
def test_tc_gemm_matches_fp64(ops, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    a, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    want = (a.double() @ b.double().T)
    got = ops.gemm_nt(a.cuda(), b.cuda())
    torch.cuda.synchronize()
    err = rel(got, want)
    # fp32 SGEMM reaches ~1e-7 on these shapes; 3xTF32 must be in the same class (one TF32 pass is ~3e-4)
    assert err < 2e-6, (M, N, K, err)


def test_tc_operand_split_is_fp32_accurate_on_wide_dynamic_range(ops):
    g = torch.Generator().manual_seed(9)
    a = torch.randn(512, 700, generator=g) * torch.logspace(-6, 6, 700)
    b = torch.randn(384, 700, generator=g) * torch.logspace(6, -6, 700)
    got = ops.gemm_nt(a.cuda(), b.cuda())
    assert rel(got, a.double() @ b.double().T) < 2e-6


def test_tc_gram_arccos_and_sweep(ops):
    from stub_ops import StubOps
    ref = StubOps()
    g = torch.Generator().manual_seed(11)
    q = torch.randn(3000, 515, generator=g)
    idx = torch.randperm(3000, generator=g)[:514].sort().values
    l = q[idx]
    got = ops.gram_arccos(q.cuda(), l.cuda(), ops.row_norms(q.cuda()), ops.row_norms(l.cuda()))
    want = ref.gram_arccos(q.double(), l.double(), ref.row_norms(q.double()), ref.row_norms(l.double()))
    assert rel(got, want) < 5e-6
    G, C = 11, 7
    bt = torch.randn(G * C, 515, generator=g)
    f = ops.nugget_sweep(q.cuda(), bt.cuda(), G, C)
    assert rel(f, ref.nugget_sweep(q.double(), bt.double(), G, C)) < 2e-6


def test_tc_engine_end_to_end_case(ops):
    import cases
    import runner
    from test_gpu_parity import run_public_api
    name = "pubmed-GCN-nys4-L2"
    case = cases.CASES[name]
    data, masks, nugget = cases.build_inputs(case)
    out, _ = run_public_api(case, data, masks, nugget)
    runner.check_against_golden(name, out, masks)


def test_tc_syrk_lower_triangle(ops):
    g = torch.Generator().manual_seed(21)
    a = torch.randn(1500, 4000, generator=g)
    got = ops.syrk(a.cuda()).cpu().double()
    want = a.double() @ a.double().T
    low = torch.tril(torch.ones(1500, 1500, dtype=torch.bool))
    assert float((got[low] - want[low]).norm() / want[low].norm()) < 2e-6
    w1, _ = torch.linalg.eigh(got)             # eigh reads the lower triangle only
    w2, _ = torch.linalg.eigh(want)
    assert float((w1 - w2).abs().max() / w2.abs().max()) < 1e-6


@pytest.mark.parametrize("n,K", [(3053, 40000), (129, 9000), (385, 70000), (2000, 200)])
def test_tc_syrk_split_k_and_odd_tile_counts(ops, n, K):
    """Lower-triangle enumeration of 128 x 256 tiles (odd and even row-block counts) and the split-K route a small
    tile grid with a long contraction takes (atomic accumulation into the zeroed output)."""
    g = torch.Generator(device="cuda").manual_seed(n + K)
    a = torch.randn(n, K, device="cuda", generator=g) + 0.25
    got = ops.syrk(a).double()
    want = a.double() @ a.double().T
    low = torch.tril(torch.ones(n, n, dtype=torch.bool, device="cuda"))
    assert float((got[low] - want[low]).norm() / want[low].norm()) < 2e-6
    # rows of the band just above the diagonal inside computed tiles are products too, never garbage
    assert torch.isfinite(got).all()


def test_laplacian_kernel_any_row_count(ops):
    """the laplacian initial kernel used to reject n > 65 535 (rows on gridDim.y)."""
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(70001, 37, device="cuda", generator=g)
    r = x[::997].contiguous()
    got = ops.initial_kernel("laplacian", x, r, gamma=0.05)
    want = torch.exp(-0.05 * torch.cdist(x.double(), r.double(), p=1))
    assert got.shape == (70001, r.shape[0]) and rel(got, want) < 1e-5
