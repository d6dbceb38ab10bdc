"""The Nyström root ``_sqrt_Nystrom`` (reference ``src/compute.py:6-17``) WITHOUT a full eigendecomposition.

The reference computes ``D, V = eigh(M)`` of the landmark block ``M = K[mask]`` and returns
``Q[mask] = V sqrt(D+)``, ``Q[~mask] = K[~mask] V diag(sqrt(D+) / max(D, tau))`` with ``tau = 1e-4 D_max``.  Every
consumer of ``Q`` sees only ``Q Qᵀ`` (the next layer's Gram, the fit's ``Q_bᵀ Q_b``), so any ``Q R`` with an orthogonal
``R`` is the same factor.  Write ``f(lam) = lam / max(lam, tau)^2`` (so that ``Q[~mask] Q[~mask]ᵀ = K f(M) Kᵀ``); below
the clamp ``f`` is LINEAR, ``f(lam) = lam / tau^2``, hence

    f(M) = M / tau^2 + V_t diag(1/lam - lam/tau^2) V_tᵀ            (V_t, lam: the eigenpairs ABOVE tau only)

and with the Cholesky factor ``M = L Lᵀ``, ``Z = Lᵀ V_t diag(lam)^-1/2`` (orthonormal columns) and ANY orthogonal
``Q_H = [Z | Z_perp]`` (here: the Householder QR of ``Z``) the matrices

    B = L Q_H          (B Bᵀ = M;  columns j < k are ± V_t sqrt(lam_j))
    S = B diag(s),     s_j = 1/lam_j (j < k),  1/tau (j >= k)        satisfy  S Sᵀ = f(M):

``Q[mask] = B`` and ``Q[~mask] = K[~mask] S`` is the reference's factor times one orthogonal matrix (``R = S_ref⁻¹ S``).
The leading k columns are EXACTLY the reference's columns for the eigenvalues above the clamp (graded like ``sqrt(D)`` /
``1/sqrt(D)``), the other n-k columns an orthogonal mix of the directions below it (whose scales differ by < 4x).  The
grading matters downstream: the fit's Gram matrix ``Q_bᵀQ_b`` is formed in fp32-equivalent arithmetic, whose elementwise
relative error preserves the small eigenvalues only when the columns carry their own scales — an ungraded ``S`` (e.g.
``S = L/tau + V_t diag(1/lam - 1/tau) V_tᵀ L``, tried first) made ``Q_bᵀQ_b + eps d I`` numerically indefinite at the
small nuggets.
On the benchmark graph 647 of 15 404 eigenvalues lie above the clamp: instead of a 15 404² eigendecomposition
(cuSOLVER fp64: 2.28 s on B200, bound by the one-stage tridiagonalisation) the route needs

  * one dense fp64 Cholesky factorisation (n³/3 flops, blocked, own DMMA GEMM),
  * the eigenpairs above ``tau`` from a block Lanczos process with full re-orthogonalisation (block 128, Krylov
    dimension ~0.11 n; every product ``M V_j`` is one fp64 GEMM) and a Rayleigh–Ritz step on the small projected matrix,
  * one skinny GEMM for ``Z``, a Householder QR of the n x k block ``Z`` and its block reflectors applied to ``L``.

Everything is fp64: the bulk columns carry the scale 1/tau against 1/lam_max = 1e-4/tau for the top direction, so they must
be orthogonal to the top Ritz directions to ~1e-10 (a leak of 1e-6 would already double the top entry of ``S Sᵀ``), which
needs the top Ritz vectors that accurate; a Ritz pair (theta, y) is accepted when ``|M y - theta y| <= tol * tau * theta_max /
theta`` (the residual bound under which the error of ``Q Qᵀ`` stays below ~1e-8, measured against the full
eigendecomposition on the benchmark's landmark block: 1.7e-10 at Krylov dimension 1 664).  When the Cholesky factorisation
meets a non-positive pivot (an indefinite or rank-deficient block, which the reference's clamp tolerates) or the Ritz
pairs above the clamp do not converge within the dimension budget (a spectrum with most of its eigenvalues above the
clamp, e.g. fraction 50), the caller takes the reference's eigen route instead.

This module is host logic only; the operators (``sym_f64``, ``dgemm``, ``potrf_f64``, ``potrf_inv_block``, ...) are
CUDA kernels behind the C ABI (``csrc/dense64.cu``); the CPU suite drives this logic through a test-only stand-in.
"""
from __future__ import annotations

import math
import os
from typing import Optional, Tuple

import torch

PANEL = 128                      # panel width of the Householder QR kernels (csrc/bandfit.cu)
DEFAULT_BLOCK = 128              # Lanczos block = the Cholesky panel width of csrc/dense64.cu
CLAMP = 1e-4                     # src/compute.py:13
RESIDUAL_TOL = 1e-7              # see the module docstring
# Krylov dimensions (fractions of n) at which a Rayleigh-Ritz check is made; past the last one the route gives up
CHECKPOINTS = (0.11, 0.16, 0.23)


def wanted(n: int, ops) -> bool:
    """Route policy: ``GNNGP_ROOT_SOLVER`` = eigh | krylov | auto (default).  auto takes the Krylov route for blocks of
    at least 4 096 landmarks (below that cuSOLVER's eigh costs less than the Cholesky factorisation plus the Lanczos
    steps: 54 ms at n = 3 052) when the operator set has the fp64 kernels."""
    mode = os.environ.get("GNNGP_ROOT_SOLVER", "auto")
    if mode == "eigh" or not hasattr(ops, "dgemm"):
        return False
    if mode == "krylov":
        return n >= 2 * int(os.environ.get("GNNGP_ROOT_BLOCK", DEFAULT_BLOCK))
    return n >= 4096


def _checkpoint_blocks(n: int, block: int):
    out = []
    for frac in CHECKPOINTS:
        blocks = int(math.ceil(frac * n / block))
        blocks = max(2, min(blocks, n // block - 1))
        if blocks * block < n and (not out or blocks > out[-1]):
            out.append(blocks)
    return out


class _Phases(object):
    """Per-phase wall times of one call (device-synchronised), only when ``stats["time"]`` is set (tools/root_probe.py)."""

    def __init__(self, stats, dev):
        self.on = bool(stats is not None and stats.get("time")) and dev.type == "cuda"
        self.stats, self.dev, self.t0 = stats, dev, None
        if self.on:
            import time
            self.clock = time.perf_counter
            torch.cuda.synchronize(dev)
            self.t0 = self.clock()
            stats["phase_ms"] = {}

    def mark(self, name):
        if self.on:
            torch.cuda.synchronize(self.dev)
            now = self.clock()
            self.stats["phase_ms"][name] = self.stats["phase_ms"].get(name, 0.0) + 1e3 * (now - self.t0)
            self.t0 = now


def _orthonormalise(ops, w: torch.Tensor, out: torch.Tensor, info: torch.Tensor, gram: torch.Tensor, tmp: torch.Tensor) -> None:
    """``out = w R⁻¹`` with orthonormal columns by two rounds of Cholesky-QR (``wᵀw = R_1ᵀR_1``, ...): three GEMMs and
    one 128 x 128 factorisation + triangular inverse per round; a breakdown (rank-deficient block) lands in ``info``."""
    ops.dgemm(w.T, w, out=gram)
    linv = ops.potrf_inv_block(gram, info)
    ops.dgemm(w, linv.T, out=tmp)
    ops.dgemm(tmp.T, tmp, out=gram)
    linv = ops.potrf_inv_block(gram, info)
    ops.dgemm(tmp, linv.T, out=out)


class RowSplit(object):
    """The replicated Krylov routes in sharded mode: products ``M @ V`` (and the rotation of the Cholesky factor) are
    independent across rows of ``M``, so rank r computes rows [r * per, (r + 1) * per) and one all-gather of the equal-sized
    row blocks hands every rank the whole result — the matrices themselves stay replicated (they come out of
    all-reduces), only the GEMM work is divided.  ``comm`` needs ``world``, ``rank`` and ``all_gather_equal``; without it
    (single GPU) every call is the plain local product."""

    def __init__(self, ops, n: int, comm=None):
        self.ops, self.n = ops, n
        self.comm = comm if (comm is not None and getattr(comm, "world", 1) > 1 and hasattr(comm, "all_gather_equal")) else None
        if self.comm is not None:
            world = self.comm.world
            self.per = -(-n // world)
            self.per += self.per & 1                                  # even: keeps the row blocks 16-byte aligned
            self.r0 = min(n, self.comm.rank * self.per)
            self.r1 = min(n, self.r0 + self.per)
            self._bufs = {}

    def _buffers(self, cols: int, like: torch.Tensor):
        hit = self._bufs.get(cols)
        if hit is None:
            piece = torch.zeros((self.per, cols), dtype=like.dtype, device=like.device)
            full = torch.empty((self.comm.world * self.per, cols), dtype=like.dtype, device=like.device)
            hit = self._bufs[cols] = (piece, full)
        return hit

    def matmul(self, m: torch.Tensor, v: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """``out = m @ v`` ([n x n] @ [n x c])."""
        if self.comm is None:
            return self.ops.dgemm(m, v, out=out)
        piece, full = self._buffers(v.shape[1], m)
        if self.r1 > self.r0:
            self.ops.dgemm(m[self.r0:self.r1], v, out=piece[:self.r1 - self.r0])
        self.comm.all_gather_equal(piece, full)
        out.copy_(full[:self.n])
        return out

    def rotate(self, x: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
        """``x <- x Q_H`` (``ops.qr_rotate``): the Householder QR of ``z`` is replicated, the reflectors are applied to
        this rank's rows of ``x`` only."""
        if self.comm is None or x.shape[0] != self.n:
            return self.ops.qr_rotate(x, z)
        piece, full = self._buffers(x.shape[1], x)
        rows = self.r1 - self.r0
        if rows > 0:
            piece[:rows].copy_(x[self.r0:self.r1])
            self.ops.qr_rotate(piece[:rows], z)
        self.comm.all_gather_equal(piece, full)
        x.copy_(full[:self.n])
        return x


_SIDE_STREAMS = {}


class _CholeskyJob(object):
    """The Cholesky factor of M, needed only by the assembly.  It is launched on a side stream just before the
    Rayleigh-Ritz step so that its GEMMs fill the GPU while the small projected eigenproblem (a chain of latency-bound
    library kernels, ~27 ms at Krylov dimension 1 792) runs on the main stream; ``get`` joins the streams and returns the
    factor, or ``None`` after a non-positive pivot."""

    def __init__(self, ops, m64):
        self.ops, self.m64, self.chol, self.info, self.side = ops, m64, None, None, None

    def _copy(self):
        n = self.m64.shape[0]
        out = torch.empty((n, n + (n & 1)), dtype=self.m64.dtype, device=self.m64.device)[:, :n]   # even pitch, like M
        out.copy_(self.m64)
        return out

    def start(self):
        if self.chol is not None:
            return
        dev = self.m64.device
        self.info = torch.zeros(1, dtype=torch.int32, device=dev)
        if dev.type != "cuda" or os.environ.get("GNNGP_ROOT_OVERLAP", "1") == "0":
            self.chol = self._copy()
            self.ops.potrf_f64(self.chol, self.info)
            return
        main = torch.cuda.current_stream(dev)
        self.side = _SIDE_STREAMS.get(dev.index)
        if self.side is None:
            self.side = _SIDE_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            self.chol = self._copy()
            self.ops.potrf_f64(self.chol, self.info)
        self.chol.record_stream(main)
        self.info.record_stream(self.side)
        self.m64.record_stream(self.side)                            # (the caller may drop M before the side stream is done)

    def get(self):
        self.start()
        if self.side is not None:
            torch.cuda.current_stream(self.m64.device).wait_stream(self.side)
            self.side = None
        return self.chol if int(self.info.item()) == 0 else None


def root_factors(ops, kmm: torch.Tensor, stats: Optional[dict] = None, block: Optional[int] = None,
                 comm=None) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
    """(``Sᵀ`` as the fp32 B operand of ``E @ S`` — the role of ``(V D~)ᵀ`` in ``Pipeline._root_factors`` — and the
    fp32 landmark rows ``L``), or ``None`` when the route does not apply to this block (the caller then runs ``eigh``).
    ``kmm``: fp32 [n x n] landmark block, lower triangle read.  ``block`` (<= 128) is the Lanczos block size.  ``comm``
    (sharded mode, the block replicated on every rank): the GEMM work of the products and of the rotation is divided
    over the ranks (``RowSplit``)."""
    n = kmm.shape[0]
    dev = kmm.device
    BLOCK = int(block if block is not None else os.environ.get("GNNGP_ROOT_BLOCK", DEFAULT_BLOCK))   # <= 128 (tests: 32)
    checkpoints = _checkpoint_blocks(n, BLOCK)
    if not checkpoints:
        return None
    f64 = torch.float64
    phases = _Phases(stats, dev)
    m64 = ops.sym_f64(kmm)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    chol = _CholeskyJob(ops, m64)                                    # launched at the first Rayleigh-Ritz step
    split = RowSplit(ops, n, comm)

    # ---- block Lanczos with full re-orthogonalisation
    dim_max = (checkpoints[-1] + 1) * BLOCK
    basis = torch.empty((n, dim_max), dtype=f64, device=dev)        # K = [V_0 V_1 ...]
    mbasis = torch.empty((n, dim_max), dtype=f64, device=dev)       # M K
    work = torch.empty((n, BLOCK), dtype=f64, device=dev)
    tmp = torch.empty((n, BLOCK), dtype=f64, device=dev)
    gram = torch.empty((BLOCK, BLOCK), dtype=f64, device=dev)
    coef = torch.empty((dim_max, BLOCK), dtype=f64, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(20240229)
    work.copy_(torch.randn((n, BLOCK), dtype=f64, device=dev, generator=gen))
    _orthonormalise(ops, work, basis[:, :BLOCK], info, gram, tmp)

    result = None
    blocks_done = 0                                                  # blocks j with M V_j stored
    for target in checkpoints:
        while blocks_done < target:
            j = blocks_done
            vj = basis[:, j * BLOCK:(j + 1) * BLOCK]
            mv = mbasis[:, j * BLOCK:(j + 1) * BLOCK]
            split.matmul(m64, vj, mv)
            blocks_done += 1
            if blocks_done == target and target == checkpoints[-1]:
                break                                                # the last block is only needed for the projection
            have = basis[:, :(j + 1) * BLOCK]
            work.copy_(mv)
            for _ in range(2):                                       # classical Gram-Schmidt, twice
                c = coef[:(j + 1) * BLOCK]
                ops.dgemm(have.T, work, out=c)
                ops.dgemm(have, c, out=work, alpha=-1.0, beta=1.0)
            _orthonormalise(ops, work, basis[:, (j + 1) * BLOCK:(j + 2) * BLOCK], info, gram, tmp)
        dim = blocks_done * BLOCK
        phases.mark("lanczos")
        got = _rayleigh_ritz(ops, chol, basis[:, :dim], mbasis[:, :dim], info, stats, BLOCK, phases, checkpoints[-1] * BLOCK, split)
        if isinstance(got, str):                                     # "breakdown" / "hopeless"
            break
        if got is not None:
            result = got
            break
        # the basis needs block `blocks_done` next: it exists unless the last product was the final one
        if target == checkpoints[-1]:
            break
    if stats is not None:
        stats["krylov_dim"] = blocks_done * BLOCK
        stats["converged"] = result is not None
    return result


def _rayleigh_ritz(ops, chol, basis, mbasis, info, stats, block, phases, dim_budget, split):
    """Project, solve the small eigenproblem, test the Ritz pairs above the clamp, assemble ``Sᵀ``.  Returns the pair of
    factors, ``None`` (not converged at this dimension), ``"breakdown"`` (non-positive pivot somewhere) or ``"hopeless"``
    (too many eigenvalues above the clamp for the dimension budget)."""
    n, dim = basis.shape
    dev = basis.device
    f64 = torch.float64
    h = torch.empty((dim, dim), dtype=f64, device=dev)
    ops.dgemm(basis.T, mbasis, out=h)
    # the first host synchronisation of the route: breakdown flag of the Cholesky-QR steps
    if int(info.item()) != 0:
        return "breakdown"
    chol.start()                                                     # side stream: overlaps the small eigenproblem
    theta, y = ops.eigh_small_f64(h)                                 # ascending
    phases.mark("projected_eigh")
    theta_host = theta.cpu()
    top = float(theta_host[-1])
    if not (top > 0.0) or not bool(torch.isfinite(theta_host).all()):
        return "breakdown"
    tau = CLAMP * top
    k = int((theta_host > tau).sum())
    k_check = int((theta_host > 0.9 * tau).sum())                    # Ritz values just below the clamp are tested too
    if stats is not None:
        stats.setdefault("ritz_above_clamp", []).append(k)
    if k_check + block > dim:                                        # the Krylov space has not reached past the clamp yet
        # ... and will not within the budget when it is saturated with wanted Ritz values this early (Ritz counts are
        # lower bounds of the true count, and convergence needs a space ~2.5x the count): hand the block back now
        return "hopeless" if k > 0.3 * dim_budget else None
    sel = theta_host[dim - k_check:]
    y_sel = y[:, dim - k_check:]
    pitch = k_check + (k_check & 1)                                  # even row pitch: the fp64 GEMM's aligned fast path
    ritz = torch.empty((n, pitch), dtype=f64, device=dev)[:, :k_check]
    resid = torch.empty((n, pitch), dtype=f64, device=dev)[:, :k_check]
    ops.dgemm(basis, y_sel, out=ritz)
    ops.dgemm(mbasis, y_sel, out=resid)
    theta_dev = sel.to(dev)
    rnorm = torch.sqrt(ops.residual_sumsq(resid, ritz, theta_dev)).cpu()
    allowed = RESIDUAL_TOL * tau * top / sel
    allowed = torch.where(sel > tau, allowed, 100.0 * allowed)
    worst = float((rnorm / allowed).max()) if k_check else 0.0
    if stats is not None:
        stats.setdefault("residual_over_allowed", []).append(worst)
    phases.mark("ritz_vectors")
    if not worst <= 1.0:
        # not converged at this dimension.  Give up early when the previous test shows that the remaining checkpoints
        # cannot bring the worst residual under its bound at the observed rate (arxiv shape: 478 -> 315 -> 197 allowed
        # residuals over 1 024 -> 1 536 -> 2 176 dimensions: more than a thousand eigenvalues above the clamp)
        history = stats.get("residual_over_allowed", []) if stats is not None else []
        if len(history) >= 2 and history[-2] > 0.0:
            gain = history[-2] / max(worst, 1e-300)
            if gain <= 1.0 or worst / gain ** 2 > 1.0:
                return "hopeless"
        return None
    chol = chol.get()
    if chol is None:
        return "breakdown"
    phases.mark("cholesky_wait")
    # ---- rotate the Cholesky factor into the graded basis: B = L Q_H with Q_H = [Z | Z_perp], Z = L^T V_t Theta^-1/2
    vt = ritz[:, k_check - k:]                                       # the k Ritz vectors above the clamp
    theta_t = sel[k_check - k:]
    k_pad = -(-max(k, 1) // PANEL) * PANEL
    if k_pad > n:
        return None
    z = torch.zeros((n, k_pad), dtype=f64, device=dev)
    if k > 0:
        vs = torch.empty((n, k + (k & 1)), dtype=f64, device=dev)[:, :k]
        ops.scale_cols_f64(vt, torch.rsqrt(theta_t).to(dev), out=vs)
        ops.dgemm(chol.T, vs, out=z[:, :k])
    split.rotate(chol, z)                                            # chol <- L Q_H: columns j < k are +-V_t sqrt(theta_j)
    scale = torch.full((n,), 1.0 / tau, dtype=f64)
    scale[:k] = 1.0 / theta_t
    st = ops.transpose_scale_f64(ops.scale_cols_f64(chol, scale.to(dev)), 1.0)
    out = ops.to_f32(st), ops.to_f32(chol)
    phases.mark("assemble")
    return out
