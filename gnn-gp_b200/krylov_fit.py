"""The GP posterior solve for ALL nuggets by a block Lanczos (block conjugate-gradient) process.

``fit_Nystrom`` (reference ``src/predict.py:57-62``) needs ``X_g = (G + eps_g d I)⁻¹ B`` for the 101 nuggets, with
``G = Q_bᵀQ_b`` (m x m) and ``B = Q_bᵀ y_b`` (m x C, C = number of classes).  Krylov spaces are SHIFT-INVARIANT:
``K_j(G, B) = K_j(G + s I, B)``, so ONE block Lanczos process started from ``B`` — one product ``G V_j`` (m x C x m) per
step, full re-orthogonalisation — serves every nugget: with the orthonormal basis ``K`` and ``H = Kᵀ G K`` (block
tridiagonal, half bandwidth 2C-1), the Galerkin solutions are ``X_g = K (H + s_g I)⁻¹ Kᵀ B``, and the small banded systems
are solved for all nuggets at once by the band fit's banded Cholesky kernel (one CTA per nugget).  The smallest nugget is
the slowest to converge; its TRUE residual ``|(G + s_0 I) X_0 - B| / |B|`` (one more product) is the stopping test.

The factor ``Q`` of the Nyström recursion has a few hundred directions that carry the signal and a large bulk of small,
nearly equal eigenvalues (Reddit shape, fraction 10: 650 of 15 405 eigenvalues of ``G`` above 1e-7 of the largest), which
is the spectrum block CG likes: the outliers are captured in the first steps, the shifted bulk is well conditioned
(numpy prototype on that matrix: residual of the smallest nugget 4.8e-7 after 20 steps, 2.8e-12 after 30).
When the residual test is not met within the step budget the caller takes the band-reduction route (``csrc/bandfit.cu``),
which is direct.  Host logic only; the operators are the fp64 kernels of ``csrc/dense64.cu`` / ``csrc/bandfit.cu``.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch

from .krylov_root import RowSplit, _orthonormalise

# Relative residual of the smallest nugget's solution at which the process stops.  The shifted systems are
# ill-conditioned (condition ~1e4 after the outliers are captured): on the benchmark's Gram matrix a residual of 5e-7
# still left a solution error of 5e-3, 3e-12 gave 4e-8 — hence 1e-10, reached after ~28 steps of 41 columns there.
RESIDUAL_TOL = 1e-10
FIRST_CHECK = 20                  # Lanczos steps before the first residual test
DECADES_PER_STEP = 0.35           # assumed convergence rate when scheduling the next test (measured: ~0.5)
MAX_STEPS = 72                    # ... and the Krylov dimension never passes MAX_DIM_FRACTION of m
MAX_DIM_FRACTION = 0.45


def wanted(m: int, n_targets: int, ops) -> bool:
    """``GNNGP_FIT_SOLVER=krylov`` forces the route (for blocks of at least 4 target columns); ``auto`` takes it for
    m >= 2 048 with 8..48 target columns (classification)."""
    mode = os.environ.get("GNNGP_FIT_SOLVER", "auto")
    if not hasattr(ops, "band_solve") or n_targets > 48:
        return False
    if mode == "krylov":
        return n_targets >= 4 and m >= 16 * n_targets
    return mode == "auto" and m >= 2048 and n_targets >= 8


def coefficients(ops, gram: torch.Tensor, qty_t: torch.Tensor, shifts: torch.Tensor,
                 stats: Optional[dict] = None, comm=None) -> Optional[torch.Tensor]:
    """The solutions in FACTORED form ``(y_t, basis_t)``: fp32 ``y_t`` [(G*C) x dim] and ``basis_t`` [dim x m] with
    ``((gram + shifts[g] I)⁻¹ rhs)[:, c] = basis_tᵀ y_t[g*C + c]`` — so ``fitted = (Q basis_tᵀ) y_tᵀ`` — or ``None`` (no
    convergence / breakdown: use the direct route).  ``gram``: fp32 [m x m], lower triangle read;
    ``qty_t``: fp32 [C x m]; ``shifts``: fp64 [G] on the device."""
    m, c = gram.shape[0], qty_t.shape[0]
    g = shifts.numel()
    dev = gram.device
    f64 = torch.float64
    g64 = ops.sym_f64(gram)
    split = RowSplit(ops, m, comm)                                   # sharded mode: the products' rows are divided over the ranks
    b = qty_t.to(f64).T.contiguous()                                 # [m x C]
    b_norm = torch.linalg.norm(b)
    max_steps = max(FIRST_CHECK, min(MAX_STEPS, int(MAX_DIM_FRACTION * m) // c))
    dim_max = (max_steps + 1) * c
    dim_max += dim_max & 1                                           # even pitch: the fp64 GEMM's aligned fast path
    basis = torch.empty((m, dim_max), dtype=f64, device=dev)
    gbasis = torch.empty((m, dim_max), dtype=f64, device=dev)
    cp = c + (c & 1)                                                 # even pitches everywhere (aligned GEMM fast path)
    work = torch.empty((m, cp), dtype=f64, device=dev)[:, :c]
    tmp = torch.empty((m, cp), dtype=f64, device=dev)[:, :c]
    gram_c = torch.empty((c, cp), dtype=f64, device=dev)[:, :c]
    coef = torch.empty((dim_max, cp), dtype=f64, device=dev)[:, :c]
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    work.copy_(b)
    _orthonormalise(ops, work, basis[:, :c], info, gram_c, tmp)
    rhs_k = torch.zeros((c, dim_max), dtype=f64, device=dev)        # (K^T B)^T: only the first block is non-zero
    ops.dgemm(b.T, basis[:, :c], out=rhs_k[:, :c])
    x0 = torch.empty((m, cp), dtype=f64, device=dev)[:, :c]
    i0 = int(torch.argmin(shifts))                                   # the smallest nugget converges last
    shift0 = shifts[i0:i0 + 1].to(f64)
    result = None
    last = None                                                      # (steps, residual) of the previous test
    next_check = min(FIRST_CHECK, max_steps)
    for j in range(max_steps):
        vj = basis[:, j * c:(j + 1) * c]
        gv = gbasis[:, j * c:(j + 1) * c]
        split.matmul(g64, vj, gv)
        steps = j + 1
        if steps == next_check or steps == max_steps:
            dim = steps * c
            if int(info.item()) != 0:
                break
            h = torch.empty((dim, dim), dtype=f64, device=dev)
            ops.dgemm(basis[:, :dim].T, gbasis[:, :dim], out=h)
            # the smallest nugget alone first (one CTA): its TRUE residual |(G + s0 I) X0 - B| / |B| decides
            y0, bad = ops.band_solve(h, rhs_k[:, :dim], shift0)     # [C x dim]
            ops.dgemm(basis[:, :dim], y0.T, out=x0)
            ops.dgemm(g64, x0, out=work)
            rel = float(torch.linalg.norm(work + shift0 * x0 - b) / b_norm)
            if stats is not None:
                stats.setdefault("residual", []).append(rel)
                stats.setdefault("checked_at", []).append(steps)
                stats["krylov_dim"] = dim
            if int(bad.item()) != 0 or not rel == rel:
                break
            if rel <= RESIDUAL_TOL:
                sol_t, bad = ops.band_solve(h, rhs_k[:, :dim], shifts)   # every nugget: [(G*C) x dim]
                if int(bad.item()) == 0:
                    # the coefficients stay FACTORED, X_g = K y_g: the caller contracts Q with K first (N x m x dim) and
                    # sweeps the nuggets over the Krylov dimension (dim ~ m / 13) instead of over m
                    result = ops.to_f32(sol_t), ops.to_f32(ops.transpose_scale_f64(basis[:, :dim], 1.0))
                break
            rate = DECADES_PER_STEP
            if last is not None and steps > last[0] and rel < last[1]:
                rate = min(1.0, math.log10(last[1] / rel) / (steps - last[0]))     # measured decades per step
            more = int(math.ceil(math.log10(rel / RESIDUAL_TOL) / rate))
            if last is not None and steps + more > max_steps:
                break                                                # will not get there within the budget: hand back now
            last = (steps, rel)
            next_check = min(max_steps, steps + max(2, min(more, 25)))
        have = basis[:, :steps * c]
        work.copy_(gv)
        for _ in range(2):
            cf = coef[:steps * c]
            ops.dgemm(have.T, work, out=cf)
            ops.dgemm(have, cf, out=work, alpha=-1.0, beta=1.0)
        _orthonormalise(ops, work, basis[:, steps * c:(steps + 1) * c], info, gram_c, tmp)
    if stats is not None:
        stats["converged"] = result is not None
    return result
