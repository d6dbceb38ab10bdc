"""TEST-ONLY stand-in for ``gnngp_b200.ops.CudaOps`` on CPU tensors.

It lets the CPU suite exercise the HOST logic of the product — the assembly of the layer recursions in
``gnngp_b200/pipeline.py`` and the row-sharded orchestration in ``gnngp_b200/shard.py`` (gloo,
world_size 2) — without a GPU.  Each method restates the contract of the CUDA operator in plain torch.
The product never imports this module; ``gnngp_b200.ops.default_ops()`` raises without the CUDA library.
"""
from __future__ import annotations

import math

import torch


class _Csr(object):
    def __init__(self, mat, n_rows, n_cols, row_offset=0):
        self.mat, self.n_rows, self.n_cols, self.row_offset = mat, n_rows, n_cols, row_offset
        self.nnz = mat._nnz()


class StubOps(object):
    name = "cpu-stub"

    def empty(self, n, m, device=None):
        return torch.zeros((n, (max(m, 1) + 3) // 4 * 4))[:, :m]

    zeros = empty

    def launch_count(self):
        return 0

    def csr_from_coo(self, row, col, val, n_rows, n_cols, row_offset=0):
        mat = torch.sparse_coo_tensor(torch.stack([row, col]), val.float(), (n_rows, n_cols)).coalesce()
        return _Csr(mat, n_rows, n_cols, row_offset)

    def csr_of(self, A):
        if isinstance(A, _Csr):
            return A
        idx = A._indices()
        return self.csr_from_coo(idx[0], idx[1], A._values(), A.shape[0], A.shape[1])

    def restrict_columns(self, csr, cols):
        dense_cols = csr.mat.index_select(1, cols).coalesce()
        return _Csr(dense_cols, csr.n_rows, int(cols.numel()), csr.row_offset)

    def spmm(self, csr, B, alpha=1.0, out=None):
        res = alpha * torch.sparse.mm(csr.mat, B.contiguous())
        if out is None:
            return res
        out.copy_(res)
        return out

    def gemm_nt(self, A, B, alpha=1.0, out=None, beta=0.0):
        res = alpha * (A @ B.T)
        if out is None:
            return res
        out.copy_(res if beta == 0.0 else res + beta * out)
        return out

    def syrk(self, A):
        return torch.tril(A @ A.T)          # like the CUDA operator: only the lower triangle is meaningful

    def gram_arccos(self, Q, L, s, t, out=None):
        c = torch.clamp(Q @ L.T / s.view(-1, 1) / t.view(1, -1), -1, 1)
        theta = torch.arccos(c)
        res = 0.5 / math.pi * (torch.sin(theta) + (math.pi - theta) * torch.cos(theta)) * s.view(-1, 1) * t.view(1, -1)
        if out is None:
            return res
        out.copy_(res)
        return out

    def initial_kernel(self, kind, X, R, gamma=0.0, c=0.0, d=1.0):
        if kind == "linear":
            return X @ R.T
        if kind == "rbf":
            return torch.exp(-gamma * torch.cdist(X, R, p=2) ** 2)
        if kind == "laplacian":
            return torch.exp(-gamma * torch.cdist(X, R, p=1))
        if kind == "sigmoid":
            return torch.tanh(gamma * X @ R.T + c)
        if kind == "polynomial":
            return (X @ R.T + c) ** d
        if kind == "arccos":
            return 1 - torch.arccos(torch.clamp(X @ R.T, -1, 1)) / math.pi
        raise ValueError(kind)

    def nugget_coeff(self, temp, w, eps, d):
        G, (m, C) = eps.numel(), temp.shape
        S = temp.T.unsqueeze(0) / (w.view(1, 1, m) + eps.view(G, 1, 1) * d.view(1, 1, 1))      # [G, C, m]
        return S.reshape(G * C, m)

    def nugget_sweep(self, Q, Bt, G, C, out=None):
        prod = Q @ Bt.T                                                                      # [n, G*C]
        return prod.reshape(Q.shape[0], G, C).permute(1, 0, 2).contiguous()

    def row_norms(self, Q, sqrt=True):
        ss = (Q * Q).sum(1)
        return torch.sqrt(ss) if sqrt else ss

    def mean(self, v, stride=1, count=None):
        return v.flatten().mean().reshape(1)

    def diag_mean(self, K):
        return torch.diag(K).mean().reshape(1)

    def fill_col(self, M, col, value):
        M[:, col] = value

    def gather_rows(self, src, idx):
        return src[idx].clone()

    def scatter_rows(self, dst, idx, src):
        dst[idx] = src

    def gather_rows_t(self, src, idx):
        return (src if idx is None else src[idx]).T.contiguous()

    def transpose(self, src):
        return src.T.contiguous()

    def gather_cols(self, src, idx):
        return src[:, idx].contiguous()

    def affine(self, out, x, y, a, b, c):
        res = a * x + c
        if y is not None:
            res = res + b * y
        out.copy_(res)
        return out

    def eigh(self, M):
        return torch.linalg.eigh(M.contiguous())

    # ---- fp64 dense operators of the Krylov Nystrom root (csrc/dense64.cu)
    def sym_f64(self, kmm):
        low = torch.tril(kmm).double()
        return low + torch.tril(low, -1).T

    def dgemm(self, A, B, out, alpha=1.0, beta=0.0, lower=False):
        res = alpha * (A @ B)
        if beta != 0.0:
            res = res + beta * out
        if lower:
            res = torch.where(torch.ones_like(res, dtype=torch.bool).tril(), res, out)
        out.copy_(res)
        return out

    def potrf_f64(self, A, info):
        factor, bad = torch.linalg.cholesky_ex(A)
        info += (bad != 0).to(torch.int32)
        A.copy_(torch.tril(factor))
        return A

    def potrf_inv_block(self, G, info):
        factor, bad = torch.linalg.cholesky_ex(G)
        info += (bad != 0).to(torch.int32)
        G.copy_(torch.tril(factor))
        if int(bad) != 0:
            return torch.eye(G.shape[0], dtype=G.dtype)
        return torch.linalg.solve_triangular(G, torch.eye(G.shape[0], dtype=G.dtype), upper=False)

    def qr_rotate(self, X, Z):
        q, _ = torch.linalg.qr(Z, mode="complete")
        X.copy_(X @ q)
        return X

    def band_solve(self, H, rhs_t, shifts):
        dim, c = H.shape[0], rhs_t.shape[0]
        low = torch.tril(H)
        sym = low + torch.tril(low, -1).T
        i = torch.arange(dim)
        sym = torch.where((i.view(-1, 1) - i.view(1, -1)).abs() <= 128, sym, torch.zeros_like(sym))
        out = torch.empty((int(shifts.numel()) * c, dim), dtype=torch.float64)
        bad = torch.zeros(1, dtype=torch.int32)
        for g in range(int(shifts.numel())):
            factor, info = torch.linalg.cholesky_ex(sym + shifts[g] * torch.eye(dim, dtype=torch.float64))
            bad += (info != 0).to(torch.int32)
            out[g * c:(g + 1) * c] = torch.cholesky_solve(rhs_t.T.contiguous(), factor).T
        return out, bad

    def eigh_small_f64(self, H):
        return torch.linalg.eigh(H)

    def residual_sumsq(self, R, V, theta):
        return ((R - V * theta.view(1, -1)) ** 2).sum(0)

    def transpose_scale_f64(self, A, alpha):
        return (alpha * A.T).contiguous()

    def scale_cols_f64(self, V, d, out=None):
        res = V * d.view(1, -1)
        if out is None:
            return res
        out.copy_(res)
        return out

    def to_f32(self, A):
        out = self.empty(A.shape[0], A.shape[1])
        out.copy_(A)
        return out

    def shifted_solves(self, gram, rhs_t, shifts):
        m, c = gram.shape[0], rhs_t.shape[0]
        sym = torch.tril(gram).double()
        sym = sym + torch.tril(sym, -1).T
        out = self.empty(int(shifts.numel()) * c, m)
        bad = torch.zeros((), dtype=torch.int64, device=gram.device)
        for i in range(int(shifts.numel())):
            factor, info = torch.linalg.cholesky_ex(sym + shifts[i] * torch.eye(m, dtype=torch.float64, device=sym.device))
            bad += (info != 0).to(torch.int64).reshape(())
            out[i * c:(i + 1) * c] = torch.cholesky_solve(rhs_t.double().T.contiguous(), factor).T.float()
        return out, bad

    def band_fit(self, gram, rhs_t, shifts):
        """Stand-in of csrc/bandfit.cu: the same shifted solves through dense fp64 Cholesky factorisations."""
        out, bad = self.shifted_solves(gram, rhs_t, shifts)
        return out, bad.to(torch.int32).reshape(1)

    def nystrom_spectrum(self, D):
        root = torch.sqrt(torch.clamp(D, 0))
        return root, root / torch.clamp(D, D[-1] * 1e-4)

    def scale_cols(self, V, d, transpose=False):
        res = V * d.view(1, -1)
        return res.T.contiguous() if transpose else res

    def arccos_dense(self, K, out=None):
        s = torch.sqrt(torch.diag(K))
        return self.gram_arccos_from(K, s, out)

    def gram_arccos_from(self, K, s, out):
        c = torch.clamp(K / s.view(-1, 1) / s.view(1, -1), -1, 1)
        theta = torch.arccos(c)
        res = 0.5 / math.pi * (torch.sin(theta) + (math.pi - theta) * torch.cos(theta)) * s.view(-1, 1) * s.view(1, -1)
        if out is None:
            return res
        out.copy_(res)
        return out

    def onehot_t(self, y, idx, C):
        return torch.nn.functional.one_hot(y[idx], C).float().T.contiguous()

    def argmax_count(self, fitted, y, membership):
        G = fitted.shape[0]
        counts = torch.zeros((G, 8), dtype=torch.int32)
        hit = torch.argmax(fitted, 2) == y.view(1, -1)
        for s in range(8):
            sel = ((membership >> s) & 1).bool()
            counts[:, s] = hit[:, sel].sum(1)
        return counts

    def sqerr_sum(self, fitted, y, membership):
        G, n = fitted.shape[0], fitted.shape[1]
        err = ((fitted.reshape(G, n, -1) - y.reshape(1, n, -1)) ** 2).double().sum(2)
        sse = torch.zeros((G, 8), dtype=torch.float64)
        for s in range(8):
            sel = ((membership >> s) & 1).bool()
            sse[:, s] = err[:, sel].sum(1)
        return sse
