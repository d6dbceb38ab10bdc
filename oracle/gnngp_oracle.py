"""CPU oracle for the GNNGP hot path — TEST INFRASTRUCTURE ONLY.

A numpy / scipy restatement of the reference's kernel recursion and GP inference
(``/root/reference/src/compute.py`` and ``src/predict.py``).  It exists to check the CUDA path;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package never does, and it raises when its
CUDA library is missing instead of falling back to anything here.

Pinning: the reference publishes no tests or golden vectors for this path; the two
notebooks' numbers need the real Cora / chameleon downloads.  The oracle is therefore pinned against
outputs of the reference itself, imported from ``/root/reference/src`` in the build container by
``tests/golden/make_golden.py`` (committed fixtures under ``tests/golden/``), and live against the
imported reference in ``tests/test_oracle.py`` whenever ``/root/reference`` is present.

Every function cites the reference lines it follows.  Arithmetic is done in the dtype of the inputs
(float32 like the reference's default, or float64 for the 1e-10 mode).
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Dict, Sequence

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_PI = math.pi


# --------------------------------------------------------------------------------------------
# sparse operator
# --------------------------------------------------------------------------------------------
class Adjacency(object):
    """Normalised adjacency as CSR.  Built from the COO triplets the reference wraps in
    ``torch.sparse_coo_tensor`` (``src/datasets.py:328``); duplicates are summed, which is what
    torch's sparse ``@`` does when it coalesces its operand."""

    def __init__(self, rows, cols, vals, n: int):
        vals = np.asarray(vals)
        mat = sp.coo_matrix((vals, (np.asarray(rows), np.asarray(cols))), shape=(n, n)).tocsr()
        mat.sum_duplicates()
        mat.sort_indices()
        self.n = n
        self.indptr = mat.indptr.astype(np.int64)
        self.indices = mat.indices.astype(np.int32)
        self.data = mat.data
        self._scipy = mat

    def scaled(self, factor) -> "Adjacency":
        """``factor * A`` as a new sparse operator — the reference scales A's stored values first
        (operator precedence of ``sigma_w * A @ X``, ``src/compute.py:142``) and multiplies after."""
        out = Adjacency.__new__(Adjacency)
        out.n = self.n
        out.indptr = self.indptr
        out.indices = self.indices
        out.data = (self.data * self.data.dtype.type(factor)).astype(self.data.dtype)
        out._scipy = sp.csr_matrix((out.data, self.indices, self.indptr), shape=(self.n, self.n))
        return out

    def astype(self, dtype) -> "Adjacency":
        out = Adjacency.__new__(Adjacency)
        out.n = self.n
        out.indptr = self.indptr
        out.indices = self.indices
        out.data = self.data.astype(dtype)
        out._scipy = sp.csr_matrix((out.data, self.indices, self.indptr), shape=(self.n, self.n))
        return out

    def matmul(self, dense: np.ndarray) -> np.ndarray:
        """``A @ dense`` (dense may be a transposed view, as in ``A @ (A @ C0).T``)."""
        dense = np.ascontiguousarray(dense)
        lib = _native()
        if lib is not None and dense.dtype == np.float32 and self.data.dtype == np.float32:
            out = np.empty((self.n, dense.shape[1]), dtype=np.float32)
            lib.oracle_spmm_csr_f32(
                self.indptr.ctypes.data_as(ctypes.c_void_p), self.indices.ctypes.data_as(ctypes.c_void_p),
                self.data.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(self.n),
                dense.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(dense.shape[1]),
                out.ctypes.data_as(ctypes.c_void_p))
            return out
        return np.asarray(self._scipy @ dense)


_LIB = False


def _native():
    """Optional OpenMP CSR SpMM (``oracle/spmm_ref.c`` → ``oracle/_build/liboracle.so``) so the CPU
    baseline uses every host core; scipy's single-threaded product is used when it is not built."""
    global _LIB
    if _LIB is False:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        try:
            _LIB = ctypes.CDLL(path) if os.path.exists(path) else None
        except OSError:
            _LIB = None
    return _LIB


# --------------------------------------------------------------------------------------------
# Nystrom square root and ReLU arc-cosine expectation
# --------------------------------------------------------------------------------------------
def nystrom_root(k_cols: np.ndarray, land: np.ndarray) -> np.ndarray:
    """Clamped Nystrom factor of a kernel given by its landmark columns.

    Follows ``_sqrt_Nystrom`` (``src/compute.py:6-17``): eigendecompose the landmark block, landmark
    rows become ``V sqrt(max(D,0))``, all other rows ``K V sqrt(max(D,0)) / max(D, 1e-4 D_max)``."""
    dt = k_cols.dtype
    evals, evecs = np.linalg.eigh(k_cols[land])
    root = np.sqrt(np.maximum(evals, dt.type(0)))
    inv = root / np.maximum(evals, evals[-1] * dt.type(1e-4))
    out = np.zeros_like(k_cols)
    out[land] = evecs * root[None, :]
    rest = ~land
    out[rest] = (k_cols[rest] @ evecs) * inv[None, :]
    return out


def _j1(theta, dt):
    return dt.type(0.5 / _PI) * (np.sin(theta) + (dt.type(_PI) - theta) * np.cos(theta))


def relu_expectation(kernel: np.ndarray) -> np.ndarray:
    """``g(K)`` for a dense kernel — ``_ExxT_ReLU`` (``src/compute.py:110-117``)."""
    dt = kernel.dtype
    s = np.sqrt(np.diag(kernel)).reshape(-1, 1)
    theta = np.arccos(np.clip(kernel / s / s.T, dt.type(-1), dt.type(1)))
    return _j1(theta, dt) * s * s.T


def relu_expectation_nystrom(q: np.ndarray, land: np.ndarray) -> np.ndarray:
    """``Ny(g(Q Qᵀ))`` — ``_ExxT_ReLU_Nystrom`` (``src/compute.py:120-127``)."""
    dt = q.dtype
    s = np.sqrt(np.sum(q ** 2, axis=1)).reshape(-1, 1)
    sl = s[land].T
    theta = np.arccos(np.clip(q @ q[land].T / s / sl, dt.type(-1), dt.type(1)))
    return nystrom_root(_j1(theta, dt) * s * sl, land)


# --------------------------------------------------------------------------------------------
# initial kernels
# --------------------------------------------------------------------------------------------
def _sqdist(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Squared Euclidean distances the way ``torch.cdist(p=2)`` forms them for large inputs
    (norms + matmul, clamped at 0, sqrt) followed by the reference's ``**2``."""
    aa = np.sum(a * a, axis=1, keepdims=True)
    bb = np.sum(b * b, axis=1, keepdims=True).T
    d2 = np.maximum(aa + bb - a.dtype.type(2) * (a @ b.T), a.dtype.type(0))
    return np.sqrt(d2) ** 2


def _l1dist(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    out = np.empty((a.shape[0], b.shape[0]), dtype=a.dtype)
    for j in range(b.shape[0]):
        out[:, j] = np.sum(np.abs(a - b[j][None, :]), axis=1)
    return out


def _initial_columns(x: np.ndarray, ref_rows: np.ndarray, initial: str, params: dict) -> np.ndarray:
    """Kernel between all rows of ``x`` and ``ref_rows`` for the six initial kernels
    (``src/compute.py:24-35`` exact, ``46-64`` Nystrom)."""
    dt = x.dtype
    if initial == "linear":
        return x @ ref_rows.T
    if initial == "rbf":
        return np.exp(-dt.type(params["gamma"]) * _sqdist(x, ref_rows))
    if initial == "laplacian":
        return np.exp(-dt.type(params["gamma"]) * _l1dist(x, ref_rows))
    if initial == "sigmoid":
        return np.tanh((dt.type(params["gamma"]) * x) @ ref_rows.T + dt.type(params["c"]))
    if initial == "polynomial":
        return (x @ ref_rows.T + dt.type(params["c"])) ** dt.type(params["d"])
    raise Exception("Unsupported kernel function!")


def initial_kernel(x: np.ndarray, initial: str = "linear", **params) -> np.ndarray:
    """``C0`` — ``_init_kernel`` (``src/compute.py:20-38``)."""
    dt = x.dtype
    if initial == "arccos":
        corr = np.corrcoef(x).astype(dt)
        return dt.type(1) - np.arccos(corr) / dt.type(_PI)
    return _initial_columns(x, x, initial, params)


def initial_factor(x: np.ndarray, land: np.ndarray, initial: str = "linear", **params) -> np.ndarray:
    """``Q0`` — ``_init_kernel_Nystrom`` (``src/compute.py:41-67``)."""
    dt = x.dtype
    if initial == "linear":
        if int(land.sum()) >= x.shape[1]:
            return x
        return nystrom_root(x @ x[land].T, land)
    if initial == "arccos":
        z = x - np.mean(x, axis=0, keepdims=True)
        z = z / np.sqrt(np.sum(z ** 2, axis=0, keepdims=True))
        c0 = dt.type(1) - np.arccos(np.clip(z @ z[land].T, dt.type(-1), dt.type(1))) / dt.type(_PI)
        return nystrom_root(c0, land)
    return nystrom_root(_initial_columns(x, x[land], initial, params), land)


# --------------------------------------------------------------------------------------------
# layer recursions
# --------------------------------------------------------------------------------------------
def _sandwich(adj: Adjacency, factor, dense: np.ndarray) -> np.ndarray:
    """``factor * A @ (A @ dense).T`` with the reference's evaluation order: the scalar scales the
    sparse operand, the inner product is transposed before the outer one (``src/compute.py:132``)."""
    inner = adj.matmul(dense)
    return adj.scaled(factor).matmul(inner.T)


def kernel_exact(c0: np.ndarray, adj: Adjacency, layers: int, sigma_b: float, sigma_w: float,
                 method: str = "GCN", **params) -> np.ndarray:
    """Dense GNNGP kernel ``K`` — ``_get_kernel`` and the four ``_*_kernel`` bodies
    (``src/compute.py:70-87, 130-136, 149-157, 174-180, 193-199``)."""
    dt = c0.dtype
    adj = adj.astype(dt)
    b2 = dt.type(sigma_b ** 2)
    w2 = sigma_w ** 2
    if method == "GGP":
        return adj.matmul(adj.matmul(c0).T)
    if method == "RBF":
        return c0
    if method == "GCN":
        k = b2 + _sandwich(adj, w2, c0)
        for _ in range(layers - 1):
            k = b2 + _sandwich(adj, w2, relu_expectation(k))
        return k
    if method == "GIN":
        k = b2 + _sandwich(adj, w2, c0)
        for _ in range(layers - 1):
            k = b2 + dt.type(w2) * relu_expectation(k)
        return k
    if method == "SAGE":
        k = b2 * c0 + _sandwich(adj, w2, c0)
        for _ in range(layers - 1):
            e = relu_expectation(k)
            k = b2 * e + _sandwich(adj, w2, e)
        return k
    if method == "GCN2":
        alpha = params.get("alpha", 0.1)
        theta = params.get("theta", 0.5)
        k = _sandwich(adj, w2, relu_expectation(c0))
        for j in range(layers - 1):
            e = relu_expectation(k)
            beta = math.log(theta / (j + 1) + 1)
            coef = (sigma_w * beta) ** 2 + (1 - beta) ** 2
            # (1-alpha)**2 * A is formed first, then the product; alpha**2 * C0 is added; coef scales
            # the sum (src/compute.py:156)
            mix = _sandwich(adj, (1 - alpha) ** 2, e) + dt.type(alpha ** 2) * c0
            k = dt.type(coef) * mix
        return k
    raise Exception("Unsupported layer type!")


def kernel_nystrom(q0: np.ndarray, adj: Adjacency, layers: int, sigma_b: float, sigma_w: float,
                   land: np.ndarray, method: str = "GCN", **params) -> np.ndarray:
    """Nystrom factor ``Q`` of the GNNGP kernel — ``_get_kernel_Nystrom`` and the four
    ``_*_kernel_Nystrom`` bodies (``src/compute.py:90-107, 139-146, 160-171, 183-190, 202-210``)."""
    dt = q0.dtype
    adj = adj.astype(dt)
    n, ni = q0.shape
    na = int(land.sum())
    if method == "GGP":
        return adj.matmul(q0)
    if method == "RBF":
        return q0
    aw = adj.scaled(sigma_w)
    if method == "GCN":
        q = np.zeros((n, na + 1), dtype=dt)
        q[:, :ni] = aw.matmul(q0)
        q[:, -1] = sigma_b
        for _ in range(layers - 1):
            e = relu_expectation_nystrom(q, land)
            q[:, :-1] = aw.matmul(e)
            q[:, -1] = sigma_b
        return q
    if method == "GIN":
        q = np.zeros((n, na + 1), dtype=dt)
        q[:, :ni] = aw.matmul(q0)
        q[:, -1] = sigma_b
        for _ in range(layers - 1):
            e = relu_expectation_nystrom(q, land)
            q[:, :-1] = dt.type(sigma_w) * e
            q[:, -1] = sigma_b
        return q
    if method == "SAGE":
        q = np.zeros((n, 2 * na), dtype=dt)
        q[:, :ni] = dt.type(sigma_b) * q0
        q[:, ni:2 * ni] = aw.matmul(q0)
        for _ in range(layers - 1):
            e = relu_expectation_nystrom(q, land)
            q[:, :na] = dt.type(sigma_b) * e
            q[:, na:] = aw.matmul(e)
        return q
    if method == "GCN2":
        alpha = params.get("alpha", 0.1)
        theta = params.get("theta", 0.5)
        e = relu_expectation_nystrom(q0, land)
        q = np.zeros((n, na + ni), dtype=dt)
        q[:, :na] = aw.matmul(e)
        for j in range(layers - 1):
            e = relu_expectation_nystrom(q, land)
            beta = math.log(theta / (j + 1) + 1)
            coef = math.sqrt((sigma_w * beta) ** 2 + (1 - beta) ** 2)
            q[:, :na] = adj.scaled((1 - alpha) * coef).matmul(e)
            q[:, na:na + ni] = dt.type(alpha * coef) * q0
        return q
    raise Exception("Unsupported layer type!")


# --------------------------------------------------------------------------------------------
# GP inference
# --------------------------------------------------------------------------------------------
def _targets(y: np.ndarray, train: np.ndarray, dt) -> np.ndarray:
    """One-hot float32 targets for integer labels (``src/predict.py:23-25``; the reference one-hots
    the *training* labels, so the class count is max(train label)+1), raw values otherwise."""
    if np.issubdtype(y.dtype, np.integer):
        yt = y[train]
        out = np.zeros((yt.shape[0], int(yt.max()) + 1), dtype=np.float32)
        out[np.arange(yt.shape[0]), yt] = 1
        return out
    return y[train]


def _sweep(left: np.ndarray, evals, evecs, proj, scale, nugget: Sequence[float], dt):
    fitted = np.zeros((len(nugget), left.shape[0]) + proj.shape[1:], dtype=dt)
    for i, eps in enumerate(nugget):
        denom = evals + dt.type(eps) * scale
        coeff = proj / (denom.reshape(-1, 1) if proj.ndim >= 2 else denom)
        fitted[i] = left @ (evecs @ coeff)
    return fitted[0] if len(nugget) == 1 else fitted


def fit_exact(kernel: np.ndarray, y: np.ndarray, train: np.ndarray, nugget) -> np.ndarray:
    """Posterior means for every nugget from the dense kernel — ``fit`` (``src/predict.py:6-33``)."""
    dt = kernel.dtype
    nugget = np.atleast_1d(np.asarray(nugget, dtype=dt))
    yb = _targets(y, train, dt)
    evals, evecs = np.linalg.eigh(kernel[train][:, train])
    scale = np.mean(np.diag(kernel))
    proj = evecs.T @ yb
    return _sweep(kernel[:, train], evals, evecs, proj, scale, nugget, dt)


def fit_nystrom(q: np.ndarray, y: np.ndarray, train: np.ndarray, nugget) -> np.ndarray:
    """Posterior means from the Nystrom factor — ``fit_Nystrom`` (``src/predict.py:36-63``)."""
    dt = q.dtype
    nugget = np.atleast_1d(np.asarray(nugget, dtype=dt))
    yb = _targets(y, train, dt)
    qb = q[train]
    evals, evecs = np.linalg.eigh(qb.T @ qb)
    scale = np.mean(np.sum(q ** 2, axis=1))
    proj = evecs.T @ (qb.T @ yb)
    return _sweep(q, evals, evecs, proj, scale, nugget, dt)


def metrics(fitted: np.ndarray, y: np.ndarray, masks: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Accuracy (integer targets) or R² (float targets) per nugget and per mask — ``result``
    (``src/predict.py:66-90``).  Returned as float32 like the reference's ``torch.zeros``."""
    classify = np.issubdtype(y.dtype, np.integer)
    out = {name: np.zeros(fitted.shape[0], dtype=np.float32) for name in masks}
    for i in range(fitted.shape[0]):
        for name, mask in masks.items():
            f, t = fitted[i][mask], y[mask]
            if classify:
                out[name][i] = np.mean((np.argmax(f, axis=1) == t).astype(np.float32))
            else:
                out[name][i] = 1 - np.sum((f - t) ** 2) / np.sum((t - np.mean(t)) ** 2)
    return out


def summary(result: Dict[str, np.ndarray], nugget) -> dict:
    """Best-validation pick — ``GNNGP.get_summary`` (``src/gnngp.py:126-135``)."""
    i = int(np.argmax(result["val"]))
    return dict(index=i, nugget=float(np.asarray(nugget)[i]), train=float(result["train"][i]),
                val=float(result["val"][i]), test=float(result["test"][i]))


# --------------------------------------------------------------------------------------------
# whole path
# --------------------------------------------------------------------------------------------
def run(x, y, rows, cols, vals, masks: Dict[str, np.ndarray], nugget, layers=2, sigma_b=0.0, sigma_w=1.0,
        nystrom=False, initial="linear", method="GCN", dtype=np.float32, **params) -> dict:
    """``GNNGP.predict`` + ``get_summary`` (``src/gnngp.py:87-135``) on numpy inputs."""
    dt = np.dtype(dtype)
    x = np.asarray(x, dtype=dt)
    adj = Adjacency(rows, cols, np.asarray(vals, dtype=dt), x.shape[0])
    if np.issubdtype(np.asarray(y).dtype, np.floating):
        y = np.asarray(y, dtype=dt)
    out = {}
    if nystrom:
        q0 = initial_factor(x, masks["landmark"], initial, **params)
        q = kernel_nystrom(q0, adj, layers, sigma_b, sigma_w, masks["landmark"], method, **params)
        fitted = fit_nystrom(q, y, masks["train"], nugget)
        out["Q"] = q
    else:
        c0 = initial_kernel(x, initial, **params)
        k = kernel_exact(c0, adj, layers, sigma_b, sigma_w, method, **params)
        fitted = fit_exact(k, y, masks["train"], nugget)
        out["K"] = k
    out["fit"] = fitted
    out["result"] = metrics(fitted, y, masks)
    out["summary"] = summary(out["result"], nugget)
    return out
