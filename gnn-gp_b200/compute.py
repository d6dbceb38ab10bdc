"""Kernel recursion of GNNGP on B200 — same function surface as the reference's ``src/compute.py``.

``_init_kernel``, ``_init_kernel_Nystrom``, ``_get_kernel`` and ``_get_kernel_Nystrom`` keep the
reference's names, positional order, defaults and error messages (``src/compute.py:20,41,70,90``); the
bodies enqueue the hand-written CUDA kernels through :mod:`gnngp_b200.pipeline`.  Inputs are CUDA
tensors (``A`` a torch sparse COO tensor as built by ``datasets.get_data``); new tensors are returned
and inputs are never modified.  ``Q`` is defined up to an orthogonal rotation (like the reference's,
whose ``eigh`` fixes no sign): compare ``Q @ Q.T``.
"""
from __future__ import annotations

import torch
from torch import Tensor

from .ops import default_ops
from .pipeline import Landmarks, Pipeline, RowLayout


def _pipeline() -> Pipeline:
    return Pipeline(default_ops())


def _landmarks(mask: Tensor) -> Landmarks:
    return Landmarks(mask, RowLayout(mask.numel(), 1, 0))


def _f32(x: Tensor) -> Tensor:
    return x if x.dtype == torch.float32 else x.to(torch.float32)


def _sqrt_Nystrom(K: Tensor, mask: Tensor) -> Tensor:
    """Nystrom approx square root ``Ny(K)`` from the landmark columns ``K`` [N, N_a]
    (reference ``src/compute.py:6-17``)."""
    return _pipeline().nystrom_root(_f32(K), _landmarks(mask))


def _init_kernel(X: Tensor, initial: str, **params) -> Tensor:
    """Initial kernel of input features ``C0`` (reference ``src/compute.py:20-38``)."""
    return _pipeline().initial_kernel(_f32(X), initial, params)


def _init_kernel_Nystrom(X: Tensor, mask: Tensor, initial: str, **params) -> Tensor:
    """Nystrom approx square root ``Q0`` of the initial kernel (reference ``src/compute.py:41-67``)."""
    return _pipeline().initial_factor(_f32(X), _landmarks(mask), initial, params)


def _get_kernel(C0: Tensor, A: Tensor, L: int, sigma_b: float, sigma_w: float, method: str, **params) -> Tensor:
    """GNNGP kernel ``K`` (reference ``src/compute.py:70-87``)."""
    pipe = _pipeline()
    return pipe.kernel_exact(_f32(C0), pipe.ops.csr_of(A), L, sigma_b, sigma_w, method, params)


def _get_kernel_Nystrom(Q0: Tensor, A: Tensor, L: int, sigma_b: float, sigma_w: float, mask: Tensor, method: str,
                        **params) -> Tensor:
    """Nystrom approx square root ``Q`` of the GNNGP kernel (reference ``src/compute.py:90-107``).
    ``_product_cache`` (not a kernel parameter; passed by ``GNNGP`` when ``reuse_products`` is on) lets a
    hyper-parameter sweep keep the first propagation ``A @ Q0``."""
    pipe = _pipeline()
    cache = params.pop("_product_cache", None)
    if cache is not None:
        pipe.product_cache = cache
    return pipe.kernel_nystrom(_f32(Q0), pipe.ops.csr_of(A), L, sigma_b, sigma_w, _landmarks(mask), method, params)


def _ExxT_ReLU(K: Tensor) -> Tensor:
    """``g(K)`` for ReLU (reference ``src/compute.py:110-117``)."""
    return default_ops().arccos_dense(_f32(K))


def _ExxT_ReLU_Nystrom(Q: Tensor, mask: Tensor) -> Tensor:
    """``Ny(g(QQ^T))`` for ReLU (reference ``src/compute.py:120-127``)."""
    return _pipeline().relu_nystrom(_f32(Q), _landmarks(mask))
