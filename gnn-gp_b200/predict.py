"""GP posterior inference of GNNGP on B200 — same function surface as the reference's
``src/predict.py`` (``fit``, ``fit_Nystrom``, ``result``; ``src/predict.py:6,36,66``).

All nuggets are evaluated by ONE contraction ``[N x m] · [m x (G·C)]`` whose epilogue scatters into
the reference's ``[G, N, C]`` layout, instead of the reference's 101-iteration Python loop; metrics
for every (nugget, mask) pair come from one fused argmax / squared-error pass.
"""
from __future__ import annotations

from typing import Dict, Union

import torch
from torch import Tensor

from .ops import default_ops
from .pipeline import Pipeline


def _nugget(nugget, device) -> Tensor:
    if isinstance(nugget, float):
        nugget = torch.tensor([nugget])
    return nugget.to(device=device, dtype=torch.float32).contiguous()


def _classes(y: Tensor, rows: Tensor) -> int:
    # torch.nn.functional.one_hot(y[mb]) infers max(train label) + 1 classes (src/predict.py:24,54)
    return int(y[rows].max()) + 1


def fit(K: Tensor, y: Tensor, train_mask: Tensor, nugget: Union[float, Tensor] = 1e-2):
    """Predictions for a range of nugget values from the dense kernel (reference ``src/predict.py:6-33``)."""
    rows = torch.nonzero(train_mask, as_tuple=False).flatten()
    classes = None if y.dtype.is_floating_point else _classes(y, rows)
    return Pipeline(default_ops()).fit_exact(K, y, rows, _nugget(nugget, K.device), classes)


def fit_Nystrom(Q: Tensor, y: Tensor, train_mask: Tensor, nugget: Union[float, Tensor] = 1e-2):
    """Predictions for a range of nugget values from the Nystrom factor (reference ``src/predict.py:36-63``)."""
    rows = torch.nonzero(train_mask, as_tuple=False).flatten()
    classes = None if y.dtype.is_floating_point else _classes(y, rows)
    return Pipeline(default_ops()).fit_nystrom(Q, y, rows, _nugget(nugget, Q.device), None, classes)


def result(fit: Tensor, y: Tensor, masks: Dict[str, Tensor]):
    """Train / validation / test (/ landmark) metric for each nugget (reference ``src/predict.py:66-90``):
    mean accuracy for integer targets, R² for float targets; CPU fp32 tensors of length ``fit.shape[0]``."""
    return Pipeline(default_ops()).metrics(fit, y, masks)
