"""``GNNGP`` model class — drop-in for the reference's ``src/gnngp.py`` (same constructor, attributes
and methods; ``src/gnngp.py:63-135``), computing on B200 through :mod:`gnngp_b200.compute` and
:mod:`gnngp_b200.predict`."""
from __future__ import annotations

import os
from typing import Union

import torch
from torch import Tensor

from . import compute, datasets, predict


class GNNGP(object):
    """Infinite-width graph neural network Gaussian process.

    Args / attributes: identical to the reference's ``GNNGP`` (``src/gnngp.py:6-62``):
    ``data`` is a PyG-``Data``-like object (``x, y, edge_index, edge_attr, train_mask, val_mask,
    test_mask`` and ``.to(device)``), ``method`` in {GCN, GCN2, GIN, SAGE, GGP, RBF}, ``initial`` in
    {linear, rbf, laplacian, arccos, sigmoid, polynomial}; ``mask["landmark"]`` is set by the caller
    for the Nystrom path.  The device must be a CUDA device: there is no CPU fallback.
    """

    _HYPER = ("L", "sigma_b", "sigma_w", "Nystrom", "initial", "method")

    def __init__(self, data, device: torch.device = None, L: int = 2, sigma_b: float = 0.1, sigma_w: float = 1.0,
                 Nystrom: bool = False, initial: str = "linear", method: str = "GCN", **params):
        self.device = device
        self.set_hyper_param(L, sigma_b, sigma_w, Nystrom, initial, method, **params)
        graph = data.to(device)
        self.N, self.X, self.y, self._A, self.mask = datasets.get_data(graph)
        # streamed upload (opt-in, datasets.GraphData._streamed_to): the edge list is still arriving — convert it
        # to a row-chunked CSR piece by piece so the first sparse product overlaps with the rest of the copy
        self._streamed = None
        self._upload_done = None
        plan = getattr(graph, "_stream_plan", None)
        if plan is not None:
            from .ops import default_ops
            self._streamed = default_ops().register_streamed(self._A, plan)
            self._upload_done = plan["all"]
        # Hyper-parameter sweeps (figure3.sh: L / sigma_b / sigma_w grids on one graph) can keep the first
        # propagation A @ Q0 across set_hyper_param calls.  Opt-in — set the attribute or
        # GNNGP_REUSE_PRODUCTS=1 — because the reference recomputes everything on every call.
        self.reuse_products = os.environ.get("GNNGP_REUSE_PRODUCTS", "0") == "1"
        self._product_cache = {}

    @property
    def A(self) -> Tensor:
        """The adjacency as the reference exposes it (sparse COO, ``src/gnngp.py:69``).  After a streamed upload the
        edge list may still be arriving on the copy stream: the first read from outside orders the caller's stream
        after the whole upload, so ``model.A @ X`` never sees a partially written tensor."""
        if self._upload_done is not None:
            torch.cuda.current_stream(self._A.device).wait_event(self._upload_done)
            self._upload_done = None
        return self._A

    @A.setter
    def A(self, value: Tensor) -> None:
        self._A, self._upload_done = value, None

    def set_hyper_param(self, L: int = None, sigma_b: float = None, sigma_w: float = None, Nystrom: bool = None,
                        initial: str = None, method: str = None, **params) -> None:
        """Update the hyper-parameters that are given, REPLACE the extra ``params`` and drop the cached
        kernel — the contract of reference ``src/gnngp.py:71-85``."""
        given = dict(zip(self._HYPER, (L, sigma_b, sigma_w, Nystrom, initial, method)))
        for name, value in given.items():
            if value is not None:
                setattr(self, name, value)
        self.params = params
        self.computed = False

    def _compute_kernel(self) -> None:
        p = self.params
        if self.Nystrom:
            land = self.mask["landmark"]
            self.Q0 = compute._init_kernel_Nystrom(self.X, land, self.initial, **p)
            self.Q = compute._get_kernel_Nystrom(self.Q0, self._A, self.L, self.sigma_b, self.sigma_w, land,
                                                 self.method, _product_cache=self._product_cache if self.reuse_products else None,
                                                 **p)
        else:
            self.C0 = compute._init_kernel(self.X, self.initial, **p)
            self.K = compute._get_kernel(self.C0, self._A, self.L, self.sigma_b, self.sigma_w, self.method, **p)
        self.computed = True

    def get_kernel(self) -> Tensor:
        """``Q`` (Nystrom) or ``K`` (exact); memoised until the next ``set_hyper_param``
        (reference ``src/gnngp.py:87-99``)."""
        if not self.computed:
            if self._streamed is None:
                self._compute_kernel()
            else:
                self._compute_kernel_streamed()
        return self.Q if self.Nystrom else self.K

    def _compute_kernel_streamed(self) -> None:
        """First kernel after a streamed upload: the row-chunked CSR is only valid if the host edge list was
        sorted by row.  The device-side check is read AFTER the recursion has been enqueued (reading it first
        would stall the host until the upload ends and forfeit the overlap); an unsorted list gives a well-formed
        but wrong CSR (possibly NaNs, which the eigensolver rejects), so a failure is treated like a failed check."""
        try:
            self._compute_kernel()
        except torch.linalg.LinAlgError:                     # the one failure a malformed CSR can cause: NaNs reach eigh
            if self._streamed_edges_ok():
                raise                                        # the edge list was fine: a genuine numerical failure
            self._compute_kernel()
            return
        if not self._streamed_edges_ok():
            self._compute_kernel()                           # recomputed from a plain CSR

    def predict(self, nugget: Union[float, Tensor] = 1e-2):
        """Posterior means for every nugget, then the per-mask metrics (reference ``src/gnngp.py:101-124``).
        Sets ``fit``, ``result`` and ``nugget``; returns ``fit``.  For a single nugget the reference's
        ``result`` mis-reads the squeezed ``fit``; here the metric is still computed per
        nugget."""
        kernel = self.get_kernel()
        grid = torch.tensor([nugget]) if isinstance(nugget, float) else nugget
        fit_fn = predict.fit_Nystrom if self.Nystrom else predict.fit
        self.fit = fit_fn(kernel, self.y, self.mask["train"], grid)
        per_nugget = self.fit if len(grid) > 1 else self.fit.unsqueeze(0)
        self.result = predict.result(per_nugget, self.y, self.mask)
        self.nugget = grid
        return self.fit

    def _streamed_edges_ok(self) -> bool:
        """The row-chunked CSR of a streamed upload assumes the host edge list was sorted by row; the pieces
        checked that on the device.  If it was not, drop the chunked CSR (``csr_of`` then converts the complete
        edge list the ordinary way) and report it so that the caller recomputes."""
        chunked, self._streamed = self._streamed, None
        torch.cuda.current_stream(chunked.in_range.device).wait_event(chunked.blocks[-1].ready)
        if bool(chunked.in_range.item()):
            return True
        import warnings
        from .ops import default_ops
        warnings.warn("gnngp_b200: streamed upload found an edge list that is not sorted by row; recomputing "
                      "from a plain CSR")
        default_ops().evict(self._A)                         # only this adjacency's entry; other models keep theirs
        self.computed = False
        return False

    def get_summary(self):
        """Hyper-parameters plus, once ``predict`` has run, the nugget with the best validation metric and
        the train / val / test metrics there (reference ``src/gnngp.py:126-135``)."""
        summary = {name: getattr(self, name) for name in self._HYPER}
        scores = getattr(self, "result", None)
        if scores is not None:
            best = torch.argmax(scores["val"])
            summary["nugget"] = self.nugget[best]
            for split in ("train", "val", "test"):
                summary[split] = scores[split][best]
        return summary
