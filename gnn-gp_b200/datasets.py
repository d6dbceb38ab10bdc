"""Data boundary of the GNNGP hot path, in pure torch (no PyG / torch_sparse / torch_scatter).

Mirrors the three pieces of the reference's ``src/datasets.py`` that the GP path touches:

* ``transition_matrix`` — self-loops + coalesce + sym/col/row degree normalisation
  (reference ``src/datasets.py:10-68``),
* ``Scale`` — column centring / unit-norm scaling (``src/datasets.py:71-81``),
* ``get_data`` — (N, X, y, A, masks) with ``A`` a sparse COO tensor (``src/datasets.py:313-330``).

Dataset download / parsing (``src/datasets.py:84-310``) is out of scope (no network); graphs
come from :mod:`gnngp_b200.synth` or from any object exposing the same attributes as a PyG ``Data``.
"""
from __future__ import annotations

import torch


class GraphData(object):
    """Minimal stand-in for ``torch_geometric.data.Data`` (attribute access + ``.to``)."""

    _tensor_fields = ("x", "y", "edge_index", "edge_attr", "train_mask", "val_mask", "test_mask")

    def __init__(self, x, y, edge_index, edge_attr=None, train_mask=None, val_mask=None,
                 test_mask=None, num_classes=None):
        self.x = x
        self.y = y
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        self.train_mask = train_mask
        self.val_mask = val_mask
        self.test_mask = test_mask
        self.num_classes = num_classes

    @property
    def num_nodes(self) -> int:
        return int(self.x.shape[0])

    def to(self, device):
        if device is None:
            return self
        out = GraphData.__new__(GraphData)
        out.num_classes = self.num_classes
        for name in self._tensor_fields:
            value = getattr(self, name)
            setattr(out, name, value.to(device) if value is not None else None)
        return out


def coalesce_edges(edge_index: torch.Tensor, edge_weight: torch.Tensor, num_nodes: int):
    """Sort edges by (row, col) and sum duplicates — what ``torch_sparse.coalesce`` does
    at reference ``src/datasets.py:41``."""
    key = edge_index[0] * num_nodes + edge_index[1]
    uniq, inverse = torch.unique(key, sorted=True, return_inverse=True)
    weight = torch.zeros(uniq.numel(), dtype=edge_weight.dtype, device=edge_weight.device)
    weight.index_add_(0, inverse, edge_weight)
    rows = torch.div(uniq, num_nodes, rounding_mode="floor")
    cols = uniq - rows * num_nodes
    return torch.stack([rows, cols]), weight


class transition_matrix(object):
    """Graph normalisation transform; same constructor and call contract as the reference's
    ``transition_matrix`` (``src/datasets.py:10-68``): ``normalization`` in
    {"sym", "col", "row", other = none}, self-loop weight added for every node."""

    def __init__(self, self_loop_weight=1, normalization="sym"):
        self.self_loop_weight = self_loop_weight
        self.normalization = normalization

    def __call__(self, data):
        n = data.num_nodes
        index = data.edge_index
        if data.edge_attr is None:
            weight = torch.ones(index.size(1), device=index.device)
        else:
            weight = data.edge_attr
        if self.self_loop_weight:
            loops = torch.arange(n, device=index.device, dtype=index.dtype)
            index = torch.cat([index, torch.stack([loops, loops])], dim=1)
            weight = torch.cat([weight, torch.full((n,), float(self.self_loop_weight),
                                                   dtype=weight.dtype, device=weight.device)])
        index, weight = coalesce_edges(index, weight, n)
        row, col = index[0], index[1]
        if self.normalization == "sym":
            deg = torch.zeros(n, dtype=weight.dtype, device=weight.device).index_add_(0, col, weight)
            scale = deg.pow(-0.5)
            scale[scale == float("inf")] = 0
            weight = scale[row] * weight * scale[col]
        elif self.normalization == "col":
            deg = torch.zeros(n, dtype=weight.dtype, device=weight.device).index_add_(0, col, weight)
            scale = 1.0 / deg
            scale[scale == float("inf")] = 0
            weight = weight * scale[col]
        elif self.normalization == "row":
            deg = torch.zeros(n, dtype=weight.dtype, device=weight.device).index_add_(0, row, weight)
            scale = 1.0 / deg
            scale[scale == float("inf")] = 0
            weight = weight * scale[row]
        data.edge_index = index
        data.edge_attr = weight
        return data


class Scale(object):
    """Feature centring / scaling (reference ``src/datasets.py:71-81``); mutates ``data.x``."""

    def __init__(self, center=True, scale=False):
        self.center = center
        self.scale = scale

    def __call__(self, data):
        if self.center:
            data.x -= torch.mean(data.x, dim=0, keepdim=True)
        if self.scale:
            data.x /= torch.sqrt(torch.sum(data.x ** 2, dim=0, keepdim=True))
        return data


def get_data(data):
    """(N, X, y, A, mask) from a graph object — reference ``src/datasets.py:313-330``.

    ``A`` is an (uncoalesced-flagged) sparse COO tensor exactly as the reference builds it; the
    CUDA path converts it to CSR once and caches that (see ``ops.csr_of``)."""
    x = data.x
    n = x.shape[0]
    if data.edge_attr is None:
        weight = torch.ones(data.edge_index.size(1), device=data.edge_index.device)
    else:
        weight = data.edge_attr
    adj = torch.sparse_coo_tensor(data.edge_index, weight, (n, n))
    mask = {"train": data.train_mask, "val": data.val_mask, "test": data.test_mask}
    return n, x, data.y, adj, mask
