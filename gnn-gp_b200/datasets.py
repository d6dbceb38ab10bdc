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

import os

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
        device = torch.device(device)
        if (device.type == "cuda" and os.environ.get("GNNGP_STREAM_UPLOAD", "0") == "1"
                and self.edge_index is not None and self.edge_attr is not None
                and not self.edge_index.is_cuda and self.edge_index.is_pinned() and self.edge_attr.is_pinned()
                and self.edge_index.size(1) >= int(os.environ.get("GNNGP_STREAM_UPLOAD_MIN_NNZ", str(1 << 22)))):
            return self._streamed_to(device)
        out = GraphData.__new__(GraphData)
        out.num_classes = self.num_classes
        for name in self._tensor_fields:
            value = getattr(self, name)
            if value is not None:                      # pinned host buffers: stream-ordered asynchronous copies
                value = value.to(device, non_blocking=bool(not value.is_cuda and value.is_pinned()))
            setattr(out, name, value)
        return out

    def _streamed_to(self, device, chunks: int = 8):
        """Host -> device copy of a large graph that lets the consumer start before the edge list has arrived
        (opt-in, ``GNNGP_STREAM_UPLOAD=1``; pinned host tensors, edge list sorted by row as ``transition_matrix``
        / ``coalesce`` leave it).  Features, labels and masks go first on a copy stream; the edge list follows
        in ``chunks`` row-aligned pieces, each with its own event.  The returned object carries
        ``_stream_plan`` = row bounds, nonzero bounds and events; ``GNNGP`` turns it into a row-chunked CSR
        whose pieces are converted and multiplied as they arrive (``ops.register_streamed``)."""
        n = self.num_nodes
        nnz = int(self.edge_index.size(1))
        main = torch.cuda.current_stream(device)
        copy = _copy_stream(device)
        row_bounds, nnz_bounds = chunk_bounds(self.edge_index[0], n, chunks)
        out = GraphData.__new__(GraphData)
        out.num_classes = self.num_classes
        small = {}
        for name in self._tensor_fields:
            value = getattr(self, name)
            if name in ("edge_index", "edge_attr") or value is None:
                continue
            small[name] = torch.empty(value.shape, dtype=value.dtype, device=device)
        edge_index = torch.empty((2, nnz), dtype=self.edge_index.dtype, device=device)
        edge_attr = torch.empty((nnz,), dtype=self.edge_attr.dtype, device=device)
        copy.wait_stream(main)                             # the buffers above are safe to write from here on
        events = []
        with torch.cuda.stream(copy):
            for name, dst in small.items():
                dst.copy_(getattr(self, name), non_blocking=True)
            ready_small = torch.cuda.Event()
            ready_small.record(copy)
            for c in range(chunks):
                a, b = nnz_bounds[c], nnz_bounds[c + 1]
                if b > a:
                    edge_index[0, a:b].copy_(self.edge_index[0, a:b], non_blocking=True)
                    edge_index[1, a:b].copy_(self.edge_index[1, a:b], non_blocking=True)
                    edge_attr[a:b].copy_(self.edge_attr[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
                events.append(ev)
        for t in list(small.values()) + [edge_index, edge_attr]:
            t.record_stream(copy)
        main.wait_event(ready_small)                       # features / labels / masks: needed first, small
        for name in self._tensor_fields:
            setattr(out, name, small.get(name))
        out.edge_index, out.edge_attr = edge_index, edge_attr
        out._stream_plan = {"row_bounds": row_bounds, "nnz_bounds": nnz_bounds, "events": events, "all": events[-1]}
        return out


def chunk_bounds(rows: torch.Tensor, n: int, chunks: int):
    """Row-aligned pieces of a row-sorted edge list: ``row_bounds[c] .. row_bounds[c+1]`` are the rows of piece
    c and ``nnz_bounds[c] .. nnz_bounds[c+1]`` its edges (binary searches on the HOST copy of the row ids).
    The edge ranges are monotone and tile ``[0, nnz)`` even when the list is NOT sorted — every edge is copied
    exactly once and the device-side range check then reports the misplaced ones."""
    nnz = int(rows.numel())
    row_bounds = [(n * c) // chunks for c in range(chunks + 1)]
    inner = torch.tensor(row_bounds[1:-1], dtype=rows.dtype)
    cuts = torch.searchsorted(rows.contiguous(), inner).tolist()
    nnz_bounds = [0]
    for c in cuts:
        nnz_bounds.append(min(nnz, max(nnz_bounds[-1], int(c))))
    nnz_bounds.append(nnz)
    return row_bounds, nnz_bounds


_COPY_STREAMS = {}


def _copy_stream(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COPY_STREAMS[key]


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
