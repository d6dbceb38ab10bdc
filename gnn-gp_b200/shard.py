"""Row-sharded multi-GPU GCNGP-X (one process per GPU, ``torch.distributed`` over NCCL / NVLink).

The path shards by node rows: rank r owns a contiguous block of rows of the adjacency
(CSR rows, all columns), of the factor ``Q`` and of the fitted values.  The exchange steps are exactly
the three the algorithm has:

  * before every sparse product, an all-gather of the row blocks of the factor being propagated
    (``A_r @ X`` gathers arbitrary rows of ``X``) — ``all_gather_into_tensor`` on equal-sized blocks;
  * the landmark rows ``Q[mask]`` and the landmark block ``E[mask]`` (N_a x m, N_a x N_a), assembled by
    an all-reduce of zero-padded contributions; the N_a x N_a eigendecomposition is replicated
    (identical inputs on every rank → identical factors);
  * in the fit, all-reduces of ``Q_bᵀQ_b`` (m x m), ``Q_bᵀy_b`` (m x C), the nugget scale and the metric
    counts.

Everything else (Gram + arc-cosine, back-projection, nugget sweep, argmax metrics) is row-local.  The
exact (non-Nystrom) path is small (N ≤ 2 708 in the configs) and is not sharded: replicas only.

``ShardedGNNGP`` mirrors the surface of :class:`gnngp_b200.gnngp.GNNGP`; ``fit`` holds the local rows
``[G, n_local, C]``, ``result`` / ``get_summary`` are global and identical on every rank.
"""
from __future__ import annotations

import os
import warnings
from typing import Union

import torch
import torch.distributed as dist
from torch import Tensor

from . import datasets
from .pipeline import Landmarks, Pipeline, RowLayout


def _round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


class TorchComm(object):
    """The collectives of the sharded path on a ``torch.distributed`` process group."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.bytes_gathered = 0
        self.bytes_reduced = 0
        self.bytes_pulled = 0
        # Fused all-gather + panel layout over peer memory (NVLink P2P loads inside our own kernel) instead of an NCCL
        # all-gather into a staging buffer followed by the panel copy.  Exchange buffers are plain device allocations
        # shared through CUDA IPC (gnngp_exchange_*); the two cross-rank ordering points per exchange are one-element
        # NCCL all-reduces, which are stream-ordered — no symmetric-memory rendezvous (round 1's torch
        # symmetric-memory version stalled with 7 GB blocks).  GNNGP_PEER_GATHER=1 turns it on (dense-row graphs, NCCL).
        self.peer_gather = (os.environ.get("GNNGP_PEER_GATHER", "0") == "1" and torch.cuda.is_available()
                            and dist.get_backend(group) == "nccl")
        self._exchange = {}
        self._token = None

    def _exchange_block(self, ops, rows: int, ld: int, device):
        """This rank's exchange buffer for [rows x ld] fp32 blocks and a device table of all ranks' pointers."""
        key = (rows, ld)
        hit = self._exchange.get(key)
        if hit is None:
            import ctypes
            with torch.cuda.device(device):
                handle = ctypes.create_string_buffer(64)
                mine = int(ops.lib.query("gnngp_exchange_alloc", rows * ld * 4, ctypes.addressof(handle)))
                ok = torch.tensor([1 if mine else 0], device=device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)      # every rank takes the same branch
                if int(ok) == 0:
                    raise RuntimeError("exchange buffer allocation failed on some rank: %s" % ops.lib.last_error())
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(handle.raw), group=self.group)
                ptrs = []
                for r, raw in enumerate(handles):
                    if r == self.rank:
                        ptrs.append(mine)
                    else:
                        peer = int(ops.lib.query("gnngp_exchange_open", raw))
                        if not peer:
                            raise RuntimeError("cudaIpcOpenMemHandle failed for rank %d: %s" % (r, ops.lib.last_error()))
                        ptrs.append(peer)
                table = torch.tensor(ptrs, dtype=torch.int64, device=device)
            hit = (mine, table)
            self._exchange[key] = hit
        return hit

    def _stream_barrier(self, device) -> None:
        """Cross-rank ordering point on the current stream: a one-element NCCL all-reduce."""
        if self._token is None:
            self._token = torch.zeros(1, device=device)
        dist.all_reduce(self._token, group=self.group)

    def pull_panels(self, ops, block: Tensor, layout: RowLayout):
        """Fused all-gather + panel layout for the sparse product: every rank publishes its row block in its
        exchange buffer, then ONE kernel per rank (``panelize_peers_kernel``) pulls all blocks over NVLink straight
        into the 64-column panels its SpMM gathers from (no gathered [N x k] buffer, no NCCL data movement).
        Returns (workspace tensor, panel pointer) or None when peer memory is unavailable (→ NCCL all-gather)."""
        if not self.peer_gather:
            return None
        n_local, k = block.shape
        ld = _round_up(max(k, 1), 4)
        try:
            mine, table = self._exchange_block(ops, layout.rows_per_rank, ld, block.device)
        except RuntimeError as exc:        # raised on every rank alike (the allocation outcome is all-reduced)
            warnings.warn("gnngp_b200: peer exchange buffers unavailable (%s); using the NCCL all-gather" % (exc,))
            self.peer_gather = False
            return None
        src = block if (k <= 1 or block.stride(1) == 1) else block.contiguous()
        ops.lib.call("gnngp_copy2d_f32", src.data_ptr(), int(src.stride(0)) if n_local > 1 else max(k, 1), n_local, k, mine, ld,
                     torch.cuda.current_stream(block.device).cuda_stream)
        self._stream_barrier(block.device)                 # every rank's block is published
        ws, ptr = ops.panels_from_peers(table.data_ptr(), self.world, layout.rows_per_rank, ld, layout.n, k, block.device)
        self._stream_barrier(block.device)                 # every rank has finished reading before a block is reused
        self.bytes_pulled += (self.world - 1) * layout.rows_per_rank * k * 4
        return ws, ptr

    def all_reduce_sum(self, t: Tensor) -> Tensor:
        if t.is_contiguous():
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        else:                                   # padded view: reduce a packed copy
            packed = t.contiguous()
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=self.group)
            t.copy_(packed)
        self.bytes_reduced += t.numel() * t.element_size()
        return t

    def all_gather_equal(self, piece: Tensor, out: Tensor) -> Tensor:
        """Equal-sized contiguous pieces (rank order) → ``out`` [world * piece.shape[0], ...] on every rank (any dtype):
        the row blocks of the replicated Krylov routes' products."""
        dist.all_gather_into_tensor(out, piece, group=self.group)
        self.bytes_gathered += out.numel() * out.element_size()
        return out

    def all_gather_rows(self, block: Tensor, layout: RowLayout) -> Tensor:
        """[n_local x k] row blocks (rank order = row order) → the full [N x k] matrix on every rank."""
        n_local, k = block.shape
        ld = _round_up(max(k, 1), 4)
        in_place = (n_local == layout.rows_per_rank and n_local > 1 and block.stride(0) == ld
                    and (k <= 1 or block.stride(1) == 1))
        if in_place:                            # an equal-sized buffer with the padded leading dimension
            slab = block.as_strided((n_local, ld), (ld, 1), block.storage_offset())
        else:
            slab = torch.zeros((layout.rows_per_rank, ld), dtype=block.dtype, device=block.device)
            slab[:n_local, :k] = block
        full = torch.empty((self.world * layout.rows_per_rank, ld), dtype=block.dtype, device=block.device)
        dist.all_gather_into_tensor(full, slab, group=self.group)
        self.bytes_gathered += full.numel() * full.element_size()
        return full[:layout.n, :k]


class ShardedGNNGP(object):
    """GCNGP-X over a process group; constructor and methods follow ``GNNGP`` (reference ``src/gnngp.py:63-135``).

    ``data`` is the FULL graph on every rank (features and labels are small next to the factor); each rank
    keeps its rows of the adjacency as CSR and its rows of every N x m matrix.  ``ops`` defaults to the CUDA
    operator set; ``group`` to the default process group."""

    _HYPER = ("L", "sigma_b", "sigma_w", "Nystrom", "initial", "method")

    def __init__(self, data, device: torch.device = None, L: int = 2, sigma_b: float = 0.1, sigma_w: float = 1.0,
                 Nystrom: bool = True, initial: str = "linear", method: str = "GCN", group=None, ops=None, **params):
        if ops is None:
            from .ops import default_ops
            ops = default_ops()
        self.ops = ops
        self.comm = TorchComm(group)
        self.device = device
        self.set_hyper_param(L, sigma_b, sigma_w, Nystrom, initial, method, **params)
        graph = data.to(device)
        plan = getattr(graph, "_stream_plan", None)
        if plan is not None:                                # streamed upload: this path reads the whole edge list at once
            torch.cuda.current_stream(graph.edge_index.device).wait_event(plan["all"])
        self.N, self.X, self.y, self.A, self.mask = datasets.get_data(graph)
        self.layout = RowLayout(self.N, self.comm.world, self.comm.rank)
        lay = self.layout
        idx, val = self.A._indices(), self.A._values()
        keep = (idx[0] >= lay.begin) & (idx[0] < lay.end)
        self.csr = ops.csr_from_coo((idx[0][keep] - lay.begin).contiguous(), idx[1][keep].contiguous(),
                                    val[keep].to(torch.float32).contiguous(), lay.n_local, self.N, lay.begin)
        self.X_local = self.X[lay.begin:lay.end].to(torch.float32).contiguous()
        self.pipe = Pipeline(ops, self.comm, lay)
        if os.environ.get("GNNGP_REUSE_PRODUCTS", "0") == "1":       # hyper-parameter sweeps: keep A @ Q0
            self.pipe.product_cache = {}
        total = torch.tensor([float(self.csr.nnz)], dtype=torch.float64, device=self.X_local.device)
        self.comm.all_reduce_sum(total)
        self.pipe.global_nnz = (int(total.item()), self.N)

    def set_hyper_param(self, L=None, sigma_b=None, sigma_w=None, Nystrom=None, initial=None, method=None, **params):
        for name, value in zip(self._HYPER, (L, sigma_b, sigma_w, Nystrom, initial, method)):
            if value is not None:
                setattr(self, name, value)
        self.params = params
        self.computed = False

    def _local(self, t: Tensor) -> Tensor:
        return t[self.layout.begin:self.layout.end]

    def get_kernel(self) -> Tensor:
        """This rank's rows of the Nystrom factor ``Q``."""
        if not self.Nystrom:
            raise RuntimeError("ShardedGNNGP shards the Nystrom path only; run the exact path as replicas with GNNGP")
        if not self.computed:
            land = Landmarks(self.mask["landmark"], self.layout)
            self.Q0 = self.pipe.initial_factor(self.X_local, land, self.initial, self.params)
            self.Q = self.pipe.kernel_nystrom(self.Q0, self.csr, self.L, self.sigma_b, self.sigma_w, land, self.method,
                                              self.params)
            self.computed = True
        return self.Q

    def predict(self, nugget: Union[float, Tensor] = 1e-2):
        q = self.get_kernel()
        grid = torch.tensor([nugget]) if isinstance(nugget, float) else nugget
        grid = grid.to(device=q.device, dtype=torch.float32).contiguous()
        train = self.mask["train"]
        rows_global = torch.nonzero(train, as_tuple=False).flatten()
        classes = None if self.y.dtype.is_floating_point else int(self.y[rows_global].max()) + 1
        rows_local = torch.nonzero(self._local(train), as_tuple=False).flatten()
        y_local = self._local(self.y)
        self.fit = self.pipe.fit_nystrom(q, y_local, rows_local, grid, self.N, classes)
        per_nugget = self.fit if grid.numel() > 1 else self.fit.unsqueeze(0)
        self.result = self.pipe.metrics(per_nugget, y_local, {k: self._local(v) for k, v in self.mask.items()})
        self.nugget = grid
        return self.fit

    def get_summary(self):
        summary = {name: getattr(self, name) for name in self._HYPER}
        scores = getattr(self, "result", None)
        if scores is not None:
            best = torch.argmax(scores["val"])
            summary["nugget"] = self.nugget[best]
            for split in ("train", "val", "test"):
                summary[split] = scores[split][best]
        return summary

    def gather_fit(self) -> Tensor:
        """All ranks' fitted rows, [G, N, ...] on every rank (for tests / small graphs)."""
        local = self.fit
        g = local.shape[0]
        flat = local.reshape(g, local.shape[1], -1).permute(1, 0, 2).reshape(local.shape[1], -1).contiguous()
        full = self.comm.all_gather_rows(flat, self.layout)
        c = flat.shape[1] // g
        return full.reshape(self.N, g, c).permute(1, 0, 2).reshape((g, self.N) + tuple(local.shape[2:]))
