"""Tensor-level operator set of the GNNGP hot path on B200.

Every method enqueues hand-written sm_100a kernels of ``libgnngp_b200.so`` (through the C ABI of
``include/gnngp_b200.h``) on torch's current CUDA stream.  PyTorch only owns memory, streams and the
symmetric eigensolver call (``torch.linalg.eigh`` → cuSOLVER; SURVEY.md §7 hard part 1).  Inputs must
be CUDA fp32 tensors; anything else raises — there is no CPU path.

Matrices produced here are views ``buf[:, :m]`` of buffers whose leading dimension is padded to a
multiple of 4 floats, so that the SpMM can gather them with 128-bit loads.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from . import _lib


def _round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


class CsrMatrix(object):
    """Device CSR of the normalised adjacency (row pointer int64, column ids int32, values fp32).
    ``row_offset`` / ``n_rows`` describe the row block a rank owns when the graph is row-sharded."""

    def __init__(self, rowptr, colidx, vals, n_rows, n_cols, row_offset=0):
        self.rowptr, self.colidx, self.vals = rowptr, colidx, vals
        self.n_rows, self.n_cols, self.row_offset = int(n_rows), int(n_cols), int(row_offset)
        self.nnz = int(colidx.numel())


class ChunkedCsr(object):
    """The adjacency as consecutive row blocks, each a :class:`CsrMatrix` over all columns (``row_offset`` =
    first row).  Built by ``CudaOps.register_streamed`` from an edge list that is still arriving from the host:
    every block is converted — and can be multiplied — as soon as ITS edges are on the device."""

    def __init__(self, n_rows: int, n_cols: int, blocks, in_range: torch.Tensor):
        self.n_rows, self.n_cols, self.blocks, self.row_offset = int(n_rows), int(n_cols), list(blocks), 0
        self.nnz = sum(b.nnz for b in self.blocks)
        self.in_range = in_range          # device flag: every edge of every block lies inside the block's row range

    def merged(self) -> CsrMatrix:
        """One CsrMatrix over all rows (concatenation; the blocks are consecutive and internally sorted)."""
        offs, total = [], 0
        for b in self.blocks:
            torch.cuda.current_stream(b.rowptr.device).wait_event(b.ready)
            offs.append(total)
            total += b.nnz
        rowptr = torch.cat([b.rowptr[:-1] + o for b, o in zip(self.blocks, offs)] +
                           [torch.tensor([total], dtype=torch.int64, device=self.blocks[0].rowptr.device)])
        return CsrMatrix(rowptr, torch.cat([b.colidx for b in self.blocks]), torch.cat([b.vals for b in self.blocks]),
                         self.n_rows, self.n_cols, 0)


class CudaOps(object):
    name = "cuda-sm100a"

    def __init__(self):
        self.lib = _lib.library()
        self._workspace = {}
        self._csr_cache = {}
        self.spmm_tile_cols = 0       # 0 = panel layout (default); 32/64/128 = row-major gather with that tile
        self.spmm_nnz_per_warp = 0
        # GNNGP_SPMM_PANELS = 1 / 0 forces / forbids the 64-column panel copy of the gathered operand (default: by shape)
        self.spmm_panels_mode = os.environ.get("GNNGP_SPMM_PANELS", "auto")
        self.spmm_flags = int(self.lib.query("gnngp_spmm_set_flags", 0))     # read the library's setting ...
        self.lib.query("gnngp_spmm_set_flags", self.spmm_flags)              # ... and put it back
        self.spmm_kernel_by_k = {}    # operand width -> name of the kernel the last product of that width ran (bench roofline)

    # ------------------------------------------------------------------ plumbing
    @staticmethod
    def _stream() -> int:
        return torch.cuda.current_stream().cuda_stream

    @staticmethod
    def _check(t: torch.Tensor, name: str, dtype=torch.float32, dims: Optional[int] = None) -> torch.Tensor:
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise RuntimeError("gnngp_b200: %s must be a CUDA tensor (the hot path has no CPU fallback)" % name)
        if t.dtype != dtype:
            raise RuntimeError("gnngp_b200: %s must be %s, got %s" % (name, dtype, t.dtype))
        if dims is not None and t.dim() != dims:
            raise RuntimeError("gnngp_b200: %s must be %d-D" % (name, dims))
        return t

    def _mat(self, t: torch.Tensor, name: str) -> Tuple[torch.Tensor, int]:
        """Row-major matrix with unit column stride → (tensor, leading dimension)."""
        self._check(t, name, dims=2)
        if t.shape[1] > 1 and t.stride(1) != 1:
            t = t.contiguous()
        if t.shape[0] > 1 and t.stride(0) < t.shape[1]:
            t = t.contiguous()
        return t, self._ld(t)

    @staticmethod
    def _ld(t: torch.Tensor) -> int:
        """Leading dimension of a row-major 2-D view (a single row may use any ld >= its width)."""
        return int(t.stride(0)) if t.shape[0] > 1 else max(int(t.shape[1]), 1)

    def _vec(self, t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
        self._check(t, name, dtype=dtype, dims=1)
        return t if t.numel() <= 1 or t.stride(0) == 1 else t.contiguous()

    def workspace(self, nbytes: int, device, slot: str = "main") -> torch.Tensor:
        key = (device.index, slot)
        buf = self._workspace.get(key)
        if buf is None or buf.numel() < nbytes:
            self._workspace[key] = None
            buf = torch.empty(int(nbytes * 1.1) + 256, dtype=torch.uint8, device=device)
            self._workspace[key] = buf
        return buf

    def release_workspace(self):
        self._workspace.clear()

    def side_stream(self, device):
        """A second, high-priority stream per device for running the eigensolver next to the sparse product,
        or None (the default).  Measured on B200 (profiles/r1/call11_overlap_ab.txt): cuSOLVER's
        tridiagonalisation kernels fill the grid and are L2-bound like the SpMM, so concurrent execution
        stretches both and the step time does not improve (0.339 s single stream vs 0.343 s overlapped at
        fraction 50; 6.83 vs 6.86 s at fraction 10).  GNNGP_OVERLAP=1 turns the overlapped schedule on."""
        if os.environ.get("GNNGP_OVERLAP") != "1":
            return None
        key = ("side", device.index)
        stream = self._workspace.get(key)
        if stream is None:
            # highest priority: the eigensolver's many small kernels must not queue behind the grid-filling
            # SpMM / tensor-core kernels of the main stream they are meant to overlap with
            stream = torch.cuda.Stream(device=device, priority=-1)
            self._workspace[key] = stream
        return stream

    def empty(self, n: int, m: int, device) -> torch.Tensor:
        buf = torch.empty((n, _round_up(max(m, 1), 4)), dtype=torch.float32, device=device)
        return buf[:, :m]

    def zeros(self, n: int, m: int, device) -> torch.Tensor:
        buf = torch.zeros((n, _round_up(max(m, 1), 4)), dtype=torch.float32, device=device)
        return buf[:, :m]

    def launch_count(self) -> int:
        return int(self.lib.query("gnngp_launch_count"))

    def tc_launch_count(self) -> int:
        return int(self.lib.query("gnngp_tc_launch_count"))

    def set_gemm_engine(self, engine: int) -> int:
        return int(self.lib.query("gnngp_set_gemm_engine", int(engine)))

    def set_spmm_flags(self, flags: int) -> int:
        """Tuning flags of the SpMM kernels (``gnngp_spmm_set_flags``: 1 streaming 16-byte stores, 2 zero-fill of cut /
        empty rows only, 4 short-row kernel); returns the previous setting."""
        self.spmm_flags = int(flags) & 7
        return int(self.lib.query("gnngp_spmm_set_flags", int(flags)))

    # ------------------------------------------------------------------ adjacency
    def csr_from_coo(self, row: torch.Tensor, col: torch.Tensor, val: torch.Tensor, n_rows: int, n_cols: int,
                     row_offset: int = 0) -> CsrMatrix:
        self._check(row, "row", torch.int64, 1)
        self._check(col, "col", torch.int64, 1)
        self._check(val, "val", torch.float32, 1)
        row, col, val = row.contiguous(), col.contiguous(), val.contiguous()
        nnz = row.numel()
        dev = row.device
        rowptr = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
        colidx = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)[:nnz]
        vals = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)[:nnz]
        need = self.lib.query("gnngp_csr_from_coo_workspace_bytes", nnz, n_rows)
        ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)     # one-off, not kept
        wptr = (ws.data_ptr() + 255) // 256 * 256
        self.lib.call("gnngp_csr_from_coo", row.data_ptr(), col.data_ptr(), val.data_ptr(), nnz, n_rows, n_cols,
                      rowptr.data_ptr(), colidx.data_ptr(), vals.data_ptr(), wptr, need, self._stream())
        ws.record_stream(torch.cuda.current_stream())
        return CsrMatrix(rowptr, colidx, vals, n_rows, n_cols, row_offset)

    def csr_of(self, A: torch.Tensor) -> CsrMatrix:
        """CSR of a torch sparse COO adjacency, converted once and cached on the tensor's storage."""
        if isinstance(A, (CsrMatrix, ChunkedCsr)):
            return A
        if not A.is_sparse:
            raise RuntimeError("gnngp_b200: adjacency must be a torch sparse COO tensor")
        idx, val = A._indices(), A._values()
        key = self._csr_key(A)
        hit = self._csr_cache.get(key)
        if hit is None:
            csr = self.csr_from_coo(idx[0], idx[1], val if val.dtype == torch.float32 else val.to(torch.float32),
                                    A.shape[0], A.shape[1])
            if len(self._csr_cache) >= 2:
                self._csr_cache.clear()
            # the entry keeps the COO tensors alive, so their addresses cannot be recycled for other data while
            # the key is in the cache; in-place edits bump the tensors' version counters, which are part of the key
            hit = (csr, idx, val)
            self._csr_cache[key] = hit
        return hit[0]

    @staticmethod
    def _csr_key(A: torch.Tensor):
        idx, val = A._indices(), A._values()
        return (idx.data_ptr(), val.data_ptr(), int(val.numel()), tuple(A.shape), idx.device.index, idx._version, val._version)

    def clear_cache(self) -> None:
        self._csr_cache.clear()

    def evict(self, A: torch.Tensor) -> None:
        """Drop the cached CSR of this adjacency only."""
        self._csr_cache.pop(self._csr_key(A), None)

    def register_streamed(self, A: torch.Tensor, plan: dict) -> ChunkedCsr:
        """Row-chunked CSR of an adjacency whose edge list is being uploaded in row-aligned pieces
        (``datasets.GraphData._streamed_to``): for every piece the current stream waits for that piece's copy
        event and converts it, so conversion and the first sparse product overlap with the rest of the upload.
        The result is cached like ``csr_of``'s, so every later ``csr_of(A)`` returns it."""
        idx, val = A._indices(), A._values()
        n = int(A.shape[0])
        dev = idx.device
        main = torch.cuda.current_stream(dev)
        convert = self._convert_stream(dev)
        convert.wait_stream(main)
        blocks = []
        rb, nb = plan["row_bounds"], plan["nnz_bounds"]
        # conversions run on their own stream, each after ITS piece's copy event, and leave a `ready` event per
        # block: the main stream only waits for a block right before it multiplies with it (_spmm_chunked), so the
        # product over block c overlaps with the upload and conversion of blocks c+1, c+2, ...
        with torch.cuda.stream(convert):
            in_range = torch.ones((), dtype=torch.bool, device=dev)
            for c, event in enumerate(plan["events"]):
                convert.wait_event(event)
                a, b = nb[c], nb[c + 1]
                r0, r1 = rb[c], rb[c + 1]
                rows = idx[0, a:b]
                if b > a:
                    in_range = in_range & (rows.min() >= r0) & (rows.max() < r1)
                # clamped so that the piece is a well-formed CSR even if the list is NOT sorted by row (the flag above
                # then makes the caller recompute from a plain CSR; the clamped block is never trusted)
                local = (rows - r0).clamp_(0, max(r1 - r0 - 1, 0))
                block = self.csr_from_coo(local, idx[1, a:b], val[a:b] if val.dtype == torch.float32 else val[a:b].float(),
                                          max(r1 - r0, 1), int(A.shape[1]), r0)
                block.n_rows = r1 - r0
                for t in (block.rowptr, block.colidx, block.vals):
                    t.record_stream(main)
                block.ready = torch.cuda.Event()
                block.ready.record(convert)
                blocks.append(block)
            in_range.record_stream(main)
        chunked = ChunkedCsr(n, int(A.shape[1]), blocks, in_range)
        key = self._csr_key(A)
        if len(self._csr_cache) >= 2:
            self._csr_cache.clear()
        self._csr_cache[key] = (chunked, idx, val)
        return chunked

    def _convert_stream(self, device):
        key = (device.index, "convert-stream")
        stream = self._workspace.get(key)
        if stream is None:
            stream = torch.cuda.Stream(device=device)
            self._workspace[key] = stream
        return stream

    def _spmm_chunked(self, csr: ChunkedCsr, B: torch.Tensor, alpha: float, out: Optional[torch.Tensor]) -> torch.Tensor:
        B, ldb = self._mat(B, "B")
        main = torch.cuda.current_stream(B.device)
        if B.shape[0] != csr.n_cols:
            raise RuntimeError("spmm: B has %d rows, adjacency has %d columns" % (B.shape[0], csr.n_cols))
        k = B.shape[1]
        if out is None:
            out = self.empty(csr.n_rows, k, B.device)
        if out.shape != (csr.n_rows, k) or (k > 1 and out.stride(1) != 1):
            raise RuntimeError("spmm: out must be a [%d x %d] row-major view" % (csr.n_rows, k))
        if int(self.spmm_tile_cols) == 0 and csr.nnz >= 64 * csr.n_rows and k >= 256:
            need = self.lib.query("gnngp_spmm_workspace_bytes", csr.n_cols, k)
            ws = self.workspace(need + 256, B.device, "spmm")
            panels = (ws.data_ptr() + 255) // 256 * 256
            self.lib.call("gnngp_panelize_f32", B.data_ptr(), ldb, B.shape[0], k, panels, self._stream())   # once for all blocks
            for blk in csr.blocks:
                main.wait_event(blk.ready)
                if blk.n_rows > 0:
                    self.spmm_panels(blk, panels, k, alpha, out[blk.row_offset:blk.row_offset + blk.n_rows])
        else:
            for blk in csr.blocks:
                main.wait_event(blk.ready)
                if blk.n_rows > 0:
                    self.spmm(blk, B, alpha, out=out[blk.row_offset:blk.row_offset + blk.n_rows])
        return out

    def restrict_columns(self, csr: CsrMatrix, cols: torch.Tensor) -> CsrMatrix:
        """CSR of ``A[:, cols]`` with column ids renumbered to positions in ``cols`` (one pass over the
        nonzeros with torch index ops; done once per landmark set)."""
        if isinstance(csr, ChunkedCsr):
            csr = csr.merged()
        dev = csr.colidx.device
        lookup = torch.full((csr.n_cols,), -1, dtype=torch.int32, device=dev)
        lookup[cols] = torch.arange(cols.numel(), dtype=torch.int32, device=dev)
        renum = lookup[csr.colidx.long()]
        keep = renum >= 0
        running = torch.zeros(csr.nnz + 1, dtype=torch.int64, device=dev)
        torch.cumsum(keep, 0, out=running[1:])
        rowptr = running[csr.rowptr].contiguous()
        return CsrMatrix(rowptr, renum[keep].contiguous(), csr.vals[keep].contiguous(), csr.n_rows, int(cols.numel()),
                         csr.row_offset)

    # ------------------------------------------------------------------ K1 SpMM
    def spmm(self, csr: CsrMatrix, B: torch.Tensor, alpha: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if isinstance(csr, ChunkedCsr):
            return self._spmm_chunked(csr, B, alpha, out)
        B, ldb = self._mat(B, "B")
        if B.shape[0] != csr.n_cols:
            raise RuntimeError("spmm: B has %d rows, adjacency has %d columns" % (B.shape[0], csr.n_cols))
        k = B.shape[1]
        if out is None:
            out = self.empty(csr.n_rows, k, B.device)
        self._check(out, "out", dims=2)
        if out.shape != (csr.n_rows, k) or (k > 1 and out.stride(1) != 1):
            raise RuntimeError("spmm: out must be a [%d x %d] row-major view" % (csr.n_rows, k))
        wptr, need = None, 0
        # Measured on B200 (profiles/r1/call7_spmm_sweep.txt): the 64-column panel copy pays off for dense rows
        # and wide / misaligned operands (Reddit shape: 108 vs 121 ms at k=3024, 22 vs 26 ms at k=602); for
        # short rows (arxiv shape, 15 nonzeros per row) gathering straight from the row-major operand wins.
        use_panels = int(self.spmm_tile_cols) == 0 and csr.nnz >= 64 * csr.n_rows and k >= 256
        if self.spmm_panels_mode in ("0", "1") and int(self.spmm_tile_cols) == 0:
            use_panels = self.spmm_panels_mode == "1" and csr.n_cols < (1 << 24)
        if use_panels:
            need = self.lib.query("gnngp_spmm_workspace_bytes", csr.n_cols, k)
            ws = self.workspace(need + 256, B.device, "spmm")
            wptr = (ws.data_ptr() + 255) // 256 * 256
        elif k >= 32 and (ldb % 4 != 0 or B.data_ptr() % 16 != 0):
            # row-major gather: one streaming copy into a 16-byte-aligned, padded buffer so the gathers are 128-bit
            padded = self.empty(B.shape[0], k, B.device)
            padded.copy_(B)
            B, ldb = padded, self._ld(padded)
        short_rows = (not use_panels and self.spmm_flags & 4 and csr.nnz < 64 * csr.n_rows and k >= 128
                      and int(self.spmm_tile_cols) in (0, 128) and ldb % 4 == 0 and B.data_ptr() % 16 == 0)
        self.spmm_kernel_by_k[int(k)] = ("spmm_panel_kernel" if use_panels else
                                         "spmm_short_rows_kernel" if short_rows else "spmm_csr_kernel")
        self.lib.call("gnngp_spmm_csr_f32", csr.rowptr.data_ptr(), csr.colidx.data_ptr(), csr.vals.data_ptr(),
                      csr.n_rows, csr.n_cols, csr.nnz, B.data_ptr(), ldb, k, float(alpha), out.data_ptr(),
                      self._ld(out), int(self.spmm_tile_cols), int(self.spmm_nnz_per_warp), wptr, need, self._stream())
        return out

    def panels_from_peers(self, peer_ptrs_dev: int, world: int, rows_per_rank: int, ld: int, n: int, k: int, device) -> torch.Tensor:
        """64-column panels of the [n x k] operand whose row blocks live in the peers' memory (NVLink pull)."""
        need = self.lib.query("gnngp_spmm_workspace_bytes", n, k)
        ws = self.workspace(need + 256, device, "spmm")
        wptr = (ws.data_ptr() + 255) // 256 * 256
        self.lib.call("gnngp_panelize_peers_f32", peer_ptrs_dev, int(world), int(rows_per_rank), int(ld), int(n), int(k), wptr,
                      self._stream())
        return ws, wptr

    def spmm_panels(self, csr: CsrMatrix, panels_ptr: int, k: int, alpha: float, out: torch.Tensor) -> torch.Tensor:
        self._check(out, "out", dims=2)
        if out.shape != (csr.n_rows, k) or (k > 1 and out.stride(1) != 1):
            raise RuntimeError("spmm_panels: out must be a [%d x %d] row-major view" % (csr.n_rows, k))
        self.spmm_kernel_by_k[int(k)] = "spmm_panel_kernel"
        self.lib.call("gnngp_spmm_panels_f32", csr.rowptr.data_ptr(), csr.colidx.data_ptr(), csr.vals.data_ptr(), csr.n_rows,
                      csr.n_cols, csr.nnz, panels_ptr, k, float(alpha), out.data_ptr(), self._ld(out),
                      int(self.spmm_nnz_per_warp), self._stream())
        return out

    # ------------------------------------------------------------------ dense contractions
    def _gemm_ws(self, M: int, N: int, K: int, device):
        need = self.lib.query("gnngp_gemm_workspace_bytes", M, N, K)
        ws = self.workspace(need + 256, device, "gemm")
        return (ws.data_ptr() + 255) // 256 * 256, need

    def gemm_nt(self, A: torch.Tensor, B: torch.Tensor, alpha: float = 1.0, out: Optional[torch.Tensor] = None,
                beta: float = 0.0) -> torch.Tensor:
        """out = alpha * A @ B.T (+ beta * out)."""
        A, lda = self._mat(A, "A")
        B, ldb = self._mat(B, "B")
        M, K = A.shape
        N = B.shape[0]
        if B.shape[1] != K:
            raise RuntimeError("gemm_nt: inner dimensions differ (%d vs %d)" % (K, B.shape[1]))
        if out is None:
            out = self.empty(M, N, A.device)
        self._check(out, "out", dims=2)
        if out.shape != (M, N) or (N > 1 and out.stride(1) != 1):
            raise RuntimeError("gemm_nt: out must be a [%d x %d] row-major view" % (M, N))
        ws, nbytes = self._gemm_ws(M, N, K, A.device)
        self.lib.call("gnngp_gemm_nt_f32", A.data_ptr(), lda, B.data_ptr(), ldb, out.data_ptr(), self._ld(out),
                      M, N, K, float(alpha), float(beta), ws, nbytes, self._stream())
        return out

    def syrk(self, A: torch.Tensor) -> torch.Tensor:
        """A @ A.T with only the lower triangle guaranteed (upper entries are zero or mirrored)."""
        A, lda = self._mat(A, "A")
        n, K = A.shape
        out = self.zeros(n, n, A.device)
        ws, nbytes = self._gemm_ws(n, n, K, A.device)
        self.lib.call("gnngp_syrk_f32", A.data_ptr(), lda, out.data_ptr(), self._ld(out), n, K, ws, nbytes, self._stream())
        return out

    def gram_arccos(self, Q: torch.Tensor, L: torch.Tensor, s: torch.Tensor, t: torch.Tensor,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """E = J1-arccos((Q L^T) / s / t^T) * s * t^T — reference ``_ExxT_ReLU_Nystrom`` before the root."""
        Q, ldq = self._mat(Q, "Q")
        L, ldl = self._mat(L, "L")
        s, t = self._vec(s, "s"), self._vec(t, "t")
        n, m = Q.shape
        nl = L.shape[0]
        if L.shape[1] != m or s.numel() != n or t.numel() != nl:
            raise RuntimeError("gram_arccos: shape mismatch")
        if out is None:
            out = self.empty(n, nl, Q.device)
        ws, nbytes = self._gemm_ws(n, nl, m, Q.device)
        self.lib.call("gnngp_gram_arccos_f32", Q.data_ptr(), ldq, n, L.data_ptr(), ldl, nl, m, s.data_ptr(), t.data_ptr(),
                      out.data_ptr(), self._ld(out), ws, nbytes, self._stream())
        return out

    KINDS = {"linear": 0, "rbf": 1, "sigmoid": 2, "polynomial": 3, "arccos": 4}

    def initial_kernel(self, kind: str, X: torch.Tensor, R: torch.Tensor, gamma: float = 0.0, c: float = 0.0,
                       d: float = 1.0) -> torch.Tensor:
        X, ldx = self._mat(X, "X")
        R, ldr = self._mat(R, "R")
        n, f = X.shape
        r = R.shape[0]
        out = self.empty(n, r, X.device)
        ldo = self._ld(out)
        if kind == "laplacian":
            self.lib.call("gnngp_laplacian_kernel_f32", X.data_ptr(), ldx, n, R.data_ptr(), ldr, r, f, float(gamma),
                          out.data_ptr(), ldo, self._stream())
            return out
        xx = rr = None
        if kind == "rbf":
            xx = self.row_norms(X, sqrt=False)
            rr = self.row_norms(R, sqrt=False)
        ws, nbytes = self._gemm_ws(n, r, f, X.device)
        self.lib.call("gnngp_initial_kernel_f32", self.KINDS[kind], X.data_ptr(), ldx, n, R.data_ptr(), ldr, r, f,
                      xx.data_ptr() if xx is not None else None, rr.data_ptr() if rr is not None else None,
                      float(gamma), float(c), float(d), out.data_ptr(), ldo, ws, nbytes, self._stream())
        return out

    def nugget_coeff(self, temp: torch.Tensor, w: torch.Tensor, eps: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
        temp, ldt = self._mat(temp, "temp")
        w, eps = self._vec(w, "w"), self._vec(eps, "eps")
        self._check(d, "d")
        m, C = temp.shape
        G = eps.numel()
        S = self.empty(G * C, m, temp.device)
        self.lib.call("gnngp_nugget_coeff_f32", temp.data_ptr(), ldt, w.data_ptr(), eps.data_ptr(), d.data_ptr(), m, C, G,
                      S.data_ptr(), self._ld(S), self._stream())
        return S

    def nugget_sweep(self, Q: torch.Tensor, Bt: torch.Tensor, G: int, C: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        Q, ldq = self._mat(Q, "Q")
        Bt, ldb = self._mat(Bt, "Bt")
        n, m = Q.shape
        if Bt.shape != (G * C, m):
            raise RuntimeError("nugget_sweep: Bt must be [%d x %d]" % (G * C, m))
        if out is None:
            out = torch.empty((G, n, C), dtype=torch.float32, device=Q.device)
        ws, nbytes = self._gemm_ws(n, G * C, m, Q.device)
        self.lib.call("gnngp_nugget_sweep_f32", Q.data_ptr(), ldq, n, m, Bt.data_ptr(), ldb, G, C, out.data_ptr(), ws, nbytes,
                      self._stream())
        return out

    # ------------------------------------------------------------------ rows / columns
    def row_norms(self, Q: torch.Tensor, sqrt: bool = True) -> torch.Tensor:
        Q, ldq = self._mat(Q, "Q")
        out = torch.empty(_round_up(max(Q.shape[0], 1), 4), dtype=torch.float32, device=Q.device)[:Q.shape[0]]
        self.lib.call("gnngp_row_sumsq_f32", Q.data_ptr(), ldq, Q.shape[0], Q.shape[1], 1 if sqrt else 0, out.data_ptr(),
                      self._stream())
        return out

    def mean(self, v: torch.Tensor, stride: int = 1, count: Optional[int] = None) -> torch.Tensor:
        self._check(v, "v")
        n = int(v.numel() if count is None else count)
        out = torch.empty(1, dtype=torch.float32, device=v.device)
        self.lib.call("gnngp_mean_f32", v.data_ptr(), n, int(stride), out.data_ptr(), self._stream())
        return out

    def diag_mean(self, K: torch.Tensor) -> torch.Tensor:
        K, ldk = self._mat(K, "K")
        return self.mean(K, stride=ldk + 1, count=K.shape[0])

    def fill_col(self, M: torch.Tensor, col: int, value: float) -> None:
        self._check(M, "M", dims=2)
        view = M[:, col]
        self.lib.call("gnngp_fill_col_f32", view.data_ptr(), self._ld(M), M.shape[0], float(value), self._stream())

    def gather_rows(self, src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        src, lds = self._mat(src, "src")
        idx = self._vec(idx, "idx", torch.int64)
        out = self.empty(idx.numel(), src.shape[1], src.device)
        self.lib.call("gnngp_gather_rows_f32", src.data_ptr(), lds, idx.data_ptr(), idx.numel(), src.shape[1],
                      out.data_ptr(), self._ld(out), self._stream())
        return out

    def scatter_rows(self, dst: torch.Tensor, idx: torch.Tensor, src: torch.Tensor) -> None:
        src, lds = self._mat(src, "src")
        self._check(dst, "dst", dims=2)
        idx = self._vec(idx, "idx", torch.int64)
        if dst.shape[1] != src.shape[1] or (dst.shape[1] > 1 and dst.stride(1) != 1):
            raise RuntimeError("scatter_rows: column mismatch")
        self.lib.call("gnngp_scatter_rows_f32", src.data_ptr(), lds, idx.data_ptr(), idx.numel(), src.shape[1],
                      dst.data_ptr(), self._ld(dst), self._stream())

    def gather_rows_t(self, src: torch.Tensor, idx: Optional[torch.Tensor]) -> torch.Tensor:
        """out[c, r] = src[idx[r], c]; ``idx=None`` is a plain transpose."""
        src, lds = self._mat(src, "src")
        n_idx = src.shape[0] if idx is None else idx.numel()
        if idx is not None:
            idx = self._vec(idx, "idx", torch.int64)
        out = self.empty(src.shape[1], n_idx, src.device)
        self.lib.call("gnngp_gather_rows_transposed_f32", src.data_ptr(), lds, idx.data_ptr() if idx is not None else None,
                      n_idx, src.shape[1], out.data_ptr(), self._ld(out),
                      self._stream())
        return out

    def transpose(self, src: torch.Tensor) -> torch.Tensor:
        return self.gather_rows_t(src, None)

    def gather_cols(self, src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        src, lds = self._mat(src, "src")
        idx = self._vec(idx, "idx", torch.int64)
        out = self.empty(src.shape[0], idx.numel(), src.device)
        self.lib.call("gnngp_gather_cols_f32", src.data_ptr(), lds, src.shape[0], idx.data_ptr(), idx.numel(), out.data_ptr(),
                      self._ld(out), self._stream())
        return out

    def affine(self, out: torch.Tensor, x: torch.Tensor, y: Optional[torch.Tensor], a: float, b: float, c: float) -> torch.Tensor:
        """out = a*x + b*y + c (elementwise, in place allowed)."""
        for name, t in (("out", out), ("x", x)) + ((("y", y),) if y is not None else ()):
            self._check(t, name, dims=2)
            if t.shape != out.shape or (t.shape[1] > 1 and t.stride(1) != 1):
                raise RuntimeError("affine: %s must be a row-major view of the output shape" % name)
        n, m = out.shape
        self.lib.call("gnngp_affine_f32", out.data_ptr(), self._ld(out), x.data_ptr(), self._ld(x),
                      y.data_ptr() if y is not None else None, self._ld(y) if y is not None else 0, n, m,
                      float(a), float(b), float(c), self._stream())
        return out

    # ------------------------------------------------------------------ fp64 dense operators (Krylov Nystrom root)
    @staticmethod
    def _f64_view(t: torch.Tensor, name: str) -> torch.Tensor:
        if not t.is_cuda or t.dtype != torch.float64 or t.dim() != 2:
            raise RuntimeError("gnngp_b200: %s must be a 2-D CUDA float64 tensor (got %s %s on %s)"
                               % (name, tuple(t.shape), t.dtype, t.device))
        return t

    def sym_f64(self, kmm: torch.Tensor) -> torch.Tensor:
        """Full symmetric fp64 copy of the LOWER triangle of an fp32 block."""
        kmm, ldk = self._mat(kmm, "kmm")
        n = kmm.shape[0]
        if kmm.shape[1] != n:
            raise RuntimeError("sym_f64: square block expected")
        ld = n + (n & 1)                                             # even pitch: the fp64 GEMM's aligned fast path
        out = torch.empty((n, ld), dtype=torch.float64, device=kmm.device)[:, :n]
        self.lib.call("gnngp_sym_f64_from_f32", kmm.data_ptr(), ldk, n, out.data_ptr(), ld, self._stream())
        return out

    def dgemm(self, A: torch.Tensor, B: torch.Tensor, out: torch.Tensor, alpha: float = 1.0, beta: float = 0.0,
              lower: bool = False) -> torch.Tensor:
        """``out = alpha A @ B + beta out`` in fp64 on the DMMA path; ``A`` / ``B`` may be transposed or column-sliced
        views (one unit stride each), ``out`` a row-major view.  ``lower``: only tiles touching the lower triangle."""
        self._f64_view(A, "A"), self._f64_view(B, "B"), self._f64_view(out, "out")
        m, k = A.shape
        k2, n = B.shape
        if k2 != k or tuple(out.shape) != (m, n) or (n > 1 and out.stride(1) != 1):
            raise RuntimeError("dgemm: shapes %s @ %s -> %s" % (tuple(A.shape), tuple(B.shape), tuple(out.shape)))
        self.lib.call("gnngp_dgemm_f64", m, n, k, A.data_ptr(), A.stride(0), A.stride(1), B.data_ptr(), B.stride(0),
                      B.stride(1), float(alpha), float(beta), out.data_ptr(), out.stride(0) if m > 1 else max(n, out.stride(0)),
                      1 if lower else 0, self._stream())
        return out

    def potrf_f64(self, A: torch.Tensor, info: torch.Tensor) -> torch.Tensor:
        """In-place lower Cholesky factor of a symmetric fp64 matrix (strict upper triangle zeroed); ``info`` (device
        int32) counts blocks with a non-positive pivot.  No host sync."""
        self._f64_view(A, "A")
        n = A.shape[0]
        if A.shape[1] != n or A.stride(1) != 1:
            raise RuntimeError("potrf_f64: square row-major matrix expected")
        need = self.lib.query("gnngp_potrf_f64_workspace_bytes", n)
        ws = self.workspace(need + 256, A.device, "potrf")
        wptr = (ws.data_ptr() + 255) // 256 * 256
        self.lib.call("gnngp_potrf_f64", A.data_ptr(), A.stride(0), n, info.data_ptr(), wptr, need, self._stream())
        return A

    def potrf_inv_block(self, G: torch.Tensor, info: torch.Tensor) -> torch.Tensor:
        """One block of at most 128 rows: Cholesky factor in place (lower triangle) and the returned ``L⁻¹``."""
        self._f64_view(G, "G")
        nb = G.shape[0]
        ldi = nb + (nb & 1)                                          # even pitch (aligned GEMM fast path)
        linv = torch.empty((nb, ldi), dtype=torch.float64, device=G.device)[:, :nb]
        self.lib.call("gnngp_potrf_inv_block_f64", G.data_ptr(), G.stride(0), nb, linv.data_ptr(), ldi, info.data_ptr(),
                      self._stream())
        return linv

    def qr_rotate(self, X: torch.Tensor, Z: torch.Tensor) -> torch.Tensor:
        """``X <- X Q_H`` in place, ``Q_H`` the orthogonal factor of the Householder QR of ``Z`` [n x kp] (kp a multiple of
        128; ``Z`` is overwritten)."""
        self._f64_view(X, "X"), self._f64_view(Z, "Z")
        rows, n = X.shape
        n2, kp = Z.shape
        if n2 != n or X.stride(1) != 1 or Z.stride(1) != 1:
            raise RuntimeError("qr_rotate: X [rows x n] and Z [n x kp] row-major expected")
        need = self.lib.query("gnngp_qr_rotate_workspace_bytes", rows, n, kp)
        ws = self.workspace(need + 256, X.device, "qr_rotate")
        wptr = (ws.data_ptr() + 255) // 256 * 256
        self.lib.call("gnngp_qr_rotate_f64", X.data_ptr(), X.stride(0), rows, Z.data_ptr(), Z.stride(0), n, kp, wptr, need,
                      self._stream())
        return X

    def krylov_root(self, kmm: torch.Tensor, stats: Optional[dict] = None, comm=None):
        """The Nystrom root of a landmark block by ``krylov_root.root_factors`` (an operator so that it shows up as ONE
        stage in the bench's timing table); ``None`` when the route hands the block back."""
        from . import krylov_root as route
        return route.root_factors(self, kmm, stats, comm=comm)

    def krylov_fit(self, gram: torch.Tensor, qty_t: torch.Tensor, shifts: torch.Tensor, stats: Optional[dict] = None, comm=None):
        """All nuggets' coefficient blocks by ``krylov_fit.coefficients`` (one timed stage); ``None`` = not converged."""
        from . import krylov_fit as route
        return route.coefficients(self, gram, qty_t, shifts, stats, comm=comm)

    def band_solve(self, H: torch.Tensor, rhs_t: torch.Tensor, shifts: torch.Tensor):
        """``(band(H) + shifts[g] I) x = rhs`` for every shift (the band |i - j| <= 128 of the lower triangle of the fp64
        matrix ``H`` is used): fp64 [(G*C) x dim] solutions (vectors as rows) and a device int32 counting non-positive
        pivots.  ``rhs_t`` fp64 [C x dim], C <= 48."""
        self._f64_view(H, "H"), self._f64_view(rhs_t, "rhs_t")
        dim, c, g = H.shape[0], rhs_t.shape[0], shifts.numel()
        if H.shape[1] != dim or rhs_t.shape[1] != dim or H.stride(1) != 1 or rhs_t.stride(1) != 1:
            raise RuntimeError("band_solve: H [dim x dim] and rhs_t [C x dim] row-major expected")
        shifts = shifts.to(torch.float64).contiguous()
        sol = torch.empty((g * c, dim), dtype=torch.float64, device=H.device)
        info = torch.zeros(1, dtype=torch.int32, device=H.device)
        need = self.lib.query("gnngp_band_solve_workspace_bytes", dim, c, g)
        ws = self.workspace(need + 256, H.device, "band_solve")
        wptr = (ws.data_ptr() + 255) // 256 * 256
        self.lib.call("gnngp_band_solve_f64", H.data_ptr(), H.stride(0), dim, rhs_t.data_ptr(), rhs_t.stride(0), c,
                      shifts.data_ptr(), g, sol.data_ptr(), dim, info.data_ptr(), wptr, need, self._stream())
        return sol, info

    def eigh_small_f64(self, H: torch.Tensor):
        """Rayleigh-Ritz step of the Krylov root: eigendecomposition of the SMALL projected matrix (dimension ~0.11 n;
        cuSOLVER through torch, 1/200 of the flops of the eigendecomposition it replaces)."""
        return torch.linalg.eigh(H)

    def residual_sumsq(self, R: torch.Tensor, V: torch.Tensor, theta: torch.Tensor) -> torch.Tensor:
        self._f64_view(R, "R"), self._f64_view(V, "V")
        n, k = R.shape
        theta = theta.contiguous()
        out = torch.empty(k, dtype=torch.float64, device=R.device)
        self.lib.call("gnngp_residual_sumsq_f64", R.data_ptr(), R.stride(0), V.data_ptr(), V.stride(0), theta.data_ptr(), n, k,
                      out.data_ptr(), self._stream())
        return out

    def transpose_scale_f64(self, A: torch.Tensor, alpha: float) -> torch.Tensor:
        self._f64_view(A, "A")
        rows, cols = A.shape
        out = torch.empty((cols, rows), dtype=torch.float64, device=A.device)
        self.lib.call("gnngp_transpose_scale_f64", A.data_ptr(), A.stride(0), rows, cols, float(alpha), out.data_ptr(), rows,
                      self._stream())
        return out

    def scale_cols_f64(self, V: torch.Tensor, d: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        self._f64_view(V, "V")
        rows, cols = V.shape
        d = d.contiguous()
        if out is None:
            out = torch.empty((rows, cols), dtype=torch.float64, device=V.device)
        self.lib.call("gnngp_scale_cols_f64", V.data_ptr(), V.stride(0), rows, cols, d.data_ptr(), out.data_ptr(), out.stride(0),
                      self._stream())
        return out

    def to_f32(self, A: torch.Tensor) -> torch.Tensor:
        self._f64_view(A, "A")
        rows, cols = A.shape
        out = self.empty(rows, cols, A.device)
        self.lib.call("gnngp_f64_to_f32", A.data_ptr(), A.stride(0), rows, cols, out.data_ptr(), self._ld(out), self._stream())
        return out

    # ------------------------------------------------------------------ Nystrom root pieces
    # dtype the eigensolver runs in.  Measured on B200 (profiles/r1/accuracy_probe.json): cuSOLVER's fp32
    # syevd/syevj through torch loses ~50x accuracy against LAPACK's ssyevd on the clamped Nystrom root
    # (Q Q^T rel. error 6e-5 vs 1e-6 for the reference's own fp32 run); in fp64 it is 2e-7.
    eigh_dtype = torch.float64

    def eigh(self, M: torch.Tensor):
        """Symmetric eigendecomposition (ascending, lower triangle like torch's default), cuSOLVER through
        torch — the one library call kept on the path (SURVEY.md §7 hard part 1).  Returns fp32."""
        self._check(M, "M", dims=2)
        src = M if M.is_contiguous() else M.contiguous()
        if self.eigh_dtype == torch.float32:
            return torch.linalg.eigh(src)
        evals, evecs = torch.linalg.eigh(src.to(self.eigh_dtype))
        return evals.to(torch.float32), evecs.to(torch.float32)

    def shifted_solves(self, gram: torch.Tensor, rhs_t: torch.Tensor, shifts: torch.Tensor):
        """``X_i = (gram + shifts[i] I)⁻¹ rhs`` for every shift, by fp64 Cholesky (cuSOLVER potrf / potrs through
        torch — library calls like ``eigh``, used by the nugget-parallel fit of the sharded mode).  ``gram`` is
        fp32 with a meaningful LOWER triangle (like ``syrk``'s output), ``rhs_t`` is [C x m].  Returns the
        solutions transposed and stacked, fp32 [(len(shifts)*C) x m], and the number of factorisations that
        reported a non-positive pivot (device scalar; no host sync here)."""
        self._check(gram, "gram", dims=2)
        self._check(rhs_t, "rhs_t", dims=2)
        return _shifted_solves(self, gram, rhs_t, shifts)

    def band_fit(self, gram: torch.Tensor, rhs_t: torch.Tensor, shifts: torch.Tensor):
        """``(gram + shifts[g] I)⁻¹ rhs`` for every shift WITHOUT an eigendecomposition (``csrc/bandfit.cu``): fp64 band
        reduction of the Gram matrix with block reflectors in GEMM form, then one banded Cholesky and two substitutions
        per shift.  ``gram`` fp32 [m x m] with a valid LOWER triangle (``syrk``'s output), ``rhs_t`` fp32 [C x m],
        ``shifts`` fp64 [G] on the device.  Returns (coefficients fp32 [(G*C) x m] — the B operand of ``nugget_sweep`` —
        and a device int32 counting shifts whose factorisation met a non-positive pivot).  No host sync."""
        gram, ldg = self._mat(gram, "gram")
        rhs_t, ldr = self._mat(rhs_t, "rhs_t")
        self._check(shifts, "shifts", dtype=torch.float64, dims=1)
        shifts = shifts.contiguous()
        m, c, g = gram.shape[0], rhs_t.shape[0], shifts.numel()
        if gram.shape[1] != m or rhs_t.shape[1] != m:
            raise RuntimeError("band_fit: gram must be [m x m] and rhs_t [C x m]")
        out = self.empty(g * c, m, gram.device)
        info = torch.zeros(1, dtype=torch.int32, device=gram.device)
        need = self.lib.query("gnngp_band_fit_workspace_bytes", m, c, g)
        ws = self.workspace(need + 256, gram.device, "bandfit")
        wptr = (ws.data_ptr() + 255) // 256 * 256
        self.lib.call("gnngp_band_fit_f64", gram.data_ptr(), ldg, m, rhs_t.data_ptr(), ldr, c, shifts.data_ptr(), g,
                      out.data_ptr(), self._ld(out), info.data_ptr(), wptr, need, self._stream())
        return out, info

    def potrs_batched(self, factors: torch.Tensor, rhs: torch.Tensor) -> torch.Tensor:
        """``X_b = (L_b L_bᵀ)⁻¹ rhs`` for a stack of fp64 lower Cholesky factors [b, m, m] and one fp64
        right-hand-side block [m, c], in one launch of ``potrs_batched_kernel``.  Returns fp64 [b, m, c]."""
        self._check(factors, "factors", dtype=torch.float64, dims=3)
        self._check(rhs, "rhs", dtype=torch.float64, dims=2)
        b, m, m2 = factors.shape
        if m != m2 or rhs.shape[0] != m:
            raise ValueError("potrs_batched: factors %s do not match rhs %s" % (tuple(factors.shape), tuple(rhs.shape)))
        factors = factors if factors.is_contiguous() else factors.contiguous()
        rhs = rhs if rhs.is_contiguous() else rhs.contiguous()
        c = rhs.shape[1]
        ldx = (c + 7) // 8 * 8
        out = torch.empty((b, m, ldx), dtype=torch.float64, device=factors.device)
        self.lib.call("gnngp_potrs_batched_f64", factors.data_ptr(), m, m * m, rhs.data_ptr(), c, 0, out.data_ptr(), ldx,
                      m * ldx, m, c, b, self._stream())
        return out[:, :, :c]

    def nystrom_spectrum(self, D: torch.Tensor):
        D = self._vec(D, "D")
        root = torch.empty_like(D)
        inv = torch.empty_like(D)
        self.lib.call("gnngp_nystrom_spectrum_f32", D.data_ptr(), D.numel(), root.data_ptr(), inv.data_ptr(), self._stream())
        return root, inv

    def scale_cols(self, V: torch.Tensor, d: torch.Tensor, transpose: bool = False) -> torch.Tensor:
        V, ldv = self._mat(V, "V")
        d = self._vec(d, "d")
        n, m = V.shape
        out = self.empty(m, n, V.device) if transpose else self.empty(n, m, V.device)
        self.lib.call("gnngp_scale_cols_f32", V.data_ptr(), ldv, n, m, d.data_ptr(), 1 if transpose else 0, out.data_ptr(),
                      self._ld(out), self._stream())
        return out

    def arccos_dense(self, K: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        K, ldk = self._mat(K, "K")
        n = K.shape[0]
        if out is None:
            out = self.empty(n, n, K.device)
        s = torch.empty(n, dtype=torch.float32, device=K.device)
        self.lib.call("gnngp_arccos_dense_f32", K.data_ptr(), ldk, n, out.data_ptr(), self._ld(out),
                      s.data_ptr(), self._stream())
        return out

    # ------------------------------------------------------------------ targets / metrics
    def onehot_t(self, y: torch.Tensor, idx: torch.Tensor, C: int) -> torch.Tensor:
        y = self._vec(y, "y", torch.int64)
        idx = self._vec(idx, "idx", torch.int64)
        out = self.empty(C, idx.numel(), y.device)
        self.lib.call("gnngp_onehot_t_f32", y.data_ptr(), idx.data_ptr(), idx.numel(), C, out.data_ptr(),
                      self._ld(out), self._stream())
        return out

    def argmax_count(self, fitted: torch.Tensor, y: torch.Tensor, membership: torch.Tensor) -> torch.Tensor:
        self._check(fitted, "fitted", dims=3)
        fitted = fitted.contiguous()
        y = self._vec(y, "y", torch.int64)
        self._check(membership, "membership", torch.uint8, 1)
        G, n, C = fitted.shape
        counts = torch.empty((G, 8), dtype=torch.int32, device=fitted.device)
        self.lib.call("gnngp_argmax_count", fitted.data_ptr(), y.data_ptr(), membership.data_ptr(), G, n, C, counts.data_ptr(),
                      self._stream())
        return counts

    def sqerr_sum(self, fitted: torch.Tensor, y: torch.Tensor, membership: torch.Tensor) -> torch.Tensor:
        self._check(fitted, "fitted")
        fitted = fitted.contiguous()
        y = self._check(y, "y").contiguous()
        G, n = fitted.shape[0], fitted.shape[1]
        k = int(fitted[0, 0].numel())
        sse = torch.empty((G, 8), dtype=torch.float64, device=fitted.device)
        self.lib.call("gnngp_sqerr_sum", fitted.data_ptr(), y.data_ptr(), membership.data_ptr(), G, n, k, sse.data_ptr(),
                      self._stream())
        return sse


def _device_of(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.device):
            if a.type == "cuda":
                return a
        elif isinstance(a, torch.Tensor):
            if a.is_cuda:
                return a.device
        elif isinstance(a, CsrMatrix):
            return a.colidx.device if a.colidx.is_cuda else None
        elif isinstance(a, ChunkedCsr):
            return a.in_range.device
    return None


def _guarded(fn):
    """Run an operator with the CUDA device of its first tensor argument current.  The C ABI launches on the
    runtime's current device and `_stream()` returns that device's current stream, so a model built with
    ``device=cuda:1`` while device 0 is current would otherwise launch on GPU 0 against GPU-1 pointers (the
    reference honours ``--device`` because torch ops guard themselves, src/main.py:68)."""
    import functools

    @functools.wraps(fn)
    def run(self, *args, **kwargs):
        dev = _device_of(args, kwargs)
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(self, *args, **kwargs)
        with torch.cuda.device(dev):
            return fn(self, *args, **kwargs)
    return run


for _name, _fn in list(vars(CudaOps).items()):
    if callable(_fn) and not _name.startswith("_") and not isinstance(_fn, (staticmethod, classmethod)) and _name not in (
            "workspace", "release_workspace", "side_stream", "launch_count", "tc_launch_count", "set_gemm_engine", "clear_cache"):
        setattr(CudaOps, _name, _guarded(_fn))
del _name, _fn


_OPS = None


def default_ops() -> CudaOps:
    """The process-wide operator set.  Raises if the CUDA library is not built."""
    global _OPS
    if _OPS is None:
        _OPS = CudaOps()
    return _OPS


_SOLVER_STREAMS = {}


def _solver_streams(device: torch.device, wanted: int):
    """Side streams for the independent per-nugget factorisations (cached per device)."""
    have = _SOLVER_STREAMS.setdefault(device.index, [])
    while len(have) < wanted:
        have.append(torch.cuda.Stream(device=device))
    return have[:wanted]


def _shifted_solves(ops, gram: torch.Tensor, rhs_t: torch.Tensor, shifts: torch.Tensor):
    """potrf per shift (cuSOLVER through torch) then ONE batched substitution launch.  One potrf of a few
    thousand rows is a chain of small kernels (1.4 ms at m = 3 025 = 6 TFLOP/s of fp64); the shifts are
    independent, so they are issued round-robin on ``GNNGP_SOLVER_STREAMS`` (default 4) side streams; the
    measured gain is modest (13 shifts at m = 3 025: 23.7 -> 19.2 ms, `profiles/r1/call25_potrs_probe.json`)
    because the host launch rate of the ~150 small kernels per potrf is the limit."""
    import os
    m, c = gram.shape[0], rhs_t.shape[0]
    dev = gram.device
    low = torch.tril(gram).to(torch.float64)
    sym = low + torch.tril(low, -1).T
    del low
    rhs = rhs_t.to(torch.float64).T.contiguous()                     # [m x C]
    count = int(shifts.numel())
    out = ops.empty(count * c, m, dev)
    bad = torch.zeros((), dtype=torch.int64, device=dev)
    # factors kept at once: <= 40 GB (all 13 of a rank at m = 15 356; one substitution launch is bound by the
    # latency of a CTA walking a factor, not by how many factors it covers, so fewer, larger launches win)
    held = max(1, min(count, int(40e9) // max(1, m * m * 8)))
    lanes = max(1, min(int(os.environ.get("GNNGP_SOLVER_STREAMS", "4")), held))
    main = torch.cuda.current_stream(dev)
    side = _solver_streams(dev, lanes) if lanes > 1 else [main]
    for s0 in range(0, count, held):
        cnt = min(held, count - s0)
        stack = torch.empty((cnt, m, m), dtype=torch.float64, device=dev)
        infos = []
        ready = torch.cuda.Event()
        ready.record(main)
        for i in range(cnt):
            stream = side[i % len(side)]
            if stream is not main:
                stream.wait_event(ready)
            with torch.cuda.stream(stream):
                shifted = sym.clone()
                shifted.diagonal().add_(shifts[s0 + i])
                factor, info = torch.linalg.cholesky_ex(shifted)     # cuSOLVER potrf
                stack[i].copy_(factor)
                infos.append(info)
                del shifted, factor
        for stream in side:
            if stream is not main:
                main.wait_stream(stream)
        bad += torch.stack([(info != 0).to(torch.int64).sum() for info in infos]).sum()
        sol = ops.potrs_batched(stack, rhs)                           # [cnt, m, C] fp64, one launch for all factors
        out[s0 * c:(s0 + cnt) * c] = sol.permute(0, 2, 1).reshape(cnt * c, m).to(torch.float32)
        del stack, sol
    return out, bad
