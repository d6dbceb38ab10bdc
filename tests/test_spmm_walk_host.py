"""The work-item walk of the SpMM kernels (``csrc/spmm.cu``), restated on the CPU to pin two invariants that the CUDA
code relies on and that no single kernel can check by itself:

* ``zero_cut_rows_kernel`` zero-fills exactly the rows the product kernels accumulate with atomics (cut by a border of
  the nnz-balanced split) or never visit (empty): its predicate ``rs / chunk == (re - 1) / chunk`` must agree with the
  kernels' ``whole = rstart >= p0 && rend <= p1`` for every row and every chunk size: a row that is ADDED to must have been
  zeroed, a row that is STORED must be stored exactly once, and no row may be left untouched;
* ``spmm_short_rows_kernel`` consumes nonzeros in fixed groups across row boundaries and finds row ends in a cache of the
  next 32 row pointers: every row must be flushed once per work item that touches it, with the right ``whole`` flag, also
  across cache refills and runs of empty rows.

This is a model of the control flow only (numpy, float64); the kernels themselves are checked on the GPU against torch
(``tests/test_gpu_ops.py``).
"""
import numpy as np
import pytest

GROUP = 4          # kShortUnroll
CACHE = 32         # row ends per register cache fill


def _first_row(rowptr, p0):
    lo, hi = 0, len(rowptr) - 2
    while lo < hi:
        mid = (lo + hi + 1) >> 1
        if rowptr[mid] <= p0:
            lo = mid
        else:
            hi = mid - 1
    return lo


def _zero_fill(rowptr, chunk, k, zero_cut_only):
    n = len(rowptr) - 1
    c = np.full((n, k), np.nan)
    for r in range(n):
        rs, re = rowptr[r], rowptr[r + 1]
        whole = re > rs and rs // chunk == (re - 1) // chunk
        if not zero_cut_only or not whole:
            c[r] = 0.0
    return c


class _Sink(object):
    def __init__(self, c):
        self.c, self.stored = c, set()

    def flush(self, row, acc, whole):
        if whole:
            assert row not in self.stored, "row %d stored twice" % row
            self.stored.add(row)
            self.c[row] = acc
        else:
            assert not np.isnan(self.c[row]).any(), "row %d is accumulated but was not zero-filled" % row
            self.c[row] += acc


def _walk_rows(rowptr, col, val, b, chunk, sink):
    """spmm_csr_kernel / spmm_panel_kernel: one row segment at a time."""
    nnz = int(rowptr[-1])
    for piece in range(-(-nnz // chunk)):
        p0 = piece * chunk
        p1 = min(p0 + chunk, nnz)
        row, p = _first_row(rowptr, p0), p0
        while p < p1:
            while rowptr[row + 1] <= p:
                row += 1
            rstart, rend = rowptr[row], rowptr[row + 1]
            seg_end = min(rend, p1)
            acc = (val[p:seg_end, None] * b[col[p:seg_end]]).sum(0)
            sink.flush(row, acc, rstart >= p0 and rend <= p1)
            p = seg_end
            if seg_end == rend:
                row += 1


def _walk_groups(rowptr, col, val, b, chunk, sink):
    """spmm_short_rows_kernel: groups of GROUP nonzeros, row ends from a cache of CACHE row pointers."""
    n, nnz = len(rowptr) - 1, int(rowptr[-1])
    for piece in range(-(-nnz // chunk)):
        p0 = piece * chunk
        length = min(p0 + chunk, nnz) - p0
        row = _first_row(rowptr, p0)
        starts_here = rowptr[row] >= p0

        def load_ends(base):
            return [int(min(rowptr[min(base + 1 + lane, n)] - p0, 1 << 30)) for lane in range(CACHE)]

        ends, cached = load_ends(row), 0
        rend = ends[0]
        acc, has = np.zeros(b.shape[1]), False
        for q in range(0, length, 32):
            cnt = min(32, length - q)
            for j in range(0, cnt, GROUP):
                for u in range(GROUP):
                    if j + u >= cnt:
                        continue
                    pos = q + j + u
                    while pos == rend:
                        if has:
                            sink.flush(row, acc, starts_here)
                        acc, has, starts_here = np.zeros(b.shape[1]), False, True
                        row += 1
                        cached += 1
                        if cached == CACHE:
                            ends, cached = load_ends(row), 0
                        rend = ends[cached]
                    acc = acc + val[p0 + pos] * b[col[p0 + pos]]
                    has = True
        if has:
            sink.flush(row, acc, starts_here and rend <= length)


def _graph(rng, n, mean_degree, hub, empty_share):
    counts = rng.poisson(mean_degree, n)
    if hub:
        counts[rng.integers(0, n)] += hub
    counts[rng.random(n) < empty_share] = 0
    rowptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    nnz = int(rowptr[-1])
    return rowptr, rng.integers(0, n, nnz), rng.standard_normal(nnz), counts


@pytest.mark.parametrize("walk", [_walk_rows, _walk_groups])
@pytest.mark.parametrize("zero_cut_only", [False, True])
def test_every_row_is_stored_once_or_zeroed_and_accumulated(walk, zero_cut_only):
    rng = np.random.default_rng(7)
    for trial in range(120):
        n = int(rng.integers(1, 150))
        chunk = int(rng.choice([32, 64, 100, 512]))
        rowptr, col, val, counts = _graph(rng, n, rng.choice([0.6, 4.0, 15.0]), int(rng.integers(0, 700)) * int(rng.random() < 0.4),
                                         rng.choice([0.0, 0.25, 0.6]))
        if rowptr[-1] == 0:
            continue
        b = rng.standard_normal((n, 3))
        want = np.zeros((n, 3))
        np.add.at(want, np.repeat(np.arange(n), counts), val[:, None] * b[col])
        sink = _Sink(_zero_fill(rowptr, chunk, 3, zero_cut_only))
        walk(rowptr, col, val, b, chunk, sink)
        assert not np.isnan(sink.c).any(), "a row was neither zero-filled nor stored (trial %d)" % trial
        assert np.allclose(sink.c, want, atol=1e-12), trial


def test_short_row_walk_refills_its_row_end_cache():
    """Many rows per work item (one nonzero each, runs of empty rows in between): the cache of 32 row ends is refilled
    several times inside one item."""
    rng = np.random.default_rng(3)
    n = 700
    counts = (rng.random(n) < 0.5).astype(np.int64)
    counts[100:180] = 0
    rowptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    nnz = int(rowptr[-1])
    col, val, b = rng.integers(0, n, nnz), rng.standard_normal(nnz), rng.standard_normal((n, 2))
    want = np.zeros((n, 2))
    np.add.at(want, np.repeat(np.arange(n), counts), val[:, None] * b[col])
    for zero_cut_only in (False, True):
        sink = _Sink(_zero_fill(rowptr, 512, 2, zero_cut_only))
        _walk_groups(rowptr, col, val, b, 512, sink)
        assert np.allclose(sink.c, want, atol=1e-12)
