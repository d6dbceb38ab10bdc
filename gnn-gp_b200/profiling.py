"""CUDA-event timing of the operator set, on the stream the kernels are launched on.

``OpTimer(ops)`` wraps the operator methods so that every call is bracketed by two events recorded on
torch's current stream (which is the stream the C ABI launches on).  Durations are read after a
synchronize; nothing here runs under a profiler.  Used by ``bench.py`` for the per-stage breakdown and
for the roofline figure of the dominant kernel."""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List

import torch

TIMED = ("csr_from_coo", "spmm", "spmm_panels", "panels_from_peers", "gemm_nt", "syrk", "restrict_columns", "gram_arccos", "initial_kernel", "nugget_coeff", "nugget_sweep", "row_norms",
         "gather_rows", "scatter_rows", "gather_rows_t", "gather_cols", "affine", "eigh", "nystrom_spectrum",
         "scale_cols", "arccos_dense", "onehot_t", "argmax_count", "sqerr_sum", "mean", "fill_col", "shifted_solves", "band_fit",
         "krylov_root", "krylov_fit")


class OpTimer(object):
    def __init__(self, ops):
        self.ops = ops
        self.records: List[tuple] = []
        self._saved = {}
        self.enabled = False

    def _wrap(self, name):
        fn = getattr(self.ops, name)

        def timed(*args, **kwargs):
            if not self.enabled:
                return fn(*args, **kwargs)
            start = torch.cuda.Event(enable_timing=True)
            stop = torch.cuda.Event(enable_timing=True)
            start.record()
            out = fn(*args, **kwargs)
            stop.record()
            info = None
            if name == "spmm":
                info = (args[0].n_rows, args[0].nnz, int(args[1].shape[1]))
            elif name == "spmm_panels":
                info = (args[0].n_rows, args[0].nnz, int(args[2]))
            elif name in ("gemm_nt", "gram_arccos"):
                info = (int(args[0].shape[0]), int(args[1].shape[0]), int(args[0].shape[1]))
            elif name == "nugget_sweep":
                info = (int(args[0].shape[0]), int(args[1].shape[0]), int(args[0].shape[1]))
            elif name == "syrk":
                info = (int(args[0].shape[0]), int(args[0].shape[0]), int(args[0].shape[1]))
            elif name == "eigh":
                info = (int(args[0].shape[0]),)
            self.records.append((name, info, start, stop))
            return out
        return timed

    def install(self):
        for name in TIMED:
            if hasattr(self.ops, name):
                self._saved[name] = getattr(self.ops, name)
                setattr(self.ops, name, self._wrap(name))
        return self

    def remove(self):
        for name in self._saved:
            try:
                delattr(self.ops, name)       # instance attribute shadows the class method
            except AttributeError:
                pass
        self._saved.clear()

    def reset(self):
        self.records = []

    def summary(self) -> Dict[str, dict]:
        torch.cuda.synchronize()
        agg = defaultdict(lambda: dict(calls=0, ms=0.0))
        for name, info, start, stop in self.records:
            agg[name]["calls"] += 1
            agg[name]["ms"] += start.elapsed_time(stop)
        return dict(agg)

    def calls(self, name) -> List[tuple]:
        torch.cuda.synchronize()
        return [(info, start.elapsed_time(stop)) for n, info, start, stop in self.records if n == name]
