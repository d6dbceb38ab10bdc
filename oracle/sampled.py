"""CPU baseline of the GCNGP-X fit+predict path on a BOUNDED SAMPLE — TEST / BENCH INFRASTRUCTURE ONLY.

The Reddit-shape workload takes minutes to hours on host cores (283 s at fraction 50 on
8 cores, out of host memory at fraction 10), so ``bench.py``'s ``cpu_baseline`` and ``--impl reference``
legs time every stage of the reference algorithm (``src/compute.py:139-146`` with ``120-127`` and
``6-17``; ``src/predict.py:36-63, 66-90``) on a row sample and scale row-separable stages by the
inverse sampling rate.  Every stage is separable over node rows except the two symmetric
eigendecompositions, which are run at full size.  The arithmetic is the oracle's
(``oracle/gnngp_oracle.py``: numpy / LAPACK for dense stages, the OpenMP CSR kernel of
``oracle/spmm_ref.c`` for the sparse products), in fp32 like the reference.
"""
from __future__ import annotations

import time
from typing import Dict

import numpy as np

from . import gnngp_oracle as orc


class RowSample(object):
    """CSR rows of the adjacency for a sample of nodes (all columns kept)."""

    def __init__(self, indptr, indices, data, n_cols):
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.data = np.ascontiguousarray(data, dtype=np.float32)
        self.n_rows = self.indptr.shape[0] - 1
        self.n_cols = n_cols

    def matmul(self, dense: np.ndarray) -> np.ndarray:
        adj = orc.Adjacency.__new__(orc.Adjacency)
        adj.n = self.n_rows
        adj.indptr, adj.indices, adj.data = self.indptr, self.indices, self.data
        import scipy.sparse as sp
        adj._scipy = sp.csr_matrix((self.data, self.indices, self.indptr), shape=(self.n_rows, self.n_cols))
        return adj.matmul(dense)


def _timed(fn):
    t0 = time.perf_counter()
    out = fn()
    return out, time.perf_counter() - t0


def estimate_gcn_nystrom_seconds(sample: RowSample, n: int, x: np.ndarray, land_rows_x: np.ndarray, n_land: int,
                                 n_train: int, n_classes: int, n_nugget: int, layers: int, sigma_b: float,
                                 sigma_w: float, dense_rows: int, seed: int = 0) -> Dict[str, float]:
    """Seconds of one GCNGP-X ``predict`` at full size, estimated from a sample.

    ``sample``     adjacency rows of ``sample.n_rows`` sampled nodes (scaled by n / n_rows),
    ``x``          full feature matrix [n, F] (gathered by the sparse product),
    ``dense_rows`` rows used for the dense row-separable stages (scaled by n / dense_rows or
                   n_train / dense_rows).
    Returns per-stage estimated full-size seconds, the measured sample seconds, and their sums."""
    rng = np.random.default_rng(seed)
    f = x.shape[1]
    m = n_land + 1
    est: Dict[str, float] = {}
    spent = 0.0
    up_sparse = n / float(sample.n_rows)
    up_dense = n / float(dense_rows)

    # layer 1: Q[:, :F] = sigma_w A X            (src/compute.py:142)
    q_rows, t = _timed(lambda: np.float32(sigma_w) * sample.matmul(x))
    est["spmm_layer1"] = t * up_sparse
    spent += t

    # a [dense_rows x m] stand-in for rows of Q and the landmark rows (values only matter for timing
    # through the eigensolver's convergence, so they are drawn with the real factor's structure:
    # propagated features plus the bias column)
    reps = int(np.ceil(dense_rows / float(q_rows.shape[0])))
    q = np.zeros((dense_rows, m), dtype=np.float32)
    q[:, :f] = np.tile(q_rows, (reps, 1))[:dense_rows] * (1 + 0.05 * rng.standard_normal((dense_rows, 1)).astype(np.float32))
    q[:, -1] = sigma_b
    ql = np.zeros((n_land, m), dtype=np.float32)
    ql[:, :f] = land_rows_x
    ql[:, -1] = sigma_b

    for layer in range(layers - 1):
        # Gram + arc-cosine on the landmark columns   (src/compute.py:124-126)
        def gram():
            s = np.sqrt(np.sum(q ** 2, axis=1)).reshape(-1, 1)
            sl = np.sqrt(np.sum(ql ** 2, axis=1)).reshape(1, -1)
            theta = np.arccos(np.clip(q @ ql.T / s / sl, np.float32(-1), np.float32(1)))
            return orc._j1(theta, np.dtype(np.float32)) * s * sl
        e_rows, t = _timed(gram)
        est["gram_arccos_l%d" % (layer + 2)] = t * up_dense
        spent += t
        # landmark block + eigh (full size)            (src/compute.py:11-13)
        def root():
            sl = np.sqrt(np.sum(ql ** 2, axis=1)).reshape(-1, 1)
            theta = np.arccos(np.clip(ql @ ql.T / sl / sl.T, np.float32(-1), np.float32(1)))
            emm = orc._j1(theta, np.dtype(np.float32)) * sl * sl.T
            evals, evecs = np.linalg.eigh(emm)
            rt = np.sqrt(np.maximum(evals, 0))
            return evecs, rt, rt / np.maximum(evals, evals[-1] * np.float32(1e-4))
        (evecs, rt, inv), t = _timed(root)
        est["eigh_landmark_l%d" % (layer + 2)] = t
        spent += t
        # back-projection of the non-landmark rows      (src/compute.py:16)
        back, t = _timed(lambda: (e_rows @ evecs) * inv[None, :])
        est["backproject_l%d" % (layer + 2)] = t * up_dense
        spent += t
        # propagation of the N x N_a factor             (src/compute.py:145)
        factor = rng.random((n, n_land), dtype=np.float32)
        _, t = _timed(lambda: np.float32(sigma_w) * sample.matmul(factor))
        est["spmm_l%d" % (layer + 2)] = t * up_sparse
        spent += t
        del factor

    # fit: train Gram (rows of the training nodes), eigh (full), nugget sweep   (src/predict.py:57-62)
    up_train = n_train / float(dense_rows)
    gram_b, t = _timed(lambda: q.T @ q)
    est["fit_gram"] = t * up_train
    spent += t
    (w, v), t = _timed(lambda: np.linalg.eigh(gram_b + np.float32(1e-3) * np.eye(m, dtype=np.float32)))
    est["fit_eigh"] = t
    spent += t
    yb = np.zeros((dense_rows, n_classes), dtype=np.float32)
    yb[np.arange(dense_rows), rng.integers(0, n_classes, dense_rows)] = 1
    temp, t = _timed(lambda: v.T @ (q.T @ yb))
    est["fit_project"] = t * up_train
    spent += t
    d = np.float32(np.mean(np.sum(q ** 2, axis=1)))

    def sweep():
        fitted = np.zeros((n_nugget, dense_rows, n_classes), dtype=np.float32)
        for i, eps in enumerate(np.logspace(-3, 1, n_nugget).astype(np.float32)):
            fitted[i] = q @ (v @ (temp / (np.abs(w) + eps * d).reshape(-1, 1)))
        return fitted
    fitted, t = _timed(sweep)
    # the m x m x C part of every nugget does not shrink with the row sample: measure it apart
    _, t_small = _timed(lambda: [v @ (temp / (np.abs(w) + np.float32(0.1) * d).reshape(-1, 1)) for _ in range(n_nugget)])
    est["nugget_sweep"] = (t - t_small) * up_dense + t_small
    spent += t + t_small
    labels = rng.integers(0, n_classes, dense_rows)
    masks = {k: rng.random(dense_rows) < p for k, p in (("train", 0.66), ("val", 0.1), ("test", 0.24), ("landmark", 0.013))}
    _, t = _timed(lambda: orc.metrics(fitted, labels, masks))
    est["metrics"] = t * up_dense
    spent += t

    est["total_estimated_s"] = float(sum(v_ for k_, v_ in est.items()))
    est["sample_measured_s"] = float(spent)
    return est
