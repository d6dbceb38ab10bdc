"""Parity cases shared by the golden generator, the oracle tests and the GPU parity tests.

Each case names a synthetic graph (``gnngp_b200.synth``), the hyper-parameters of
``GNNGP.set_hyper_param`` (reference ``src/gnngp.py:71-85``) and how the landmark mask is drawn.
``fraction=None`` → exact path; ``fraction=1`` → landmark = train (``src/main.py:111`` with
``rand < 1.0``; ``notebooks/chameleon.ipynb`` cell 2)."""
from __future__ import annotations

import numpy as np
import torch

from gnngp_b200 import synth

GCN2 = dict(alpha=0.1, theta=0.5)            # src/main.py:106
GGP = dict(c=5.0, d=3.0)                     # src/main.py:107 (initial="polynomial")


def _case(shape, method, fraction, L=2, sigma_b=0.1, sigma_w=1.0, initial="linear", seed=0, norm=None, **params):
    if norm is None:
        norm = "row" if method in ("SAGE", "GGP") else "sym"       # src/main.py:66
    return dict(shape=shape, method=method, fraction=fraction, L=L, sigma_b=sigma_b, sigma_w=sigma_w,
                initial=initial, seed=seed, norm=norm, params=params)


CASES = {}
for _m, _p in (("GCN", {}), ("GCN2", GCN2), ("GIN", {}), ("SAGE", {})):
    for _L in (2, 3):
        CASES["tiny-%s-exact-L%d" % (_m, _L)] = _case("tiny", _m, None, L=_L, sigma_b=0.2, sigma_w=1.1, **_p)
        CASES["tiny-%s-nys2-L%d" % (_m, _L)] = _case("tiny", _m, 2, L=_L, sigma_b=0.2, sigma_w=1.1, **_p)
    CASES["tinyreg-%s-exact-L2" % _m] = _case("tiny-reg", _m, None, L=2, sigma_b=0.316, **_p)
    CASES["tinyreg-%s-nys1-L3" % _m] = _case("tiny-reg", _m, 1, L=3, sigma_b=0.316, **_p)
CASES["tiny-GGP-exact"] = _case("tiny", "GGP", None, initial="polynomial", **GGP)
CASES["tiny-GGP-nys2"] = _case("tiny", "GGP", 2, initial="polynomial", **GGP)
CASES["tiny-RBF-exact"] = _case("tiny", "RBF", None, initial="rbf", gamma=0.01)
CASES["tiny-RBF-nys2"] = _case("tiny", "RBF", 2, initial="rbf", gamma=0.01)
for _init, _p in (("laplacian", dict(gamma=0.02)), ("arccos", {}), ("sigmoid", dict(gamma=0.01, c=0.1)),
                  ("polynomial", dict(c=1.0, d=2.0))):
    CASES["tiny-GCN-exact-%s" % _init] = _case("tiny", "GCN", None, initial=_init, **_p)
    CASES["tiny-GCN-nys2-%s" % _init] = _case("tiny", "GCN", 2, initial=_init, **_p)
CASES["tiny-GCN-nys8-L2"] = _case("tiny", "GCN", 8, L=2, sigma_b=0.0)         # N_a < F: compute.py:47
CASES["small-GCN-nys4-L2"] = _case("small", "GCN", 4, L=2, sigma_b=0.0)
CASES["small-SAGE-nys4-L4"] = _case("small", "SAGE", 4, L=4, sigma_b=0.3)
# BASELINE.json configs[0]: Cora-shape exact GCNGP, CLI defaults sigma_b=0, sigma_w=1 (src/main.py:21-24)
CASES["cora-GCN-exact-L2"] = _case("cora", "GCN", None, L=2, sigma_b=0.0)
# configs[1]: PubMed-shape GCNGP-X fraction=4
CASES["pubmed-GCN-nys4-L2"] = _case("pubmed", "GCN", 4, L=2, sigma_b=0.0)
CASES["pubmedpublic-GCN-nys4-L2"] = _case("pubmed-public", "GCN", 4, L=2, sigma_b=0.0)
# configs[2]: chameleon-shape, GIN/SAGE/GCN2, depth sweep (table4.sh:15 uses sigma_b=0.316)
for _m, _p in (("GIN", {}), ("SAGE", {}), ("GCN2", GCN2)):
    for _L in (2, 6, 10):
        CASES["chameleon-%s-exact-L%d" % (_m, _L)] = _case("chameleon", _m, None, L=_L, sigma_b=0.316, **_p)
    CASES["chameleon-%s-nys1-L2" % _m] = _case("chameleon", _m, 1, L=2, sigma_b=0.316, **_p)

# cases small enough for the default CPU suite (the rest are exercised on the GPU box / by make_golden)
FAST = [k for k in CASES if k.startswith(("tiny", "small"))]
MEDIUM = [k for k in CASES if not k.startswith(("tiny", "small"))]

SAMPLE = 64   # size of the sampled K / QQ^T block stored in the fixtures


def build_inputs(case: dict):
    """Graph, masks (with landmark when Nystrom), nugget grid — all CPU torch tensors."""
    data = synth.make_graph(case["shape"], seed=case["seed"], normalization=case["norm"])
    masks = {"train": data.train_mask, "val": data.val_mask, "test": data.test_mask}
    if case["fraction"] is not None:
        if case["fraction"] == 1:
            masks["landmark"] = data.train_mask.clone()
        else:
            masks["landmark"] = synth.draw_landmarks(data.train_mask, case["fraction"], seed=case["seed"])
    nugget = synth.nugget_grid()
    return data, masks, nugget


def sample_index(n: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(12345 + seed)
    return np.sort(rng.choice(n, size=min(SAMPLE, n), replace=False))


def input_checksum(data, masks) -> np.ndarray:
    vals = [float(data.x.double().sum()), float(data.edge_attr.double().sum()), float(data.edge_index.size(1)),
            float(data.y.double().sum()), float(masks["train"].sum()),
            float(masks["landmark"].sum()) if "landmark" in masks else -1.0]
    return np.asarray(vals, dtype=np.float64)


# ---------------------------------------------------------------------------------------------------------
# Full-size cases (BASELINE.json configs[3] and configs[4]).  Their goldens come from ONE run each of the
# unmodified reference on the build container's CPU (tests/golden/make_golden_big.py: minutes, tens of GB),
# so they are kept out of CASES (the parametrised small-case suites) and hold samples instead of full arrays.
# ---------------------------------------------------------------------------------------------------------
BIG = {
    # configs[4], the headline workload: figure4.sh:2 / table4.sh:25 (fraction 50 is the largest landmark set the
    # reference's CPU path fits in this container's 62 GB; fraction 10 needs ~85 GB)
    "reddit-GCN-nys50-L2": _case("reddit", "GCN", 50, L=2, sigma_b=0.0),
    # configs[3] as stated: ogbn-arxiv shape, fraction 10
    "arxiv-GCN-nys10-L2": _case("arxiv", "GCN", 10, L=2, sigma_b=0.0),
}
BIG_ROWS = 4096   # rows of the posterior mean kept in a big fixture (all N argmax predictions are kept)


def big_row_sample(n: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(54321 + seed)
    return np.sort(rng.choice(n, size=min(BIG_ROWS, n), replace=False))
