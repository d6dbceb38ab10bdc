"""Seeded synthetic graphs with the shapes of the five ``BASELINE.json`` configs.

There is no network in the build or bench environment, so the Planetoid / MUSAE / OGB / Reddit
downloads of the reference (``src/datasets.py:84-310``) are replaced by generated graphs with the
same node / edge / feature / class / split counts.  Degrees follow a power law
(hub rows stress the SpMM row-split scheduler), edges and features are label-correlated (planted
partition) so that GP predictions have non-degenerate margins.

Everything is drawn from a ``torch.Generator`` on the requested device; the same seed on the same
device type reproduces the same graph.  Small shapes are generated on CPU (shared with the oracle
and the golden fixtures), the Reddit shape on the GPU.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .datasets import GraphData, transition_matrix

# name -> nodes, undirected edges, features, classes (0 = regression), split sizes
SHAPES: Dict[str, dict] = {
    # BASELINE.json configs[0]: Cora-shape, 10 556 directed edges
    "cora": dict(n=2708, und=5278, f=1433, c=7, split=(140, 500, 1000), feat="bow"),
    # configs[1]: PubMed-shape, 88 648 directed edges; dense-label split so fraction=4 gives ~1.5k landmarks
    "pubmed": dict(n=19717, und=44324, f=500, c=3, split=(5915, 500, 1000), feat="bow"),
    # PubMed-shape with the public 60/500/1000 split: N_a < F exercises compute.py:47
    "pubmed-public": dict(n=19717, und=44324, f=500, c=3, split=(60, 500, 1000), feat="bow"),
    # configs[2]: chameleon-shape (regression), 36 101 directed edges
    "chameleon": dict(n=2277, und=31400, f=3132, c=0, split=(1092, 729, 456), feat="bow"),
    # configs[3]: ogbn-arxiv-shape, 1.17 M directed edges symmetrised
    "arxiv": dict(n=169343, und=1166243, f=128, c=40, split=(90941, 29799, 48603), feat="gauss", sep=0.25, p_same=0.6,
                  exact_edges=True, label_noise=0.08),
    # configs[4]: Reddit-shape, 114.6 M directed edges
    "reddit": dict(n=232965, und=57307946, f=602, c=41, split=(153431, 23831, 55703), feat="gauss", sep=0.1, p_same=0.5,
                   exact_edges=True, label_noise=0.06),
    # tiny shapes for unit tests / smoke
    "tiny": dict(n=257, und=1100, f=40, c=4, split=(120, 60, 77), feat="gauss", sep=0.3, p_same=0.5),
    "tiny-reg": dict(n=203, und=900, f=24, c=0, split=(100, 50, 53), feat="gauss", sep=0.5, p_same=0.6),
    "small": dict(n=1500, und=12000, f=96, c=5, split=(700, 300, 500), feat="gauss", sep=0.15, p_same=0.5),
    # smoke(): the smallest graph on which the auto dispatch takes the production kernels — dense rows (average degree
    # ~84) and >= 256 operand columns for the panel SpMM, >= 64 output tiles for the tcgen05 contractions
    "smoke": dict(n=6000, und=250000, f=300, c=5, split=(2400, 1200, 2400), feat="gauss", sep=0.15, p_same=0.5),
}


def _powerlaw_endpoints(n: int, count: int, power: float, gen: torch.Generator, device) -> torch.Tensor:
    """Draw ``count`` node ids with P(rank r) ∝ (r+1)^-power, ranks mapped through a random permutation."""
    weights = torch.arange(1, n + 1, device=device, dtype=torch.float64).pow(-power)
    cdf = torch.cumsum(weights, 0)
    cdf /= cdf[-1].clone()
    u = torch.rand(count, generator=gen, device=device, dtype=torch.float64)
    ranks = torch.searchsorted(cdf, u).clamp_(max=n - 1)
    perm = torch.randperm(n, generator=gen, device=device)
    return perm[ranks]


def make_graph(shape: str = "tiny", seed: int = 0, device: Optional[torch.device] = None,
               normalization: Optional[str] = "sym", power: float = 0.7, p_same: float = 0.8,
               overrides: Optional[dict] = None) -> GraphData:
    """Generate one graph of the named shape.  ``normalization`` is applied with
    :class:`gnngp_b200.datasets.transition_matrix` ("sym" for GCN/GCN2/GIN, "row" for SAGE/GGP —
    reference ``src/main.py:66``); ``None`` leaves raw unit edge weights."""
    spec = dict(SHAPES[shape])
    if overrides:
        spec.update(overrides)
    device = torch.device("cpu") if device is None else torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(1000003 * seed + 17)
    n, und, f, c = spec["n"], spec["und"], spec["f"], spec["c"]
    groups = c if c > 0 else 5
    p_same = spec.get("p_same", p_same)
    sep = spec.get("sep", 0.9)

    label = torch.randint(0, groups, (n,), generator=gen, device=device)
    order = torch.argsort(label, stable=True)
    counts = torch.bincount(label, minlength=groups)
    starts = torch.cumsum(counts, 0) - counts

    # edges: power-law source, planted-partition destination
    draw = int(und * 1.04) + 8
    src = _powerlaw_endpoints(n, draw, power, gen, device)
    same = torch.rand(draw, generator=gen, device=device) < p_same
    g = label[src]
    within = (torch.rand(draw, generator=gen, device=device, dtype=torch.float64)
              * counts[g].to(torch.float64)).long()
    within = torch.minimum(within, counts[g] - 1)
    dst_same = order[starts[g] + within]
    dst_any = torch.randint(0, n, (draw,), generator=gen, device=device)
    dst = torch.where(same, dst_same, dst_any)
    keep = src != dst
    lo = torch.minimum(src[keep], dst[keep])
    hi = torch.maximum(src[keep], dst[keep])
    key = torch.unique(lo * n + hi)
    if spec.get("exact_edges"):
        # hub-heavy shapes lose many draws to duplicates: top up until the undirected edge count is reached
        rounds = 0
        while key.numel() < und and rounds < 12:
            extra = int((und - key.numel()) * 1.3) + 1024
            s2 = _powerlaw_endpoints(n, extra, power, gen, device)
            d2 = torch.where(torch.rand(extra, generator=gen, device=device) < p_same,
                             order[starts[label[s2]] + torch.minimum((torch.rand(extra, generator=gen, device=device, dtype=torch.float64)
                                                                      * counts[label[s2]].to(torch.float64)).long(), counts[label[s2]] - 1)],
                             torch.randint(0, n, (extra,), generator=gen, device=device))
            ok = s2 != d2
            key = torch.unique(torch.cat([key, torch.minimum(s2[ok], d2[ok]) * n + torch.maximum(s2[ok], d2[ok])]))
            rounds += 1
    if key.numel() > und:
        pick = torch.randperm(key.numel(), generator=gen, device=device)[:und]
        key = key[pick]
    lo = torch.div(key, n, rounding_mode="floor")
    hi = key - lo * n
    edge_index = torch.stack([torch.cat([lo, hi]), torch.cat([hi, lo])])
    del key, lo, hi, src, dst, dst_same, dst_any, within, same, keep

    # features
    if spec["feat"] == "gauss":
        means = torch.randn(groups, f, generator=gen, device=device)
        x = torch.randn(n, f, generator=gen, device=device) + sep * means[label]
    else:  # sparse binary bag-of-words with class prototype words
        base = torch.rand(n, f, generator=gen, device=device) < 0.012
        proto = torch.rand(groups, f, generator=gen, device=device) < 0.05
        hit = torch.rand(n, f, generator=gen, device=device) < 0.35
        x = (base | (proto[label] & hit)).to(torch.float32)
        empty = x.sum(1) == 0
        if bool(empty.any()):  # keep every row non-zero (sigma_b=0 would give 0/0 otherwise)
            x[empty, 0] = 1.0

    # targets
    if c > 0:
        y = label.clone()
        noise = spec.get("label_noise", 0.0)
        if noise > 0:     # observed labels differ from the planted community for a fraction of nodes
            flip = torch.rand(n, generator=gen, device=device) < noise
            y = torch.where(flip, torch.randint(0, c, (n,), generator=gen, device=device), y)
    else:
        level = torch.randn(groups, generator=gen, device=device)
        y = (level[label] + 0.5 * torch.randn(n, generator=gen, device=device)).to(torch.float32)

    # fixed-count random split
    nt, nv, ns = spec["split"]
    perm = torch.randperm(n, generator=gen, device=device)
    train = torch.zeros(n, dtype=torch.bool, device=device)
    val = torch.zeros(n, dtype=torch.bool, device=device)
    test = torch.zeros(n, dtype=torch.bool, device=device)
    train[perm[:nt]] = True
    val[perm[nt:nt + nv]] = True
    test[perm[nt + nv:nt + nv + ns]] = True

    data = GraphData(x, y, edge_index, None, train, val, test, num_classes=c)
    if normalization is not None:
        data = transition_matrix(normalization=normalization)(data)
    return data


def draw_landmarks(train_mask: torch.Tensor, fraction: int, seed: int = 0) -> torch.Tensor:
    """``train & (rand(N) < 1/fraction)`` (reference ``src/main.py:111``) from a seeded generator on
    the mask's device, so that both sides of a parity run can be handed the same landmark set."""
    gen = torch.Generator(device=train_mask.device)
    gen.manual_seed(7919 * seed + 101)
    u = torch.rand(train_mask.numel(), generator=gen, device=train_mask.device)
    return train_mask & (u < 1.0 / fraction)


def nugget_grid(device=None) -> torch.Tensor:
    """``logspace(-3, 1, 101)`` — the nugget sweep of reference ``src/main.py:104``."""
    return torch.logspace(-3, 1, 101, device=device)


def describe(data: GraphData) -> dict:
    return dict(nodes=data.num_nodes, nnz=int(data.edge_index.size(1)), feats=int(data.x.shape[1]),
                classes=int(data.num_classes or 0), train=int(data.train_mask.sum()),
                val=int(data.val_mask.sum()), test=int(data.test_mask.sum()))
