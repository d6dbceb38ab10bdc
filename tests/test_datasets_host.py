"""Data boundary on the host: ``transition_matrix`` / ``Scale`` / ``get_data`` of
``gnngp_b200.datasets`` against an independent DENSE numpy construction of what reference
``src/datasets.py:27-68`` computes (self-loops with the given weight, duplicate edges summed by the coalesce,
degree = column sums for "sym" / "col" and row sums for "row", ``inf -> 0`` for isolated nodes)."""
import numpy as np
import pytest
import torch

from gnngp_b200.datasets import GraphData, Scale, get_data, transition_matrix


def dense_reference(n, rows, cols, weights, loop_weight, normalization):
    a = np.zeros((n, n), dtype=np.float64)
    for r, c, w in zip(rows, cols, weights):                 # coalesce: duplicates add up (src/datasets.py:41)
        a[r, c] += w
    if loop_weight:
        a[np.arange(n), np.arange(n)] += loop_weight        # add_self_loops(fill_value) (src/datasets.py:36-39)
    with np.errstate(divide="ignore"):
        if normalization == "sym":
            s = a.sum(axis=0) ** -0.5                        # scatter_add over col (src/datasets.py:45-48)
            s[np.isinf(s)] = 0
            a = s[:, None] * a * s[None, :]
        elif normalization == "col":
            s = 1.0 / a.sum(axis=0)
            s[np.isinf(s)] = 0
            a = a * s[None, :]
        elif normalization == "row":
            s = 1.0 / a.sum(axis=1)
            s[np.isinf(s)] = 0
            a = a * s[:, None]
    return a


def random_edges(n, e, seed, weighted, isolated=()):
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, n, e)
    cols = rng.integers(0, n, e)
    keep = ~np.isin(rows, isolated) & ~np.isin(cols, isolated)
    rows, cols = rows[keep], cols[keep]
    dup = rng.integers(0, len(rows), len(rows) // 5)          # 20 % duplicated edges, some of them self-loops already
    rows, cols = np.concatenate([rows, rows[dup], [1, 1]]), np.concatenate([cols, cols[dup], [1, 1]])
    weights = rng.random(len(rows)) + 0.25 if weighted else None
    return rows, cols, weights


@pytest.mark.parametrize("normalization", ["sym", "col", "row", "none"])
@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("loop_weight", [1, 0, 2.5])
def test_transition_matrix_matches_dense_construction(normalization, weighted, loop_weight):
    n = 37
    isolated = (5, 20)                                        # no incident edge: degree = loop weight, or 0 -> inf -> 0
    rows, cols, weights = random_edges(n, 300, 7, weighted, isolated)
    data = GraphData(torch.zeros(n, 3), torch.zeros(n, dtype=torch.int64), torch.tensor(np.stack([rows, cols])),
                     None if weights is None else torch.tensor(weights, dtype=torch.float32))
    out = transition_matrix(self_loop_weight=loop_weight, normalization=normalization)(data)
    want = dense_reference(n, rows, cols, np.ones(len(rows)) if weights is None else weights.astype(np.float32).astype(np.float64),
                           loop_weight, normalization)
    idx = out.edge_index.numpy()
    # coalesced output: sorted by (row, col), no duplicates, exactly the non-zero pattern of the dense matrix
    key = idx[0] * n + idx[1]
    assert np.all(np.diff(key) > 0)
    pattern = np.zeros((n, n), dtype=bool)
    pattern[idx[0], idx[1]] = True
    dense_pattern = np.zeros((n, n), dtype=bool)
    dense_pattern[rows, cols] = True
    if loop_weight:
        dense_pattern[np.arange(n), np.arange(n)] = True
    assert np.array_equal(pattern, dense_pattern)
    got = np.zeros((n, n))
    got[idx[0], idx[1]] = out.edge_attr.numpy().astype(np.float64)
    np.testing.assert_allclose(got, want, rtol=3e-6, atol=1e-7)
    assert np.isfinite(got).all()
    if not loop_weight and normalization in ("sym", "col", "row"):
        assert got[list(isolated)].sum() == 0 and got[:, list(isolated)].sum() == 0


def test_sym_normalised_undirected_graph_is_symmetric_and_rows_of_row_norm_sum_to_one():
    n = 64
    rng = np.random.default_rng(3)
    r, c = rng.integers(0, n, 400), rng.integers(0, n, 400)
    ei = torch.tensor(np.stack([np.concatenate([r, c]), np.concatenate([c, r])]))
    sym = transition_matrix(normalization="sym")(GraphData(torch.zeros(n, 1), None, ei.clone(), None))
    a = torch.sparse_coo_tensor(sym.edge_index, sym.edge_attr, (n, n)).to_dense()
    assert torch.allclose(a, a.T, atol=1e-7)
    row = transition_matrix(normalization="row")(GraphData(torch.zeros(n, 1), None, ei.clone(), None))
    b = torch.sparse_coo_tensor(row.edge_index, row.edge_attr, (n, n)).to_dense()
    assert torch.allclose(b.sum(1), torch.ones(n), atol=1e-6)


def test_scale_and_get_data_follow_the_reference_contract():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(50, 7, generator=g) + 3.0
    data = GraphData(x.clone(), torch.arange(50) % 3, torch.tensor([[0, 1, 2], [1, 2, 0]]), None,
                     torch.arange(50) < 20, (torch.arange(50) >= 20) & (torch.arange(50) < 35), torch.arange(50) >= 35)
    out = Scale(center=True, scale=True)(data)
    xc = x.double() - x.double().mean(0, keepdim=True)                     # src/datasets.py:77-80
    xs = xc / xc.pow(2).sum(0, keepdim=True).sqrt()
    assert torch.allclose(out.x.double(), xs, atol=1e-6)
    n, feats, y, adj, mask = get_data(out)                                  # src/datasets.py:313-330
    assert n == 50 and feats is out.x and y is out.y and set(mask) == {"train", "val", "test"}
    assert adj.is_sparse and adj.shape == (50, 50) and not adj.is_coalesced()
    assert torch.equal(adj._values(), torch.ones(3))                         # edge_attr None -> unit weights
