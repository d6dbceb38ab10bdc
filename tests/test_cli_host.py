"""Host-side checks of the CLI driver and the drop-in shims (no GPU)."""
import sys
import os

import pytest


def test_cli_flags_and_name_resolution_match_reference():
    from gnngp_b200 import main as cli
    args = cli.build_parser().parse_args(["redd", "gcn", "gp", "--fraction=50"])
    assert (args.fraction, args.num_layers, args.sigma_b, args.sigma_w, args.device, args.runs) == (50, 2, 0.0, 1.0, 0, 0)
    assert cli._pick("redd", cli.DATASETS, "data", True) == "Reddit"
    assert cli._pick("c", cli.DATASETS, "data", True) == "Cora"           # first prefix match wins (src/main.py:41-44)
    assert cli._pick("sage", cli.METHODS, "method", False) == "SAGE"
    with pytest.raises(Exception, match="Unsupported data"):
        cli._pick("nope", cli.DATASETS, "data", True)
    with pytest.raises(Exception, match="Unsupported method"):
        cli._pick("gc", cli.METHODS, "method", False)
    assert cli._line([1.5, 0.25, 0.5, 0.75]) == "time: 1.5000, train: 0.2500, val: 0.5000, test: 0.7500"


def test_dropin_modules_expose_reference_names():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "gnn-gp_b200", "dropin"))
    saved = {m: sys.modules.pop(m, None) for m in ("gnngp", "compute", "predict", "datasets")}
    try:
        import compute, datasets, gnngp, predict
        for name in ("_init_kernel", "_init_kernel_Nystrom", "_get_kernel", "_get_kernel_Nystrom", "_sqrt_Nystrom",
                     "_ExxT_ReLU", "_ExxT_ReLU_Nystrom"):
            assert hasattr(compute, name)
        for name in ("fit", "fit_Nystrom", "result"):
            assert hasattr(predict, name)
        assert hasattr(gnngp, "GNNGP") and hasattr(datasets, "get_data") and hasattr(datasets, "transition_matrix")
    finally:
        sys.path.pop(0)
        for mod, old in saved.items():
            sys.modules.pop(mod, None)
            if old is not None:
                sys.modules[mod] = old


def test_hyper_param_contract():
    """set_hyper_param keeps unspecified values, replaces params, drops the cache (src/gnngp.py:71-85)."""
    from gnngp_b200.gnngp import GNNGP
    from gnngp_b200 import synth
    model = GNNGP(synth.make_graph("tiny"), device=None, Nystrom=True, alpha=0.3)
    assert (model.L, model.sigma_b, model.sigma_w, model.Nystrom, model.initial, model.method) == (2, 0.1, 1.0, True, "linear", "GCN")
    assert model.params == {"alpha": 0.3} and not model.computed
    model.computed = True
    model.set_hyper_param(sigma_b=0.5, method="SAGE")
    assert (model.L, model.sigma_b, model.method, model.params, model.computed) == (2, 0.5, "SAGE", {}, False)
    summary = model.get_summary()
    assert set(summary) == {"L", "sigma_b", "sigma_w", "Nystrom", "initial", "method"}


def test_streamed_upload_chunk_bounds_tile_the_edge_list():
    """Host logic of the streamed upload: row-aligned pieces of a row-sorted list; monotone tiling of [0, nnz)
    even for an unsorted list (so that every edge is always copied exactly once)."""
    import torch
    from gnngp_b200.datasets import chunk_bounds
    g = torch.Generator().manual_seed(0)
    n, nnz = 1000, 20000
    rows = torch.sort(torch.randint(0, n, (nnz,), generator=g)).values
    rb, nb = chunk_bounds(rows, n, 8)
    assert rb[0] == 0 and rb[-1] == n and nb[0] == 0 and nb[-1] == nnz and len(rb) == len(nb) == 9
    for c in range(8):
        piece = rows[nb[c]:nb[c + 1]]
        assert nb[c] <= nb[c + 1]
        if piece.numel():
            assert int(piece.min()) >= rb[c] and int(piece.max()) < rb[c + 1]
    shuffled = rows[torch.randperm(nnz, generator=g)]
    rb, nb = chunk_bounds(shuffled, n, 8)
    assert nb[0] == 0 and nb[-1] == nnz and all(a <= b for a, b in zip(nb, nb[1:]))
    # degenerate shapes: fewer rows than pieces, empty list
    rb, nb = chunk_bounds(torch.tensor([0, 0, 1, 2]), 3, 8)
    assert nb[0] == 0 and nb[-1] == 4 and all(a <= b for a, b in zip(nb, nb[1:]))
    rb, nb = chunk_bounds(torch.zeros(0, dtype=torch.int64), 5, 4)
    assert nb == [0, 0, 0, 0, 0]
