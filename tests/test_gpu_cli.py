"""The reference's command lines (table4.sh / figure4.sh GP rows) through gnngp_b200.main on the B200."""
import re

import pytest

pytestmark = pytest.mark.gpu


def run_cli(argv, capsys):
    from gnngp_b200 import main as cli
    table = cli.main(argv)
    return table, capsys.readouterr().out


def test_cora_exact_gcn_prints_the_reference_format(capsys):
    table, out = run_cli(["cora", "gcn", "gp"], capsys)            # table4.sh:12
    assert "Dataset: Cora" in out and "Method: GCN-GP" in out and "Number of runs: 1 for deterministic results." in out
    assert re.search(r"Run: 01, time: \d+\.\d{4}, train: \d\.\d{4}, val: \d\.\d{4}, test: \d\.\d{4}", out)
    assert table.shape == (1, 4) and 0.5 < float(table[0, 3]) <= 1.0


def test_pubmed_nystrom_runs_with_random_landmarks(capsys):
    table, out = run_cli(["pub", "gcn", "gp", "--fraction=4", "--runs", "3"], capsys)
    assert "Method: GCN-GP-X" in out and "Using Nystrom: 1/4 of training nodes chosen as landmark nodes." in out
    assert "Mean:" in out and "SD:" in out
    assert table.shape == (3, 4) and float(table[:, 3].min()) > 0.5


@pytest.mark.parametrize("method", ["gin", "sage", "gcn2", "ggp"])
def test_chameleon_methods(capsys, method):
    table, out = run_cli(["cham", method, "gp", "--sigma_b=0.316", "--num_layers=3"], capsys)   # table4.sh:15 style
    assert "Dataset: chameleon" in out
    assert table.shape == (1, 4) and float(table[0, 2]) > 0.0


def test_gnn_action_is_rejected(capsys):
    from gnngp_b200 import main as cli
    with pytest.raises(Exception, match="outside this package's scope"):
        cli.main(["cora", "gcn", "gnn"])


@pytest.mark.parametrize("argv, shape, fraction", [(["cora", "gcn", "gp"], "cora", 0),
                                                   (["pub", "gcn", "gp", "--fraction=4", "--runs", "1"], "pubmed", 4)])
def test_cli_metrics_match_the_oracle_on_the_same_graph(capsys, argv, shape, fraction):
    """the train / val / test figures the CLI prints at the selected nugget equal what the CPU
    oracle (restatement of compute.py / predict.py, pinned to the reference's goldens; run in float64 here) gets on the SAME synthetic
    graph and the SAME landmark draw (the CLI's own: ``torch.manual_seed(123)`` then ``rand(N) < 1/fraction``,
    src/main.py:95,111)."""
    import numpy as np
    import torch
    from gnngp_b200 import synth
    from gnngp_b200.datasets import transition_matrix
    from oracle import gnngp_oracle as orc
    table, _ = run_cli(argv, capsys)
    dev = torch.device("cuda:0")
    data = transition_matrix(normalization="sym")(synth.make_graph(shape, seed=0, device=dev, normalization=None))
    masks = {"train": data.train_mask.cpu().numpy(), "val": data.val_mask.cpu().numpy(), "test": data.test_mask.cpu().numpy()}
    if fraction:
        torch.manual_seed(123)
        masks["landmark"] = (data.train_mask & (torch.rand(data.num_nodes, device=dev) < 1 / fraction)).cpu().numpy()
    ref = orc.run(data.x.cpu().numpy(), data.y.cpu().numpy(), data.edge_index[0].cpu().numpy(), data.edge_index[1].cpu().numpy(),
                  data.edge_attr.cpu().numpy(), masks, torch.logspace(-3, 1, 101).numpy(), layers=2, sigma_b=0.0, sigma_w=1.0,
                  nystrom=bool(fraction), method="GCN", dtype=np.float64)   # fp64: free of the fp32 noise at small nuggets
    # The selected nugget is an argmax over a validation curve with exact ties (accuracies are multiples of 1/size) and
    # the plateau reaches down to the smallest nuggets, where ANY fp32 factor Q (the CLI's, the reference's) gives a
    # noise-dominated fit against this float64 oracle — only the training accuracy moves along the plateau (0.987 ... 1.0).
    # Checked: the best validation accuracy agrees to 2.5 samples, the test accuracy printed next to it agrees with the
    # oracle's test accuracy somewhere on ITS plateau, and the training accuracy is not below the plateau's.
    tol = {key: 2.5 / int(masks[key].sum()) + 1e-6 for key in ("train", "val", "test")}
    curves = {key: np.asarray(ref["result"][key], dtype=np.float64) for key in ("train", "val", "test")}
    assert abs(float(table[0, 2]) - float(ref["summary"]["val"])) <= tol["val"], (float(table[0, 2]), ref["summary"]["val"])
    plateau = curves["val"] >= curves["val"].max() - tol["val"]
    assert np.abs(curves["test"][plateau] - float(table[0, 3])).min() <= 2 * tol["test"] + 2e-3, (float(table[0, 3]), curves["test"][plateau])
    assert float(table[0, 1]) >= curves["train"][plateau].min() - tol["train"]
