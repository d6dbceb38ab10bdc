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
