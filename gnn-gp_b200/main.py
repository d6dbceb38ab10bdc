"""Command-line driver for the GP branch — same positional arguments, flags, name resolution, run
counts and print format as the reference's ``src/main.py`` (lines 6-131), so the GP lines of
``table4.sh`` / ``table5.sh`` / ``figure3.sh`` / ``figure4.sh`` run unchanged:

    python -m gnngp_b200.main redd gcn gp --fraction=50
    python -m gnngp_b200.main cora gcn gp

Differences, all forced by the environment: the dataset is a synthetic graph with the shape of the
named dataset (``gnngp_b200.synth``; no network for the real downloads, reference
``src/datasets.py:84-310``); the ``GNN`` action (PyG baselines, ``src/main_gnn.py`` / ``main_batch.py``)
is out of scope; time is wall-clock with a device synchronize (the reference prints ``process_time``,
CPU-seconds without a sync, ``src/main.py:101-102``); a CUDA device is required.
"""
from __future__ import annotations

import time
from argparse import ArgumentParser

import torch

from . import synth
from .datasets import Scale
from .gnngp import GNNGP

DATASETS = ["Cora", "CiteSeer", "PubMed", "chameleon", "crocodile", "squirrel", "arxiv", "Reddit"]
METHODS = ["GCN", "GCN2", "GIN", "SAGE", "GGP", "RBF"]
ACTIONS = ["GNN", "GP"]
SHAPE_OF = {"Cora": "cora", "PubMed": "pubmed-public", "chameleon": "chameleon", "arxiv": "arxiv", "Reddit": "reddit"}


def _pick(value: str, options, what: str, prefix: bool):
    for option in options:
        if (option.upper().startswith(value.upper()) if prefix else option.upper() == value.upper()):
            return option
    raise Exception("Unsupported %s! Possible values: %s" % (what, " ".join(options)))


def build_parser() -> ArgumentParser:
    parser = ArgumentParser(description="GNNGP arguments")
    parser.add_argument("data")
    parser.add_argument("method")
    parser.add_argument("action")
    parser.add_argument("--device", type=int, default=0)
    parser.add_argument("--runs", type=int, default=0)
    parser.add_argument("--verbose", action="store_true")
    parser.add_argument("--center", action="store_true")
    parser.add_argument("--scale", action="store_true")
    parser.add_argument("--num_layers", type=int, default=2)
    parser.add_argument("--sigma_b", type=float, default=0.0)
    parser.add_argument("--sigma_w", type=float, default=1.0)
    parser.add_argument("--fraction", type=int, default=0)
    # accepted for command-line compatibility; only used by the (out-of-scope) GNN action
    parser.add_argument("--dim_hidden", type=int, default=256)
    parser.add_argument("--dropout", type=float, default=0.5)
    parser.add_argument("--lr", type=float, default=0.01)
    parser.add_argument("--epochs", type=int, default=100)
    parser.add_argument("--batch_size", type=int, default=0)
    return parser


def _line(values) -> str:
    return "time: %.4f, train: %.4f, val: %.4f, test: %.4f" % tuple(float(v) for v in values[:4])


def main(argv=None):
    args = build_parser().parse_args(argv)
    nystrom = args.fraction > 0
    action = args.action.upper()

    name = _pick(args.data, DATASETS, "data", prefix=True)
    print("Dataset: %s" % name)
    method = _pick(args.method, METHODS, "method", prefix=False)
    label = method
    if action == "GP":
        if method not in ("GGP", "RBF"):
            label += "-GP"
        if nystrom:
            label += "-X"
    print("Method: %s" % label)
    if action not in ACTIONS:
        raise Exception("Unsupported action! Possible values: " + " ".join(ACTIONS))
    if action == "GNN":
        raise Exception("The GNN baselines (src/main_gnn.py, src/main_batch.py) are outside this package's scope")
    if name not in SHAPE_OF:
        raise Exception("No synthetic shape for dataset %s; available: %s" % (name, " ".join(SHAPE_OF)))
    if not (args.device >= 0 and torch.cuda.is_available()):
        raise Exception("gnngp_b200 needs a CUDA device (the reference's --device -1 CPU path is not provided)")

    device = torch.device("cuda:%s" % args.device)
    norm = "sym" if method not in ("SAGE", "GGP") else "row"                       # src/main.py:66
    data = synth.make_graph(SHAPE_OF[name], seed=0, device=device, normalization=None)
    if args.center or args.scale:
        data = Scale(args.center, args.scale)(data)
    from .datasets import transition_matrix
    data = transition_matrix(normalization=norm)(data)
    print("Dataset loaded to: %s" % device)

    runs = args.runs
    if runs <= 0:
        if args.fraction <= 1:
            print("Number of runs: 1 for deterministic results.")
            runs = 1
        else:
            print("Number of runs: 10 for random results.")
            runs = 10
    else:
        print("Number of runs: %d" % runs)
    if nystrom:
        print("Using Nystrom: 1/%d of training nodes chosen as landmark nodes." % args.fraction)
    print("----")
    if args.verbose:
        print("Preprocessing: center=%s, scale=%s" % (args.center, args.scale))
        print("GP arguments:  num_layers=%d, sigma_b=%.2f, sigma_w=%.2f, Nystrom=%s, fraction=%d" %
              (args.num_layers, args.sigma_b, args.sigma_w, nystrom, args.fraction))
        print("----")

    torch.manual_seed(123)

    def clock() -> float:
        torch.cuda.synchronize(device)
        return time.perf_counter()

    now = clock()
    table = torch.zeros((runs, 4))
    epsilon = torch.logspace(-3, 1, 101, device=device)
    params = {}
    if method == "GCN2":
        params = {"alpha": 0.1, "theta": 0.5}
    if method == "GGP":
        params = {"initial": "polynomial", "c": 5.0, "d": 3.0}
    model = GNNGP(data, device=device, Nystrom=nystrom)
    for j in range(runs):
        if nystrom:
            model.mask["landmark"] = model.mask["train"] & (torch.rand(model.N, device=device) < 1 / args.fraction)
        if method != "RBF":
            model.set_hyper_param(args.num_layers, args.sigma_b, args.sigma_w, method=method, **params)
            model.predict(epsilon)
            summary = model.get_summary()
            table[j] = torch.tensor([clock() - now, summary["train"], summary["val"], summary["test"]])
        else:
            for gamma in torch.logspace(-2, 2, 101):                               # src/main.py:118-124
                model.set_hyper_param(initial="rbf", method=method, gamma=float(gamma))
                model.predict(epsilon)
                summary = model.get_summary()
                if summary["val"] > table[j][2]:
                    table[j] = torch.tensor([0.0, summary["train"], summary["val"], summary["test"]])
            table[j][0] = clock() - now
        now = clock()
        print("Run: %02d," % (j + 1), _line(table[j]))
    if runs > 1:
        print("----")
        print("Mean:   ", _line(torch.mean(table, dim=0)))
        print("SD:     ", _line(torch.std(table, dim=0)))
    print("----")
    return table


if __name__ == "__main__":
    main()
