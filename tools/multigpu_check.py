"""GPU-box check under torchrun (N >= 2): the NCCL row-sharded path against the single-GPU path on the same
inputs (parity cases + an arxiv-shape graph).  Rank 0 prints the relative differences."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from gnngp_b200 import synth  # noqa: E402
from gnngp_b200.gnngp import GNNGP  # noqa: E402
from gnngp_b200.shard import ShardedGNNGP  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
report = {}


def compare(name, data, masks, nugget, hyper):
    sharded = ShardedGNNGP(data, device=dev, Nystrom=True)
    sharded.mask["landmark"] = masks["landmark"].to(dev)
    sharded.set_hyper_param(**hyper)
    sharded.predict(nugget.to(dev))
    fit = sharded.gather_fit()
    q = sharded.comm.all_gather_rows(sharded.Q, sharded.layout)
    res = sharded.result
    row = None
    if rank == 0:
        single = GNNGP(data, device=dev, Nystrom=True)
        single.mask["landmark"] = masks["landmark"].to(dev)
        single.set_hyper_param(**hyper)
        ref = single.predict(nugget.to(dev))
        idx = torch.randperm(data.num_nodes, device=dev)[:1500]
        a, b = q[idx].double(), single.Q[idx].double()
        ka, kb = a @ a.T, b @ b.T
        row = {"qqT_rel": float((ka - kb).norm() / kb.norm()),
               "fit_mid_rel": float((fit[50] - ref[50]).norm() / ref[50].norm()),
               "val_curve_maxdiff": float((res["val"] - single.result["val"]).abs().max()),
               "test_curve_maxdiff": float((res["test"] - single.result["test"]).abs().max()),
               "gathered_GB": sharded.comm.bytes_gathered / 1e9, "reduced_GB": sharded.comm.bytes_reduced / 1e9,
               "pulled_over_nvlink_GB": sharded.comm.bytes_pulled / 1e9, "peer_gather": bool(sharded.comm.peer_gather),
               "fit_solver": getattr(sharded.pipe, "last_fit_solver", None),
               "best_nugget_sharded_vs_single": [float(sharded.get_summary()["nugget"]), float(single.get_summary()["nugget"])],
               "test_at_best_sharded_vs_single": [float(sharded.get_summary()["test"]), float(single.get_summary()["test"])]}
        report[name] = row
        print(name, json.dumps(row), flush=True)
    dist.barrier()


for name in ("small-GCN-nys4-L2", "small-SAGE-nys4-L4", "pubmed-GCN-nys4-L2", "pubmedpublic-GCN-nys4-L2", "chameleon-GCN2-nys1-L2"):
    case = cases.CASES[name]
    data, masks, nugget = cases.build_inputs(case)
    compare(name, data, masks, nugget, dict(L=case["L"], sigma_b=case["sigma_b"], sigma_w=case["sigma_w"], initial=case["initial"],
                                            method=case["method"], **case["params"]))

data = synth.make_graph("arxiv", seed=0, device=dev)
land = synth.draw_landmarks(data.train_mask, 50, seed=0)
compare("arxiv-f50", data, {"landmark": land}, synth.nugget_grid(dev), dict(L=2, sigma_b=0.0, sigma_w=1.0, method="GCN"))
if "--reddit" in sys.argv:
    del data
    torch.cuda.empty_cache()
    data = synth.make_graph("reddit", seed=0, device=dev)
    land = synth.draw_landmarks(data.train_mask, 50, seed=0)
    compare("reddit-f50", data, {"landmark": land}, synth.nugget_grid(dev), dict(L=2, sigma_b=0.0, sigma_w=1.0, method="GCN"))
if rank == 0:
    tag = os.environ.get("GNNGP_CHECK_TAG", "")
    with open(os.path.join(ROOT, "gpurun_out", "multigpu_check_n%d%s.json" % (world, tag)), "w") as fh:
        json.dump(report, fh, indent=1)
dist.destroy_process_group()
