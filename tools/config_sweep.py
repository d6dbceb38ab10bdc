"""GPU-box tool: seconds per `set_hyper_param + predict + get_summary` for the five BASELINE.json configs
(and the Reddit-shape landmark sweep), next to the numpy oracle on the host for the shapes it finishes
in seconds.  Writes gpurun_out/config_sweep.json."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnngp_b200 import synth                 # noqa: E402
from gnngp_b200.gnngp import GNNGP           # noqa: E402
from gnngp_b200.ops import default_ops       # noqa: E402
from oracle import gnngp_oracle as orc       # noqa: E402

dev = torch.device("cuda:0")
ops = default_ops()
GCN2 = dict(alpha=0.1, theta=0.5)
out = []


def gpu_time(shape, method, fraction, layers, sigma_b, norm, params, reps=3, landmark_is_train=False, cpu=False):
    data = synth.make_graph(shape, seed=0, device=dev, normalization=norm)
    nugget = synth.nugget_grid(dev)
    nystrom = fraction is not None
    model = GNNGP(data, device=dev, Nystrom=nystrom)
    land = None
    if nystrom:
        land = data.train_mask.clone() if landmark_is_train else synth.draw_landmarks(data.train_mask, fraction, seed=0)
        model.mask["landmark"] = land

    def step():
        model.set_hyper_param(layers, sigma_b, 1.0, method=method, **params)
        model.predict(nugget)
        return model.get_summary()
    step()
    torch.cuda.synchronize()
    l0 = ops.launch_count()
    t0 = time.perf_counter()
    for _ in range(reps):
        s = step()
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / reps
    row = dict(shape=shape, method=method + ("-GP-X" if nystrom else "-GP"), fraction=fraction, L=layers, sigma_b=sigma_b,
               nodes=data.num_nodes, nnz=int(data.edge_index.shape[1]), landmarks=int(land.sum()) if land is not None else 0,
               gpu_s=round(sec, 5), launches=(ops.launch_count() - l0) // reps,
               val=float(s["val"]), test=float(s["test"]), nugget=float(s["nugget"]))
    if cpu:
        masks = {"train": data.train_mask.cpu().numpy(), "val": data.val_mask.cpu().numpy(), "test": data.test_mask.cpu().numpy()}
        if land is not None:
            masks["landmark"] = land.cpu().numpy()
        t0 = time.perf_counter()
        ref = orc.run(data.x.cpu().numpy(), data.y.cpu().numpy(), data.edge_index[0].cpu().numpy(), data.edge_index[1].cpu().numpy(),
                      data.edge_attr.cpu().numpy(), masks, nugget.cpu().numpy(), layers=layers, sigma_b=sigma_b, sigma_w=1.0,
                      nystrom=nystrom, method=method, **params)
        row["cpu_oracle_s"] = round(time.perf_counter() - t0, 3)
        row["cpu_val"] = ref["summary"]["val"]
        row["cpu_test"] = ref["summary"]["test"]
    out.append(row)
    print(json.dumps(row), flush=True)
    del model, data
    torch.cuda.empty_cache()


# configs[0]: Cora-shape exact GCNGP
gpu_time("cora", "GCN", None, 2, 0.0, "sym", {}, cpu=True)
# configs[1]: PubMed-shape GCNGP-X fraction 4
gpu_time("pubmed", "GCN", 4, 2, 0.0, "sym", {}, cpu=True)
# configs[2]: chameleon-shape GIN / SAGE / GCN2, depth sweep, exact and landmark = train
for method, norm, params in (("GIN", "sym", {}), ("SAGE", "row", {}), ("GCN2", "sym", GCN2)):
    for layers in (2, 4, 6, 8, 10):
        gpu_time("chameleon", method, None, layers, 0.316, norm, params, cpu=(layers in (2, 10)))
    gpu_time("chameleon", method, 1, 2, 0.316, norm, params, landmark_is_train=True, cpu=True)
# configs[3]: arxiv-shape GCNGP-X fraction 10
gpu_time("arxiv", "GCN", 10, 2, 0.0, "sym", {}, reps=2)
gpu_time("arxiv", "GCN", 50, 2, 0.0, "sym", {}, reps=2)
# configs[4]: Reddit-shape landmark sweep
for fraction in (50, 25, 10) + ((5,) if "--f5" in sys.argv else ()):
    gpu_time("reddit", "GCN", fraction, 2, 0.0, "sym", {}, reps=2)

with open(os.path.join(ROOT, "gpurun_out", "config_sweep.json"), "w") as fh:
    json.dump(out, fh, indent=1)
