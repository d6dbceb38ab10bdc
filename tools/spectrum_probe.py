"""GPU-box probe: spectra of the two eigendecompositions of the benchmark step (landmark block E_mm, train Gram
Q_b^T Q_b) at fractions 50 and 10 — how many eigenvalues are non-positive / below the fp32 noise / below the clamp,
and how close the Gram's most negative eigenvalue comes to the smallest shift eps*d (poles of the posterior solve)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnngp_b200 import synth  # noqa: E402
from gnngp_b200.gnngp import GNNGP  # noqa: E402
from gnngp_b200.ops import default_ops  # noqa: E402

dev = torch.device("cuda:0")
ops = default_ops()
data = synth.make_graph("reddit", seed=0).to(dev)
out = {}
for fraction in (50, 10):
    seen = []
    real = ops.eigh

    def spy(m, _real=real):
        w, v = _real(m)
        seen.append(w.double().cpu())
        return w, v

    ops.eigh = spy
    model = GNNGP(data, device=dev, Nystrom=True)
    model.mask["landmark"] = synth.draw_landmarks(data.train_mask.cpu(), fraction, seed=0).to(dev)
    model.set_hyper_param(2, 0.0, 1.0, method="GCN")
    model.predict(synth.nugget_grid(dev))
    del ops.eigh
    q = model.Q
    d = float((q.double() ** 2).sum(1).mean())
    rec = {}
    for label, w in zip(("landmark_block", "train_gram"), seen):
        top = float(w[-1])
        rec[label] = {"n": int(w.numel()), "max": top, "min": float(w[0]), "nonpositive": int((w <= 0).sum()),
                      "below_1e-7_max": int((w.abs() < 1e-7 * top).sum()), "below_1e-6_max": int((w < 1e-6 * top).sum()),
                      "below_1e-5_max": int((w < 1e-5 * top).sum()), "below_clamp_1e-4_max": int((w < 1e-4 * top).sum())}
    rec["nugget_scale_d"] = d
    rec["smallest_shift_eps_d"] = 1e-3 * d
    rec["gram_min_over_smallest_shift"] = float(seen[1][0]) / (1e-3 * d)
    out["fraction_%d" % fraction] = rec
    print(fraction, json.dumps(rec), flush=True)
    del model, q
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "spectrum_probe.json"), "w") as fh:
    json.dump(out, fh, indent=1)
