"""The headline workloads at FULL size through the public API on the B200, against golden fixtures written by
ONE run each of the unmodified reference on CPU (tests/golden/make_golden_big.py):

  * BASELINE.json configs[4]: Reddit shape, GCNGP-X, L = 2, fraction 50  (figure4.sh:2, table4.sh:25)
  * BASELINE.json configs[3]: ogbn-arxiv shape, GCNGP-X, L = 2, fraction 10

The graph is regenerated on the host from the same seed (checked through the fixture's input checksum), moved to
the GPU, and run exactly like ``src/main.py:108-116`` does.  Compared: the sampled Q Qᵀ block (rel 1e-4), the
posterior mean on 4 096 sampled rows at the val-selected nugget and at nugget 0.1, ALL N argmax predictions at
the selected nugget (flip count reported), and the three accuracy curves over the 101 nuggets."""
import json
import os

import numpy as np
import pytest
import torch

import cases
import parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(cases.BIG))
def test_full_size_run_matches_reference_golden(name):
    from gnngp_b200.gnngp import GNNGP
    golden = parity.load_big(name)
    if golden is None:
        pytest.skip("fixture tests/golden/big/%s.npz not generated" % name)
    case = cases.BIG[name]
    data, masks, nugget = cases.build_inputs(case)
    np.testing.assert_allclose(cases.input_checksum(data, masks), golden["checksum"], rtol=1e-7,
                               err_msg="the synthetic generator does not reproduce the fixture's inputs on this machine")
    dev = torch.device("cuda:0")
    model = GNNGP(data, device=dev, Nystrom=True)
    model.mask["landmark"] = masks["landmark"].to(dev)
    model.set_hyper_param(case["L"], case["sigma_b"], case["sigma_w"], initial=case["initial"], method=case["method"],
                          **case["params"])
    fit = model.predict(nugget.to(dev))
    q = model.get_kernel()
    idx = torch.from_numpy(golden["sample_index"]).to(dev)
    rows = torch.from_numpy(golden["row_sample"]).to(dev)
    best = int(golden["best"])
    qs = q[idx].double()
    sizes = {k: int(masks[k].sum()) for k in ("train", "val", "test")}
    rep = parity.check_big(golden, (qs @ qs.T).cpu().numpy(), fit[best][rows].cpu().numpy(), fit[50][rows].cpu().numpy(),
                           torch.argmax(fit[best], 1).cpu().numpy(), {k: v.numpy() for k, v in model.result.items()}, sizes)
    rep["best_nugget_index_ours"] = int(torch.argmax(model.result["val"]))
    rep["best_nugget_index_reference"] = best
    from gnngp_b200.ops import default_ops
    rep["root_solver"] = getattr(default_ops(), "last_root_solver", None)       # "krylov" on the arxiv shape (9 094 landmarks)
    rep["root_stats"] = dict(getattr(default_ops(), "last_root_stats", None) or {}) or None   # Ritz counts / residuals of the attempt
    rep["fit_solver"] = getattr(default_ops(), "last_fit_solver", None)
    print("\n%s: %s" % (name, json.dumps(rep)))
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "parity_big_%s.json" % name), "w") as fh:
        json.dump(rep, fh, indent=1)
    del model, fit, q
    torch.cuda.empty_cache()
