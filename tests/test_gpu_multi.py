"""The NCCL row-sharded path against the single-GPU path, run by pytest whenever the box has two or more GPUs
(``tools/multigpu_check.py`` under torchrun with 2 ranks): sampled Q Qᵀ, posterior means and the accuracy curves of
six parity cases plus the arxiv-shape graph must agree.  Skipped on a single-GPU box (NCCL cannot put two ranks on
one device); the world_size 2/3 gloo tests in tests/test_shard_gloo.py cover the host logic on CPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs for two NCCL ranks")
def test_sharded_nccl_run_matches_single_gpu_run():
    env = dict(os.environ, GNNGP_CHECK_TAG="_pytest")
    out = os.path.join(ROOT, "gpurun_out", "multigpu_check_n2_pytest.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if os.path.exists(out):
        os.remove(out)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "multigpu_check.py")]
    proc = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    with open(out) as fh:
        report = json.load(fh)
    assert len(report) >= 6
    for name, row in report.items():
        assert row["qqT_rel"] <= 1e-5, (name, row)
        assert row["fit_mid_rel"] <= 1e-3, (name, row)
        assert row["val_curve_maxdiff"] <= 2e-3 and row["test_curve_maxdiff"] <= 2e-3, (name, row)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs a second GPU")
def test_model_on_a_non_current_device_matches_device_zero():
    """every operator must run on the device of its tensors, whatever the current device is
    (the reference honours ``--device`` because torch ops guard themselves, src/main.py:68)."""
    import cases
    from gnngp_b200.gnngp import GNNGP
    case = cases.CASES["small-GCN-nys4-L2"]
    data, masks, nugget = cases.build_inputs(case)
    outs = []
    for dev in (torch.device("cuda:0"), torch.device("cuda:1")):
        torch.cuda.set_device(0)                                   # current device stays 0 for both models
        model = GNNGP(data, device=dev, Nystrom=True)
        model.mask["landmark"] = masks["landmark"].to(dev)
        model.set_hyper_param(case["L"], case["sigma_b"], case["sigma_w"], method="GCN")
        fit = model.predict(nugget.to(dev))
        assert fit.device == dev and model.Q.device == dev
        outs.append((model.Q.double().cpu(), fit.cpu(), {k: v.clone() for k, v in model.result.items()}))
    qa, qb = outs[0][0][:300], outs[1][0][:300]
    assert float((qa @ qa.T - qb @ qb.T).norm() / (qb @ qb.T).norm()) < 1e-5
    assert float((outs[0][1][50] - outs[1][1][50]).norm() / outs[1][1][50].norm()) < 1e-3
    for key in ("train", "val", "test"):
        assert float((outs[0][2][key] - outs[1][2][key]).abs().max()) < 5e-3
