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
