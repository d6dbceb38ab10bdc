"""The numpy oracle against the golden fixtures made from the unmodified reference, and live against
the imported reference when /root/reference is present (build container only).  CPU only."""
import os
import sys
import types

import numpy as np
import pytest
import torch

import cases
import parity
from oracle import gnngp_oracle as orc

HAVE_REF = os.path.isdir("/root/reference/src")


def oracle_run(case, data, masks, nugget, dtype=np.float32):
    np_masks = {k: v.numpy() for k, v in masks.items()}
    y = data.y.numpy()
    return orc.run(data.x.numpy(), y, data.edge_index[0].numpy(), data.edge_index[1].numpy(),
                   data.edge_attr.numpy(), np_masks, nugget.numpy(), layers=case["L"], sigma_b=case["sigma_b"],
                   sigma_w=case["sigma_w"], nystrom=case["fraction"] is not None, initial=case["initial"],
                   method=case["method"], dtype=dtype, **case["params"])


def block_of(out, idx):
    if "Q" in out:
        q = out["Q"][idx]
        return q @ q.T
    return out["K"][np.ix_(idx, idx)]


def run_case(name):
    golden = parity.load_golden(name)
    if golden is None:
        pytest.skip("fixture %s.npz not generated" % name)
    case = cases.CASES[name]
    data, masks, nugget = cases.build_inputs(case)
    np.testing.assert_allclose(cases.input_checksum(data, masks), golden["checksum"], rtol=1e-9,
                               err_msg="synthetic generator drifted from the fixture's inputs")
    out = oracle_run(case, data, masks, nugget)
    idx = golden["sample_index"]
    parity.check_kernel_block(golden, block_of(out, idx))
    best = int(golden["best"])
    parity.check_fit(golden, out["fit"][best], out["fit"][50])
    parity.check_predictions(golden, out["fit"][best])
    sizes = {k: int(masks[k].sum()) for k in ("train", "val", "test")}
    parity.check_curves(golden, out["result"], sizes)


@pytest.mark.parametrize("name", cases.FAST)
def test_oracle_matches_golden_fast(name):
    run_case(name)


@pytest.mark.slow
@pytest.mark.parametrize("name", cases.MEDIUM)
def test_oracle_matches_golden_medium(name):
    run_case(name)


@pytest.mark.parametrize("name", ["tiny-GCN-nys2-L2", "tiny-SAGE-exact-L2"])
def test_oracle_fp64_mode(name):
    """fp64 oracle against the fp64 reference block: the 1e-10 tolerance of north_star."""
    golden = parity.load_golden(name)
    if golden is None:
        pytest.skip("fixture missing")
    case = cases.CASES[name]
    data, masks, nugget = cases.build_inputs(case)
    out = oracle_run(case, data, masks, nugget, dtype=np.float64)
    err = parity.rel_fro(block_of(out, golden["sample_index"]), golden["ksamp64"])
    assert err <= parity.REL_FP64, err


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference only exists in the build container")
def test_oracle_pieces_live_against_reference():
    """Function-by-function: nystrom_root, relu_expectation[_nystrom], fit_*, metrics on random inputs."""
    sys.path.insert(0, "/root/reference/src")
    import compute as ref_compute
    import predict as ref_predict
    rng = np.random.default_rng(3)
    n, na = 90, 30
    q = rng.standard_normal((n, 17)).astype(np.float32)
    land = np.zeros(n, dtype=bool)
    land[rng.choice(n, na, replace=False)] = True
    tq, tl = torch.from_numpy(q), torch.from_numpy(land)
    ref = ref_compute._ExxT_ReLU_Nystrom(tq, tl).numpy()
    got = orc.relu_expectation_nystrom(q, land)
    assert parity.rel_fro(got @ got[land].T, ref @ ref[land].T) < 1e-5
    k = (q @ q.T).astype(np.float32)
    assert parity.rel_fro(orc.relu_expectation(k), ref_compute._ExxT_ReLU(torch.from_numpy(k)).numpy()) < 1e-5
    kc = k[:, land]
    r1, r2 = orc.nystrom_root(kc, land), ref_compute._sqrt_Nystrom(torch.from_numpy(kc), tl).numpy()
    assert parity.rel_fro(r1 @ r1.T, r2 @ r2.T) < 1e-4
    y = rng.integers(0, 3, n)
    train = np.zeros(n, dtype=bool)
    train[:50] = True
    nug = np.logspace(-2, 1, 7).astype(np.float32)
    f1 = orc.fit_nystrom(q, y, train, nug)
    f2 = ref_predict.fit_Nystrom(tq, torch.from_numpy(y), torch.from_numpy(train), torch.from_numpy(nug)).numpy()
    assert parity.rel_fro(f1, f2) < 1e-4
    kk = k + 0.5 * np.eye(n, dtype=np.float32)
    f1 = orc.fit_exact(kk, y, train, nug)
    f2 = ref_predict.fit(torch.from_numpy(kk), torch.from_numpy(y), torch.from_numpy(train), torch.from_numpy(nug)).numpy()
    assert parity.rel_fro(f1, f2) < 1e-4
    masks = {"train": train, "rest": ~train}
    m1 = orc.metrics(f1, y, masks)
    m2 = ref_predict.result(torch.from_numpy(f2), torch.from_numpy(y), {k_: torch.from_numpy(v) for k_, v in masks.items()})
    for key in masks:
        np.testing.assert_allclose(m1[key], m2[key].numpy(), atol=1e-6)


def test_native_spmm_matches_scipy():
    rng = np.random.default_rng(0)
    n = 300
    rows = rng.integers(0, n, 4000)
    cols = rng.integers(0, n, 4000)
    vals = rng.standard_normal(4000).astype(np.float32)
    adj = orc.Adjacency(rows, cols, vals, n)
    dense = rng.standard_normal((n, 37)).astype(np.float32)
    ref = np.asarray(adj._scipy @ dense)
    got = adj.matmul(dense)
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-5)
