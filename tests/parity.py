"""Parity metrics of SURVEY.md §8c: compare K / Q Qᵀ blocks (never Q itself — eigh leaves a rotation
free), posterior means at fixed nuggets, metric curves and argmax predictions against a golden
fixture written by ``tests/golden/make_golden.py`` from the unmodified reference."""
from __future__ import annotations

import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerances
REL_FP32 = 1e-4
REL_FP64 = 1e-10


def load_golden(name: str):
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    if not os.path.exists(path):
        return None
    return dict(np.load(path))


def rel_fro(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def check_kernel_block(golden, block, tol=REL_FP32):
    err = rel_fro(block, golden["ksamp"])
    err64 = rel_fro(block, golden["ksamp64"])
    assert err <= tol, "K block differs from the fp32 reference: rel %.3e > %.1e" % (err, tol)
    assert err64 <= tol, "K block differs from the fp64 reference: rel %.3e > %.1e" % (err64, tol)
    return err, err64


def check_fit(golden, fit_best, fit_mid, tol=REL_FP32):
    """Posterior means at the val-selected nugget and at nugget 0.1.  The reference's own fp32 run
    deviates from its fp64 run by more than 1e-4 at small nuggets (SURVEY.md §7 hard part 3): there the error
    is fp32 rounding noise times the conditioning of (K_bb + eps d I), and two fp32 implementations are two
    draws of that noise (measured spread between our own engines on chameleon-SAGE-exact-L10 at nugget
    1e-3: 2.6e-3 ... 2.6e-2 against the reference's 6.5e-3).  The bound is max(tol, 6 x the reference's
    self-error) against the fp64 run; well-conditioned nuggets keep the plain 1e-4."""
    out = {}
    for key, value in (("fit_best", fit_best), ("fit_mid", fit_mid)):
        floor = rel_fro(golden[key], golden[key + "64"])
        bound = max(tol, 6.0 * floor)
        e32 = rel_fro(value, golden[key])
        e64 = rel_fro(value, golden[key + "64"])
        assert e64 <= bound, "%s vs fp64 reference: rel %.3e > %.3e (reference self-error %.2e)" % (key, e64, bound, floor)
        assert e32 <= 2 * bound, "%s vs fp32 reference: rel %.3e > %.3e" % (key, e32, 2 * bound)
        out[key] = (e32, e64, floor)
    return out


def check_predictions(golden, fit_best, exact=True):
    if "pred_best" not in golden:
        return None
    pred = np.argmax(np.asarray(fit_best), axis=1)
    agree = float(np.mean(pred == golden["pred_best"]))
    if exact and agree < 1.0:
        # a flip is only tolerated where the fp64 reference's top-2 margin is inside the reference's own
        # fp32-vs-fp64 error at this nugget (its fp32 run flips such nodes against its fp64 run too)
        top2 = np.sort(golden["fit_best64"], axis=1)[:, -2:]
        margin = top2[:, 1] - top2[:, 0]
        noise = float(np.max(np.abs(golden["fit_best"].astype(np.float64) - golden["fit_best64"])))
        scale = float(np.max(np.abs(golden["fit_best64"])))
        bad = (pred != golden["pred_best"]) & (margin > 2.0 * noise + 1e-5 * scale)
        assert not bad.any(), "argmax differs from the reference on %d nodes with a clear margin" % int(bad.sum())
    return agree


def check_curves(golden, result, sizes, slack=2):
    """Accuracy / R² curves over the 101 nuggets.  Accuracy may move by ``slack`` samples at a nugget
    (near-ties); R² by 1e-3 absolute or 3x the curve's own fp32 noise."""
    for key in ("train", "val", "test"):
        ref = golden["result_" + key]
        got = np.asarray(result[key], dtype=np.float32)
        assert got.shape == ref.shape
        if "pred_best" in golden:
            bound = (slack + 0.5) / max(sizes[key], 1)
            ok = np.abs(got - ref) <= bound
            # indefinite initial kernels (sigmoid) put poles of 1/(w + eps d) inside the grid: require the
            # val-selected nugget and 90 % of the grid
            assert ok[int(golden["best"])], "%s accuracy at the selected nugget off by > %d samples" % (key, slack)
            assert ok.mean() >= 0.9, "%s accuracy curve: only %.0f%% of nuggets within %d samples" % (key, 100 * ok.mean(), slack)
        else:
            ok = np.abs(got - ref) <= 2e-3 + 2e-3 * np.abs(ref)
            # the smallest nuggets amplify fp32 noise in the reference itself: require 90 % of the grid
            assert ok.mean() >= 0.9, "%s R2 curve: only %.0f%% of nuggets within 2e-3" % (key, 100 * ok.mean())


# ---------------------------------------------------------------------------------------------------------
# Full-size fixtures (tests/golden/big): samples of one run of the unmodified reference on the headline shapes
# ---------------------------------------------------------------------------------------------------------
BIG_DIR = os.path.join(GOLDEN_DIR, "big")


def load_big(name: str):
    path = os.path.join(BIG_DIR, name + ".npz")
    if not os.path.exists(path):
        return None
    return dict(np.load(path))


def check_big(golden, ksamp, fit_best_rows, fit_mid_rows, pred_all, curves, sizes, slack=2):
    """The checks of ``runner.check_against_golden`` on the sampled quantities of a big fixture.  Returns a report
    dict; argmax flips are COUNTED and reported, and only tolerated on nodes whose reference top-2 margin lies
    inside the reference's own fp32 noise (fp64 companion present) or below 1e-3 of the largest mean (absent)."""
    rep = {}
    rep["k_rel_fp32"] = rel_fro(ksamp, golden["ksamp"])
    assert rep["k_rel_fp32"] <= REL_FP32, "Q Q^T block vs the fp32 reference: rel %.3e" % rep["k_rel_fp32"]
    has64 = "ksamp64" in golden
    if has64:
        rep["k_rel_fp64"] = rel_fro(ksamp, golden["ksamp64"])
        assert rep["k_rel_fp64"] <= REL_FP32, "Q Q^T block vs the fp64 reference: rel %.3e" % rep["k_rel_fp64"]
    for key, value in (("fit_best", fit_best_rows), ("fit_mid", fit_mid_rows)):
        e32 = rel_fro(value, golden[key])
        if has64:
            floor = rel_fro(golden[key], golden[key + "64"])
            e64 = rel_fro(value, golden[key + "64"])
            bound = max(REL_FP32, 6.0 * floor)
            assert e64 <= bound, "%s vs fp64 reference: rel %.3e > %.3e (reference self-error %.2e)" % (key, e64, bound, floor)
            assert e32 <= 2 * bound, "%s vs fp32 reference: rel %.3e > %.3e" % (key, e32, 2 * bound)
            rep[key] = (e32, e64, floor)
        else:
            # no fp64 companion (the run does not fit the build container): the plain tolerance at nugget 0.1,
            # 10x that at the val-selected nugget when it is one of the small, noise-amplifying ones
            bound = REL_FP32 * (10.0 if key == "fit_best" and int(golden["best"]) < 25 else 1.0)
            assert e32 <= bound, "%s vs fp32 reference: rel %.3e > %.1e" % (key, e32, bound)
            rep[key] = (e32, None, None)
    pred_all = np.asarray(pred_all).astype(np.int64)
    flips = pred_all != golden["pred_best"].astype(np.int64)
    rep["argmax_flips"] = int(flips.sum())
    rep["argmax_nodes"] = int(flips.size)
    scale = float(golden["fit_scale"])
    if has64:
        noise = float(np.max(np.abs(golden["fit_best"].astype(np.float64) - golden["fit_best64"])))
        rep["reference_fp32_vs_fp64_flips"] = int((golden["pred_best"] != golden["pred_best64"]).sum())
        tol_margin = 2.0 * noise + 1e-5 * scale
    else:
        tol_margin = 1e-3 * scale
    clear = flips & (golden["margin_best"] > tol_margin)
    rep["argmax_flips_clear_margin"] = int(clear.sum())
    assert not clear.any(), "argmax differs from the reference on %d nodes with a clear margin" % int(clear.sum())
    for key in ("train", "val", "test"):
        ref = golden["result_" + key]
        got = np.asarray(curves[key], dtype=np.float32)
        bound = (slack + 0.5) / max(sizes[key], 1) + 1e-4
        ok = np.abs(got - ref) <= bound
        if "result64_" + key in golden:
            # the reference's fp32 curve is noise at the smallest nuggets (poles of 1 / (w + eps d) at rounding-noise
            # eigenvalues); a nugget also counts when the curve agrees with the reference's own fp64 run there
            ok64 = np.abs(got - golden["result64_" + key]) <= bound
            rep["curve_%s_nuggets_matching_fp32_run" % key] = int(ok.sum())
            rep["curve_%s_nuggets_matching_fp64_run" % key] = int(ok64.sum())
            ok = ok | ok64
        rep["curve_%s_max_abs" % key] = float(np.abs(got - ref).max())
        assert ok[int(golden["best"])], "%s accuracy at the selected nugget differs: %.5f vs %.5f" % (key, got[int(golden["best"])], ref[int(golden["best"])])
        assert ok.mean() >= 0.9, "%s accuracy curve: only %.0f%% of nuggets within tolerance" % (key, 100 * ok.mean())
    return rep
