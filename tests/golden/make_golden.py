"""Generate the golden fixtures by running the UNMODIFIED reference (imported from
/root/reference/src) on the synthetic parity cases of ``tests/cases.py``.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py            # all cases -> tests/golden/*.npz

``gnngp.py`` of the reference imports ``datasets`` (PyG-dependent, absent here); a stub module that
provides ``get_data`` with the same body contract (``src/datasets.py:313-330``) is installed under
that name so ``gnngp.GNNGP`` imports unchanged.  ``compute.py`` and ``predict.py`` import as they are.

Stored per case: a sampled block of K (or Q Qᵀ) in fp32 and from an fp64 reference run, the
accuracy / R² curves over the 101 nuggets, the best-validation summary, the fitted values and
argmax predictions at that nugget, and a checksum of the generated inputs.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from gnngp_b200 import datasets as my_datasets  # noqa: E402

REF = "/root/reference/src"


def import_reference():
    stub = types.ModuleType("datasets")
    stub.get_data = my_datasets.get_data
    sys.modules["datasets"] = stub
    sys.path.insert(0, REF)
    import compute as ref_compute  # noqa
    import predict as ref_predict  # noqa
    import gnngp as ref_gnngp  # noqa
    return ref_compute, ref_predict, ref_gnngp


def run_reference(ref_gnngp, case, data, masks, nugget, fp64=False):
    nystrom = case["fraction"] is not None
    if fp64:
        torch.set_default_dtype(torch.float64)
        d = my_datasets.GraphData(data.x.double(), data.y, data.edge_index, data.edge_attr.double(),
                                  data.train_mask, data.val_mask, data.test_mask)
        if d.y.dtype == torch.int64:   # reference's one_hot(...).to(torch.float) is fp32: pass float targets
            d.y = torch.nn.functional.one_hot(d.y, int(d.y[d.train_mask].max()) + 1).double()
        else:
            d.y = d.y.double()
        nug = nugget.double()
    else:
        d, nug = data, nugget
    try:
        model = ref_gnngp.GNNGP(d, device=None, Nystrom=nystrom)
        if nystrom:
            model.mask["landmark"] = masks["landmark"]
        model.set_hyper_param(case["L"], case["sigma_b"], case["sigma_w"], initial=case["initial"],
                              method=case["method"], **case["params"])
        fit = model.predict(nug)
        kern = model.get_kernel()
        summ = model.get_summary()
        return model, fit, kern, summ
    finally:
        torch.set_default_dtype(torch.float32)


def make_one(name, ref_gnngp):
    case = cases.CASES[name]
    data, masks, nugget = cases.build_inputs(case)
    idx = cases.sample_index(data.num_nodes, case["seed"])
    tidx = torch.from_numpy(idx)
    out = {"checksum": cases.input_checksum(data, masks), "sample_index": idx}

    model, fit, kern, summ = run_reference(ref_gnngp, case, data, masks, nugget, fp64=False)
    block = kern[tidx] @ kern[tidx].T if case["fraction"] is not None else kern[tidx][:, tidx]
    out["ksamp"] = block.numpy().astype(np.float32)
    out["kdiag_mean"] = np.float64((kern ** 2).sum(1).mean() if case["fraction"] is not None else torch.diag(kern).mean())
    for key in ("train", "val", "test"):
        out["result_" + key] = model.result[key].numpy().astype(np.float32)
    best = int(torch.argmax(model.result["val"]))
    out["best"] = np.int64(best)
    out["fit_best"] = fit[best].numpy().astype(np.float32)
    # a mid-grid nugget too (index 50 = 1e-1): less sensitive than the val-selected one when that is tiny
    out["fit_mid"] = fit[50].numpy().astype(np.float32)
    if data.y.dtype == torch.int64:
        out["pred_best"] = torch.argmax(fit[best], 1).numpy().astype(np.int16)

    _, fit64, kern64, _ = run_reference(ref_gnngp, case, data, masks, nugget, fp64=True)
    block64 = kern64[tidx] @ kern64[tidx].T if case["fraction"] is not None else kern64[tidx][:, tidx]
    out["ksamp64"] = block64.numpy().astype(np.float64)
    out["fit_best64"] = fit64[best].numpy().astype(np.float64)
    out["fit_mid64"] = fit64[50].numpy().astype(np.float64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    rel = np.linalg.norm(out["ksamp"] - out["ksamp64"]) / np.linalg.norm(out["ksamp64"])
    print("%-34s best=%3d val=%.4f test=%.4f  fp32-vs-fp64 K rel=%.2e" %
          (name, best, float(summ["val"]), float(summ["test"]), rel), flush=True)


def main(argv):
    _, _, ref_gnngp = import_reference()
    names = argv[1:] or list(cases.CASES)
    torch.manual_seed(0)
    for name in names:
        make_one(name, ref_gnngp)


if __name__ == "__main__":
    main(sys.argv)
