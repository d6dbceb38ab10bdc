"""fp64 companion of a big golden fixture, from the UNMODIFIED reference (baseline/_ref) run in float64 on the GPU
box (CUDA tensors): the arxiv fraction-10 run needs ~74 GB in fp64, which the build container does not have.

    python tools/golden_fp64_on_gpu.py arxiv-GCN-nys10-L2      ->  gpurun_out/<name>_fp64.npz

The keys (ksamp64, fit_best64, fit_mid64, pred_best64) are then merged into tests/golden/big/<name>.npz with
`python tools/golden_fp64_on_gpu.py --merge <name>` in the build container."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(argv):
    if argv[1] == "--merge":
        name = argv[2]
        path = os.path.join(ROOT, "tests", "golden", "big", name + ".npz")
        base = dict(np.load(path))
        extra = dict(np.load(os.path.join(ROOT, "gpurun_out", name + "_fp64.npz")))
        assert np.array_equal(base["sample_index"], extra["sample_index"]) and int(base["best"]) == int(extra["best"])
        for k in ("ksamp64", "fit_best64", "fit_mid64", "pred_best64", "result64_train", "result64_val", "result64_test"):
            base[k] = extra[k]
        base["fp64_companion_source"] = np.asarray("unmodified reference in float64 on CUDA tensors (tools/golden_fp64_on_gpu.py)")
        np.savez_compressed(path, **base)
        rel = np.linalg.norm(base["ksamp"] - base["ksamp64"]) / np.linalg.norm(base["ksamp64"])
        print("merged; reference fp32-vs-fp64: K rel %.2e, fit_best rel %.2e, fit_mid rel %.2e, argmax flips %d" % (
            rel, np.linalg.norm(base["fit_best"] - base["fit_best64"]) / np.linalg.norm(base["fit_best64"]),
            np.linalg.norm(base["fit_mid"] - base["fit_mid64"]) / np.linalg.norm(base["fit_mid64"]),
            int((base["pred_best"] != base["pred_best64"]).sum())))
        return
    import cases
    from baseline.make_ref import import_reference
    from gnngp_b200 import datasets as my_datasets
    name = argv[1]
    case = cases.BIG[name]
    golden = dict(np.load(os.path.join(ROOT, "tests", "golden", "big", name + ".npz")))
    _, _, ref_gnngp = import_reference()
    data, masks, nugget = cases.build_inputs(case)
    np.testing.assert_allclose(cases.input_checksum(data, masks), golden["checksum"], rtol=1e-7)
    dev = torch.device("cuda:0")
    torch.set_default_dtype(torch.float64)
    d = my_datasets.GraphData(data.x.double().to(dev), None, data.edge_index.to(dev), data.edge_attr.double().to(dev),
                              data.train_mask.to(dev), data.val_mask.to(dev), data.test_mask.to(dev))
    d.y = torch.nn.functional.one_hot(data.y, int(data.y[data.train_mask].max()) + 1).double().to(dev)   # see make_golden.py
    model = ref_gnngp.GNNGP(d, device=dev, Nystrom=True)
    model.mask["landmark"] = masks["landmark"].to(dev)
    model.set_hyper_param(case["L"], case["sigma_b"], case["sigma_w"], initial=case["initial"], method=case["method"], **case["params"])
    fit = model.predict(nugget.double().to(dev))
    kern = model.get_kernel()
    idx = torch.from_numpy(golden["sample_index"]).to(dev)
    rows = torch.from_numpy(golden["row_sample"]).to(dev)
    best = int(golden["best"])
    out = {"sample_index": golden["sample_index"], "best": golden["best"],
           "ksamp64": (kern[idx] @ kern[idx].T).cpu().numpy(), "fit_best64": fit[best][rows].cpu().numpy(),
           "fit_mid64": fit[50][rows].cpu().numpy(), "pred_best64": torch.argmax(fit[best], 1).cpu().numpy().astype(np.int16)}
    # accuracy curves of the fp64 run (its own `result` reports R^2 because the targets had to be passed as floats)
    labels = data.y.to(dev)
    for key in ("train", "val", "test"):
        mask = getattr(data, key + "_mask").to(dev)
        acc = [(torch.argmax(fit[g][mask], 1) == labels[mask]).double().mean().item() for g in range(fit.shape[0])]
        out["result64_" + key] = np.asarray(acc, dtype=np.float32)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", name + "_fp64.npz"), **out)
    print("wrote gpurun_out/%s_fp64.npz; peak GPU memory %.1f GB" % (name, torch.cuda.max_memory_allocated() / 1e9))


if __name__ == "__main__":
    main(sys.argv)
