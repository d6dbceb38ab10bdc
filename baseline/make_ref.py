"""Vendor the reference's hot-path modules, UNMODIFIED, into ``baseline/_ref/`` (git-ignored; it travels to
the GPU box with the repo snapshot, where ``/root/reference`` does not exist).

    python baseline/make_ref.py            # run in the build container; __graft_entry__.build() calls it

The reference has no ``setup.py`` / ``pyproject.toml`` (nothing for ``pip install`` to build) and its package
entry imports torch_geometric / torch_sparse / torch_scatter, none of which is in the image or the wheelhouse.
The GP path itself needs only three of its files, which import nothing but torch and numpy:

    src/gnngp.py   src/compute.py   src/predict.py      (byte-for-byte copies, sha256 recorded in MANIFEST.json)

``gnngp.py`` does ``from datasets import get_data``; the reference's ``datasets.py`` cannot be imported (PyG), so
``bench.py`` / ``tests`` install a stub module of that name whose ``get_data`` is this repo's restatement of
``src/datasets.py:313-330`` (18 lines: build the sparse COO adjacency and the mask dict) before importing it.
Nothing under ``baseline/_ref`` is imported by the product; it is the ``--impl reference`` / ``gpu_reference``
arm of ``bench.py`` only.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src"
DST = os.path.join(HERE, "_ref")
FILES = ("gnngp.py", "compute.py", "predict.py")


def make(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print("baseline/make_ref: %s not present (GPU box); using the prebuilt baseline/_ref" % SRC)
        return os.path.isdir(DST)
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for name in FILES:
        shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
        with open(os.path.join(DST, name), "rb") as fh:
            manifest[name] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "sha256": manifest}, fh, indent=1)
    if verbose:
        print("baseline/_ref: %s copied unmodified from %s" % (", ".join(FILES), SRC))
    return True


def import_reference():
    """(compute, predict, gnngp) modules of the vendored reference, with the ``datasets.get_data`` stub installed."""
    import types
    if not all(os.path.exists(os.path.join(DST, f)) for f in FILES):
        raise RuntimeError("baseline/_ref is missing; run `python baseline/make_ref.py` in the build container")
    root = os.path.dirname(HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    from gnngp_b200 import datasets as my_datasets
    stub = types.ModuleType("datasets")
    stub.get_data = my_datasets.get_data
    saved = {k: sys.modules.get(k) for k in ("datasets", "compute", "predict", "gnngp")}
    sys.modules["datasets"] = stub
    sys.path.insert(0, DST)
    try:
        for k in ("compute", "predict", "gnngp"):
            sys.modules.pop(k, None)
        import compute as ref_compute
        import predict as ref_predict
        import gnngp as ref_gnngp
    finally:
        sys.path.remove(DST)
        for k, v in saved.items():            # do not leave modules named like the product's drop-in shims behind
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)
    return ref_compute, ref_predict, ref_gnngp


if __name__ == "__main__":
    make()
