"""The C-ABI library loads and exports every symbol include/gnngp_b200.h declares (no compute calls: no GPU here)."""
import ctypes
import os

import pytest

from gnngp_b200 import _lib


def test_header_declares_the_documented_entry_points():
    protos = _lib.parse_header()
    for name in ("gnngp_csr_from_coo", "gnngp_spmm_csr_f32", "gnngp_gram_arccos_f32", "gnngp_gemm_nt_f32",
                 "gnngp_nugget_sweep_f32", "gnngp_argmax_count", "gnngp_arccos_dense_f32", "gnngp_last_error"):
        assert name in protos
    assert len(protos) >= 29


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        from gnngp_b200.build import build
        build(verbose=False)
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    missing = [name for name in _lib.parse_header() if not hasattr(cdll, name)]
    assert not missing, "declared in include/gnngp_b200.h but not exported: %s" % missing
    lib = _lib.Library()
    assert lib.cdll.gnngp_abi_version() == 1
    assert lib.query("gnngp_gemm_workspace_bytes", 128, 256, 32) >= (128 + 256) * 32 * 8
    assert lib.query("gnngp_csr_from_coo_workspace_bytes", 1000, 10) > 24000


def test_product_fails_loudly_without_cuda_or_library(monkeypatch, tmp_path):
    import torch
    from gnngp_b200 import ops as ops_mod
    with pytest.raises(RuntimeError, match="missing"):
        _lib.Library(str(tmp_path / "nope.so"))
    if not torch.cuda.is_available():
        o = ops_mod.default_ops()
        with pytest.raises(RuntimeError, match="CUDA tensor"):
            o.row_norms(torch.zeros(3, 3))
        from gnngp_b200 import compute
        with pytest.raises(RuntimeError):
            compute._init_kernel(torch.zeros(4, 4), "linear")


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "gnn-gp_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in text and "from oracle" not in text and "stub_ops" not in text, fn
