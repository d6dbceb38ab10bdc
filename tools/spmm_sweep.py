"""GPU-box probe: CSR SpMM time against tile width / nonzeros per warp on the Reddit and arxiv shapes."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnngp_b200 import synth            # noqa: E402
from gnngp_b200.ops import default_ops  # noqa: E402

dev = torch.device("cuda:0")
ops = default_ops()
out = {}


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for shape, ks in (("reddit", (602, 1024, 3024)), ("arxiv", (128, 2048))):
    data = synth.make_graph(shape, seed=0, device=dev)
    n = data.num_nodes
    csr = ops.csr_from_coo(data.edge_index[0], data.edge_index[1], data.edge_attr, n, n)
    deg = csr.rowptr[1:] - csr.rowptr[:-1]
    print(shape, "nnz", csr.nnz, "max deg", int(deg.max()), "mean", float(deg.float().mean()), flush=True)
    for k in ks:
        b = torch.randn(n, (k + 3) // 4 * 4, device=dev)[:, :k]
        c = ops.empty(n, k, dev)
        for tile in (64, 128):
            for chunk in (256, 512):
                ops.spmm_tile_cols, ops.spmm_nnz_per_warp = tile, chunk
                ms = timeit(lambda: ops.spmm(csr, b, 1.0, out=c))
                gather_tbs = 4.0 * csr.nnz * k / (ms * 1e-3) / 1e12
                out["%s k=%d tile=%d chunk=%d" % (shape, k, tile, chunk)] = ms
                print("%s k=%4d tile=%3d chunk=%4d  %8.3f ms  gather %.2f TB/s" % (shape, k, tile, chunk, ms, gather_tbs), flush=True)
        for chunk in (128, 256, 512, 1024, 2048):
            ops.spmm_tile_cols, ops.spmm_nnz_per_warp = 0, chunk
            ms = timeit(lambda: ops.spmm(csr, b, 1.0, out=c))
            out["%s k=%d panel chunk=%d" % (shape, k, chunk)] = ms
            print("%s k=%4d PANEL    chunk=%4d  %8.3f ms  gather %.2f TB/s" % (shape, k, chunk, ms, 4.0 * csr.nnz * k / (ms * 1e-3) / 1e12), flush=True)
        ops.spmm_tile_cols, ops.spmm_nnz_per_warp = 0, 0
        # cuSPARSE through torch, for context only (the reference's GPU path)
        a = torch.sparse_coo_tensor(data.edge_index, data.edge_attr, (n, n)).coalesce()
        bc = b.contiguous()
        ms = timeit(lambda: torch.sparse.mm(a, bc))
        print("%s k=%4d torch.sparse.mm (coalesced COO, cuSPARSE): %8.3f ms" % (shape, k, ms), flush=True)
        acsr = a.to_sparse_csr()
        ms = timeit(lambda: acsr @ bc)
        print("%s k=%4d torch CSR @ dense (cuSPARSE): %8.3f ms" % (shape, k, ms), flush=True)
        out["%s k=%d cusparse_csr" % (shape, k)] = ms
        del a, acsr, b, c, bc
    del data, csr
    torch.cuda.empty_cache()

with open(os.path.join(ROOT, "gpurun_out", "spmm_sweep.json"), "w") as fh:
    json.dump(out, fh, indent=1)
