"""GPU-box driver for ncu: one launch each of the hot kernels on Reddit-shape operands (kept short: ncu
replays every profiled launch ~40 times).  Run it plain first, then under ncu with -k regex filters."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnngp_b200 import synth            # noqa: E402
from gnngp_b200.ops import default_ops  # noqa: E402

dev = torch.device("cuda:0")
ops = default_ops()
shape = sys.argv[1] if len(sys.argv) > 1 else "reddit"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 256
data = synth.make_graph(shape, seed=0, device=dev)
n = data.num_nodes
csr = ops.csr_from_coo(data.edge_index[0], data.edge_index[1], data.edge_attr, n, n)
b = torch.randn(n, k, device=dev)
c = ops.empty(n, k, dev)
for _ in range(2):
    ops.spmm(csr, b, 1.0, out=c)
torch.cuda.synchronize()

# dense kernels on a 32 768-row block of a Reddit-shape factor (m = 3 107, N_a = 3 106, G*C = 4 141)
rows, na = 32768, 3106
q = torch.randn(rows, na + 1, device=dev)
l = q[:na].contiguous()
s, t = ops.row_norms(q), ops.row_norms(l)
ops.set_gemm_engine(1)
for _ in range(2):
    e = ops.gram_arccos(q, l, s, t)
    w = ops.gemm_nt(e, l[:, :na].contiguous())
    f = ops.nugget_sweep(q, torch.randn(101 * 41, na + 1, device=dev), 101, 41)
torch.cuda.synchronize()
print("done")
