"""GPU-box driver for ncu: ONE launch of the tcgen05 back-projection of the bench's fraction-10 step (a 16 384-row block of
`K[~mask] @ S`, K = N = 15 404) and one of the lower-triangle train Gram on a 4 096-column slice of the same factor."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnngp_b200.ops import default_ops  # noqa: E402

dev = torch.device("cuda:0")
ops = default_ops()
ops.set_gemm_engine(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 15404
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
a = ops.empty(rows, n, dev)
a.normal_()
s = ops.empty(n, n, dev)
s.normal_()
out = ops.gemm_nt(a, s)
torch.cuda.synchronize()
print("done", float(out[0, 0]))
