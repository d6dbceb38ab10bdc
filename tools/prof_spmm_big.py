"""GPU-box driver for ncu: ONE launch of the SpMM at a layer-2 shape of the bench:  prof_spmm_big.py [k] [shape]
(default: k = 3 024 on the Reddit shape; `9200 arxiv` is the short-row case)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnngp_b200 import synth            # noqa: E402
from gnngp_b200.ops import default_ops  # noqa: E402

dev = torch.device("cuda:0")
ops = default_ops()
k = int(sys.argv[1]) if len(sys.argv) > 1 else 3024
data = synth.make_graph(sys.argv[2] if len(sys.argv) > 2 else "reddit", seed=0, device=dev)
n = data.num_nodes
csr = ops.csr_from_coo(data.edge_index[0], data.edge_index[1], data.edge_attr, n, n)
b = ops.empty(n, k, dev)
b.normal_()
c = ops.empty(n, k + 1, dev)
ops.spmm(csr, b, 1.0, out=c[:, :k])
torch.cuda.synchronize()
print("done")
