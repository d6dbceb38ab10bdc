"""GPU-box probe: the one-CTA block factorisation + inverse (potrf_inv_block) and the blocked Cholesky at the benchmark size."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnngp_b200.ops import default_ops  # noqa: E402

dev = torch.device("cuda:0")
ops = default_ops()


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


x = torch.randn(128, 200, device=dev, dtype=torch.float64)
spd = x @ x.T + torch.eye(128, device=dev, dtype=torch.float64)
info = torch.zeros(1, dtype=torch.int32, device=dev)
work = spd.clone()


def block():
    work.copy_(spd)
    ops.potrf_inv_block(work, info)


t_copy = timeit(lambda: work.copy_(spd), 200)
print("potrf_inv_block 128: %.1f us (copy %.1f us)" % (1e3 * (timeit(block, 200) - t_copy), 1e3 * t_copy), flush=True)
for n in (3052, 15404):
    y = torch.randn(n, n + 64, device=dev, dtype=torch.float64)
    m = y @ y.T
    del y
    buf = torch.empty_like(m)

    def own():
        buf.copy_(m)
        ops.potrf_f64(buf, info)

    def lib():
        torch.linalg.cholesky(m)

    t_c = timeit(lambda: buf.copy_(m), 3)
    print("potrf n=%d: own %.1f ms, cusolver %.1f ms" % (n, timeit(own, 3) - t_c, timeit(lib, 3)), flush=True)
    del m, buf
    torch.cuda.empty_cache()
