"""The row-sharded orchestration (gnngp_b200/shard.py + pipeline.py collectives) under world_size 2
and 3 with the gloo backend on CPU, driven by the test-only CPU operator stand-in, against the same
pipeline on one rank.  The CUDA kernels are not involved here; the NCCL path is measured by bench.py."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, name, out_dir, solver="auto"):
    os.environ["GNNGP_FIT_SOLVER"] = solver
    if solver == "krylov":                                   # both Krylov routes, with a small Lanczos block for the root
        os.environ["GNNGP_ROOT_SOLVER"] = "krylov"
        os.environ["GNNGP_ROOT_BLOCK"] = "16"
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import cases
    from stub_ops import StubOps
    from gnngp_b200.shard import ShardedGNNGP
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    case = cases.CASES[name]
    data, masks, nugget = cases.build_inputs(case)
    model = ShardedGNNGP(data, device=None, Nystrom=True, ops=StubOps())
    model.mask["landmark"] = masks["landmark"]
    model.set_hyper_param(case["L"], case["sigma_b"], case["sigma_w"], initial=case["initial"], method=case["method"],
                          **case["params"])
    model.predict(nugget)
    fit = model.gather_fit()
    q_full = model.comm.all_gather_rows(model.Q, model.layout)
    summ = model.get_summary()
    if rank == 0:
        torch.save({"fit": fit, "Q": q_full.clone(), "result": model.result,
                    "summary": {k: float(summ[k]) for k in ("nugget", "train", "val", "test")},
                    "rows": (model.layout.begin, model.layout.end)}, os.path.join(out_dir, "rank0.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,name,solver", [(2, "tiny-GCN-nys2-L3", "auto"), (2, "tiny-SAGE-nys2-L2", "auto"),
                                               (3, "tiny-GCN2-nys2-L2", "auto"), (2, "tinyreg-GIN-nys1-L3", "auto"),
                                               (2, "tiny-GCN-nys8-L2", "auto"), (2, "tiny-GCN-nys2-arccos", "auto"),
                                               # nugget-parallel fit: every rank solves its share of the nugget grid
                                               (3, "tiny-GCN-nys2-L3", "cholesky"), (2, "tinyreg-GIN-nys1-L3", "cholesky"),
                                               # replicated Krylov root / Krylov fit with the common convergence decision
                                               (2, "small-GCN-nys4-L2", "krylov"), (3, "small-SAGE-nys4-L4", "krylov")])
def test_sharded_matches_single_rank(tmp_path, world, name, solver):
    import cases
    import runner
    from stub_ops import StubOps
    mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path), solver), nprocs=world, join=True)
    got = torch.load(os.path.join(str(tmp_path), "rank0.pt"))
    case = cases.CASES[name]
    data, masks, nugget = cases.build_inputs(case)
    ref = runner.run_pipeline(StubOps(), case, data, masks, nugget)
    kq, kr = got["Q"] @ got["Q"].T, ref["Q"] @ ref["Q"].T
    assert float((kq - kr).norm() / kr.norm()) < 1e-4
    assert got["fit"].shape == ref["fit"].shape
    assert float((got["fit"][50] - ref["fit"][50]).norm() / ref["fit"][50].norm()) < 1e-3
    for key in ("train", "val", "test", "landmark"):
        assert float((got["result"][key] - ref["result"][key]).abs().max()) < 0.03
    # against the reference's golden outputs too
    runner.check_against_golden(name, {"Q": got["Q"], "fit": got["fit"], "result": got["result"]}, masks)


def test_row_layout_partitions_every_row_once():
    from gnngp_b200.pipeline import RowLayout
    for n in (1, 7, 257, 1000):
        for world in (1, 2, 3, 8):
            spans = [(RowLayout(n, world, r).begin, RowLayout(n, world, r).end) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            mask = torch.zeros(n, dtype=torch.bool)
            mask[::3] = True
            pos_all = []
            for r in range(world):
                rows, pos, count = RowLayout(n, world, r).local_index(mask)
                assert count == int(mask.sum())
                pos_all.append(pos)
                assert bool(mask[rows + RowLayout(n, world, r).begin].all())
            assert torch.equal(torch.cat(pos_all), torch.arange(int(mask.sum())))


def _rowsplit_worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    from stub_ops import StubOps
    from gnngp_b200.krylov_root import RowSplit
    from gnngp_b200.shard import TorchComm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gen = torch.Generator().manual_seed(4)
    n, c, k = 301, 7, 256
    m = torch.randn(n, n, generator=gen, dtype=torch.float64)
    v = torch.randn(n, c, generator=gen, dtype=torch.float64)
    zq, _ = torch.linalg.qr(torch.randn(n, 40, generator=gen, dtype=torch.float64))
    z = torch.zeros(n, k, dtype=torch.float64)
    z[:, :40] = zq
    ops = StubOps()
    split = RowSplit(ops, n, TorchComm())
    out = torch.empty(n, c, dtype=torch.float64)
    split.matmul(m, v, out)
    x = m.clone()
    split.rotate(x, z.clone())
    want_x = m.clone()
    ops.qr_rotate(want_x, z.clone())
    ok = bool(torch.allclose(out, m @ v, rtol=1e-12, atol=1e-12)) and bool(torch.allclose(x, want_x, rtol=1e-10, atol=1e-10))
    torch.save({"ok": ok, "rows": (split.r0, split.r1)}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_split_products_and_rotation(tmp_path, world):
    """krylov_root.RowSplit: the rank-parallel product / rotation of the replicated Krylov routes equals the local one
    on every rank (odd row count, last rank short)."""
    mp.spawn(_rowsplit_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    spans = []
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r))
        assert got["ok"], r
        spans.append(got["rows"])
    assert spans[0][0] == 0 and spans[-1][1] == 301 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
