"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU modes (sharding bounds,
combination of per-shard diagnostics and moments).  The NCCL data path itself needs GPUs and
is covered by tests/test_gpu_multi.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import turing_jl_b200 as T
    from turing_jl_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        full = rng.standard_normal((50, 3, 10))            # [iters, params, chains]
        b = D.chain_bounds(10, world)
        mine = full[:, :, b[rank]:b[rank + 1]]
        mean, var = D.gather_param_moments(dist, mine)
        rh, ess = D.combine_diagnostics(dist, [1.0 + 0.01 * rank] * 3, [100.0 * (rank + 1)] * 3)
        model = T.models.synthetic_logistic(101, 4, seed=3)
        local = D.shard_rows(model, world, rank)
        import torch
        n = torch.tensor([local.N])
        dist.all_reduce(n)
        q.put((rank, mean, var, rh, ess, int(n.item()), local.N))
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    full = rng.standard_normal((50, 3, 10))
    flat = full.transpose(1, 0, 2).reshape(3, -1)
    for rank, mean, var, rh, ess, ntot, nloc in res:
        np.testing.assert_allclose(mean, flat.mean(axis=1), rtol=1e-12)
        np.testing.assert_allclose(var, flat.var(axis=1, ddof=1), rtol=1e-10)
        np.testing.assert_allclose(rh, [1.01] * 3)
        np.testing.assert_allclose(ess, [300.0] * 3)
        assert ntot == 101 and nloc in (50, 51)


def test_bounds():
    sys.path.insert(0, ROOT)
    from turing_jl_b200 import distributed as D
    assert D.chain_bounds(65536, 8) == [8192 * i for i in range(9)]
    b = D.chain_bounds(10, 4)
    assert b[0] == 0 and b[-1] == 10 and all(2 <= b[i + 1] - b[i] <= 3 for i in range(4))
    assert D.row_bounds(50_000_000, 8)[1] == 6_250_000


def _mesh_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from turing_jl_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cg, rr, ng = D.mesh_coords(rank, world, 2)
        grp, leader = D.mesh_row_group(dist, rank, world, 2)
        # every row group gets its own id from its own leader (a fake id here: no GPU)
        uid = D.broadcast_unique_id(dist, rank, None, src=leader, group=grp,
                                    make_id=lambda: np.full(128, 10 + rank, dtype=np.uint8))
        q.put((rank, cg, rr, ng, leader, int(uid[0]), int(uid[-1])))
    finally:
        dist.destroy_process_group()


def test_mesh_groups_four_ranks():
    """chains x rows mesh 2 x 2 over gloo: coordinates and the per-row-group unique id"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_mesh_worker, args=(r, 4, port, q)) for r in range(4)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 0), (0, 1), (1, 0), (1, 1)]
    assert [r[4] for r in res] == [0, 0, 2, 2]
    assert [r[5] for r in res] == [10, 10, 12, 12] and all(r[5] == r[6] for r in res)
    sys.path.insert(0, ROOT)
    from turing_jl_b200 import distributed as D
    with pytest.raises(ValueError):
        D.mesh_coords(0, 6, 4)
