"""torchrun script: row-sharded logistic regression over NCCL vs a single-GPU engine, and
chain-sharded weak scaling sanity.  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node=N --master-addr 127.0.0.1 \
      --master-port 29611 scripts/multi_gpu_check.py --rows [--big]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import turing_jl_b200 as T
from turing_jl_b200 import distributed as D


def diag_mode(rank, world, local):
    """chains sharded over the ranks: pooled (cross-GPU) R-hat / bulk ESS == single-engine values"""
    model = T.EightSchools()
    spl = T.NUTS(100, 0.8)
    C = 96
    eng = D.make_chain_sharded_engine(model, spl, C, 100, 5, rank, world, local)
    D.attach_comm(eng, dist, device="cuda:%d" % local)
    eng.init()
    eng.run(100 + 299, 100, 1)
    rhat, ess = eng.diagnostics()          # collective: every rank gets the pooled values
    ok = True
    if rank == 0:
        single = T.Engine(model, spl, C, 100, 5, device=local)
        single.init()
        single.run(100 + 299, 100, 1)
        r1, e1 = single.diagnostics()
        single.close()
        dr, de = np.abs(rhat - r1).max(), np.abs(ess / e1 - 1).max()
        print("pooled diagnostics over %d ranks vs one engine: max |d rhat| %.2e, max rel d ess %.2e" % (world, dr, de))
        ok = dr < 1e-9 and de < 1e-6
    eng.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and flag.item() == 1:
        print("DIAG-OK")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


def mesh_mode(rank, world, local):
    """chains x rows mesh (SURVEY 8e "Hybrid"): world = chain_groups x 2 row shards.  Ranks of one
    chain group must hold identical draws; the union over the groups must equal (to the rounding of
    the regrouped row sums) a single-GPU engine running all chains on all rows."""
    rs = 2
    N, Dm, C = 4001, 24, 96
    model = T.models.synthetic_logistic(N, Dm, seed=3)
    spl = T.NUTS(10, 0.8)
    n_tr = 14
    eng = D.make_mesh_engine(model, spl, C, 10, 7, rank, world, local, dist, rs)
    cg, rr, ng = eng.mesh
    eng.set_option("keep_theta", 1)
    eng.init()
    eng.run(n_tr, 1, 1)
    out = eng.fetch(theta=True)
    th = torch.from_numpy(out["theta"].copy()).cuda()          # [K, D, C_group]
    b = D.chain_bounds(C, ng)
    # gather every rank's draws on every rank (groups may differ in size by one chain: pad)
    cmax = max(b[g + 1] - b[g] for g in range(ng))
    pad = torch.zeros(th.shape[0], th.shape[1], cmax, dtype=th.dtype, device="cuda")
    pad[:, :, :th.shape[2]] = th
    allth = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(allth, pad)
    ok = True
    for g in range(ng):
        ok = ok and bool(torch.equal(allth[g * rs], allth[g * rs + 1]))
    if rank == 0:
        print("mesh %d chain groups x %d row shards: identical within every group: %s" % (ng, rs, ok))
        single = T.Engine(model, spl, C, 10, 7, device=local, strategy=2)
        single.set_option("keep_theta", 1)
        single.init()
        single.run(n_tr, 1, 1)
        o1 = single.fetch(theta=True)
        union = np.concatenate([allth[g * rs][:, :, :b[g + 1] - b[g]].cpu().numpy() for g in range(ng)], axis=2)
        d = np.abs(o1["theta"][:6] - union[:6]).max()
        print("  union of the groups vs single GPU: max |dtheta| first 6 transitions %.2e" % d)
        ok = ok and d < 1e-8
        single.close()
    eng.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and flag.item() == 1:
        print("MESH-OK")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if "--diag" in sys.argv:
        return diag_mode(rank, world, local)
    if "--mesh" in sys.argv:
        return mesh_mode(rank, world, local)
    big = "--big" in sys.argv
    N, Dm, C = (2_000_000, 256, 256) if big else (4001, 24, 64)
    model = T.models.synthetic_logistic(N, Dm, seed=3)
    spl = T.NUTS(10, 0.8)
    n_tr = 14
    eng = D.make_row_sharded_engine(model, spl, C, 10, 7, rank, world, local, dist)
    eng.set_option("keep_theta", 1)
    tc = "--tc" in sys.argv      # tcgen05 likelihood on every rank's row slice
    if tc:
        eng.set_option("logistic_tc", 1)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    eng.init()
    eng.run(n_tr, 1, 1)
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    out = eng.fetch(theta=True)
    grads = eng.grad_evals
    # every rank holds every chain and must have produced identical draws
    t = torch.from_numpy(out["theta"].copy()).cuda()
    ref = t.clone()
    dist.broadcast(ref, src=0)
    same_across_ranks = bool(torch.equal(t, ref))
    ok = same_across_ranks
    if rank == 0:
        print("row-sharded: world %d N=%d D=%d C=%d: %.2fs, %d grad evals (%.3e grad/s, %.3e flop/s), identical across ranks: %s"
              % (world, N, Dm, C, dt, grads, grads / dt, grads / dt * 4 * N * Dm, same_across_ranks))
        if not big:
            single = T.Engine(model, spl, C, 10, 7, device=local, strategy=2)
            single.set_option("keep_theta", 1)
            if tc:
                single.set_option("logistic_tc", 1)
            single.init()
            single.run(n_tr, 1, 1)
            o1 = single.fetch(theta=True)
            d = np.abs(o1["theta"][:6] - out["theta"][:6]).max()
            same = (o1["stats"][:, [0, 6], :] == out["stats"][:, [0, 6], :]).mean()
            print("  vs single GPU: max |dtheta| first 6 transitions %.2e, identical tree stats %.3f" % (d, same))
            # tcgen05 mode: the ranks' fp32 partial sums group differently from the single-GPU cut
            ok = ok and (d < 1e-2 and same > 0.8 if tc else d < 1e-8 and same > 0.95)
            single.close()
    eng.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and flag.item() == 1:
        print("ROWS-OK")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
