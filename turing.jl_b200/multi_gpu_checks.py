"""Equality checks of the multi-GPU modes against a single-GPU engine.

Every function is collective over an ALREADY INITIALISED `torch.distributed` NCCL process group
(one rank per GPU), returns `(ok, details)` with the same value on every rank, and compares the
sharded run with ONE engine that runs all chains on all rows on rank 0's GPU: no oracle involved,
these are self-consistency checks of the product (the single-GPU engine is what tests/ compare with
the oracle).  Used by scripts/multi_gpu_check.py (pytest, torchrun) and by bench.py --gpus N, which
puts the outcome into its JSON line as `parity_checks`.
"""
import numpy as np

from . import distributed as D
from .engine import Engine
from .models import EightSchools, synthetic_logistic
from .samplers import NUTS


def _all_ok(dist, ok):
    import torch
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item() == 1)


def rows_check(dist, rank, world, local, tc=False, big=False, wide=False):
    """rows sharded over all ranks, per-leapfrog NCCL all-reduce: identical chains on every rank and
    (fp64 path) equal to the single-GPU engine to rounding over the first transitions.
    wide: 128 < D <= 256, where the tcgen05 mode runs the CTA-pair kernel (config 5's regime)"""
    import torch
    N, Dm, C = (2_000_000, 256, 256) if big else ((6001, 160, 64) if wide else (4001, 24, 64))
    model = synthetic_logistic(N, Dm, seed=3)
    spl = NUTS(10, 0.8)
    n_tr = 14
    eng = D.make_row_sharded_engine(model, spl, C, 10, 7, rank, world, local, dist)
    eng.set_option("keep_theta", 1)
    if tc:
        eng.set_option("logistic_tc", 1)
    eng.init()
    eng.run(n_tr, 1, 1)
    out = eng.fetch(theta=True)
    t = torch.from_numpy(out["theta"].copy()).cuda()
    ref = t.clone()
    dist.broadcast(ref, src=0)
    same_across_ranks = bool(torch.equal(t, ref))
    ok = same_across_ranks
    det = {"world": world, "N": N, "D": Dm, "chains": C, "identical_across_ranks": same_across_ranks,
           "grad_evals": int(eng.grad_evals)}
    if rank == 0 and not big:
        single = Engine(model, spl, C, 10, 7, device=local, strategy=2)
        single.set_option("keep_theta", 1)
        if tc:
            single.set_option("logistic_tc", 1)
        single.init()
        single.run(n_tr, 1, 1)
        o1 = single.fetch(theta=True)
        d = float(np.abs(o1["theta"][:6] - out["theta"][:6]).max())
        same = float((o1["stats"][:, [0, 6], :] == out["stats"][:, [0, 6], :]).mean())
        det["max_abs_dtheta_first6_vs_single_gpu"] = d
        det["identical_tree_stats_vs_single_gpu"] = same
        # tcgen05 mode: the ranks' fp32 partial sums group differently from the single-GPU cut
        ok = ok and ((d < 1e-2 and same > 0.8) if tc else (d < 1e-8 and same > 0.95))
        single.close()
    eng.close()
    ok = _all_ok(dist, ok)
    det["ok"] = ok
    return ok, det


def diag_check(dist, rank, world, local):
    """chains sharded over the ranks: pooled (cross-GPU) R-hat / bulk ESS == single-engine values"""
    model = EightSchools()
    spl = NUTS(100, 0.8)
    C = 96
    eng = D.make_chain_sharded_engine(model, spl, C, 100, 5, rank, world, local)
    D.attach_comm(eng, dist, device="cuda:%d" % local)
    eng.init()
    eng.run(100 + 299, 100, 1)
    rhat, ess = eng.diagnostics()          # collective: every rank gets the pooled values
    ok = True
    det = {"world": world, "chains": C}
    if rank == 0:
        single = Engine(model, spl, C, 100, 5, device=local)
        single.init()
        single.run(100 + 299, 100, 1)
        r1, e1 = single.diagnostics()
        single.close()
        dr, de = float(np.abs(rhat - r1).max()), float(np.abs(ess / e1 - 1).max())
        det["max_abs_drhat"], det["max_rel_dess"] = dr, de
        ok = dr < 1e-9 and de < 1e-6
    eng.close()
    ok = _all_ok(dist, ok)
    det["ok"] = ok
    return ok, det


def mesh_check(dist, rank, world, local, row_shards=2):
    """chains x rows mesh (SURVEY 8e "Hybrid"): world = chain_groups x row_shards.  Ranks of one
    chain group must hold identical draws; the union over the groups must equal (to the rounding of
    the regrouped row sums) a single-GPU engine running all chains on all rows."""
    import torch
    rs = row_shards
    N, Dm, C = 4001, 24, 96
    model = synthetic_logistic(N, Dm, seed=3)
    spl = NUTS(10, 0.8)
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
        for r in range(1, rs):
            ok = ok and bool(torch.equal(allth[g * rs], allth[g * rs + r]))
    det = {"world": world, "chain_groups": ng, "row_shards": rs, "identical_within_groups": ok}
    if rank == 0:
        single = Engine(model, spl, C, 10, 7, device=local, strategy=2)
        single.set_option("keep_theta", 1)
        single.init()
        single.run(n_tr, 1, 1)
        o1 = single.fetch(theta=True)
        union = np.concatenate([allth[g * rs][:, :, :b[g + 1] - b[g]].cpu().numpy() for g in range(ng)], axis=2)
        d = float(np.abs(o1["theta"][:6] - union[:6]).max())
        det["max_abs_dtheta_first6_union_vs_single_gpu"] = d
        ok = ok and d < 1e-8
        single.close()
    eng.close()
    ok = _all_ok(dist, ok)
    det["ok"] = ok
    return ok, det


def run_all(dist, rank, world, local):
    """the checks that apply to this world size; {name: details}, every value carries "ok" """
    out = {}
    if world < 2:
        return out
    _, out["rows_sharded_fp64"] = rows_check(dist, rank, world, local, tc=False)
    _, out["rows_sharded_tcgen05"] = rows_check(dist, rank, world, local, tc=True)
    _, out["rows_sharded_tcgen05_wide"] = rows_check(dist, rank, world, local, tc=True, wide=True)
    _, out["pooled_diagnostics"] = diag_check(dist, rank, world, local)
    if world >= 4 and world % 2 == 0:
        _, out["mesh_chains_x_rows"] = mesh_check(dist, rank, world, local, 2)
    return out
