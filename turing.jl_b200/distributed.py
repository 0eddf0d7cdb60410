"""Multi-GPU host logic: one process per GPU (torchrun), `torch.distributed` only as plumbing.

Two ways the path shards:
  * chains: rank g owns chains [lo, hi); the model is replicated; NO collective on the data
    path.  Random draws are addressed by the GLOBAL chain id, so the union of the shards is
    bit-identical to a single-GPU run of all chains.
  * rows:   every rank owns all chains and a slice of the data rows; libtnuts sums the
    per-leapfrog [D+1][chains] likelihood/gradient block with ncclAllReduce on its own stream
    (tnuts_comm_init); the host only broadcasts the 128-byte ncclUniqueId.
  * hybrid: a chains x rows mesh (SURVEY.md 8e "Hybrid"): `world = chain_groups * row_shards`;
    rank r is row shard r % row_shards of chain group r // row_shards.  Every chain group owns a
    slice of the chains and a private NCCL communicator over its row shards: the all-reduce stays
    inside the group, nothing crosses groups on the data path.
"""
import ctypes as C

import numpy as np

from . import _lib
from .engine import Engine


def chain_bounds(n_chains, world):
    """[lo_0, lo_1, ..., n_chains]: contiguous, sizes differ by at most one"""
    return [n_chains * i // world for i in range(world + 1)]


def row_bounds(n_rows, world):
    return [n_rows * i // world for i in range(world + 1)]


def shard_rows(model, world, rank):
    """a copy of a row-separable model (logistic regression) holding only this rank's rows"""
    from .models import LogisticRegression
    if not isinstance(model, LogisticRegression):
        raise TypeError("row sharding is defined for LogisticRegression")
    b = row_bounds(model.N, world)
    return LogisticRegression(model.X[b[rank]:b[rank + 1]], model.y[b[rank]:b[rank + 1]],
                              model.hyper[0])


def _new_unique_id():
    buf = np.zeros(128, dtype=np.uint8)
    rc = _lib.load().tnuts_comm_unique_id(buf.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise _lib.TnutsError(rc, "tnuts_comm_unique_id failed")
    return buf


def broadcast_unique_id(dist, rank, device=None, src=0, group=None, make_id=_new_unique_id):
    """global rank `src` creates the ncclUniqueId through the C ABI; everybody in `group` (default:
    the whole process group) receives the 128 bytes"""
    import torch
    buf = make_id() if rank == src else np.zeros(128, dtype=np.uint8)
    t = torch.from_numpy(np.ascontiguousarray(buf, dtype=np.uint8))
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=src, group=group)
    return np.ascontiguousarray(t.cpu().numpy())


def mesh_coords(rank, world, row_shards):
    """(chain_group, row_rank, n_chain_groups) of `rank` in a chains x rows mesh"""
    if row_shards < 1 or world % row_shards != 0:
        raise ValueError("world size %d is not a multiple of row_shards %d" % (world, row_shards))
    return rank // row_shards, rank % row_shards, world // row_shards


def mesh_row_group(dist, rank, world, row_shards):
    """the torch process group of this rank's row shards (every rank creates every group, as
    torch.distributed requires) and the global rank of its leader"""
    cg, _, ngroups = mesh_coords(rank, world, row_shards)
    mine = None
    for g in range(ngroups):
        ranks = list(range(g * row_shards, (g + 1) * row_shards))
        grp = dist.new_group(ranks)
        if g == cg:
            mine = grp
    return mine, cg * row_shards


def attach_comm(engine, dist, device=None):
    """give `engine` an NCCL communicator spanning the process group"""
    world, rank = dist.get_world_size(), dist.get_rank()
    uid = broadcast_unique_id(dist, rank, device)
    engine.comm_init(uid.ctypes.data_as(C.c_void_p), world, rank)
    engine._uid = uid
    return engine


def make_chain_sharded_engine(model, spl, n_chains_total, n_adapts, seed, rank, world, device):
    b = chain_bounds(n_chains_total, world)
    return Engine(model, spl, b[rank + 1] - b[rank], n_adapts, seed, device=device,
                  chain_offset=b[rank])


def make_row_sharded_engine(model, spl, n_chains, n_adapts, seed, rank, world, device, dist):
    local = shard_rows(model, world, rank)
    e = Engine(local, spl, n_chains, n_adapts, seed, device=device, chain_offset=0,
               row_shards=world, row_rank=rank, strategy=2)
    e._local_model = local
    if world > 1:
        attach_comm(e, dist, device="cuda:%d" % device if device is not None else None)
    return e


def make_mesh_engine(model, spl, n_chains_total, n_adapts, seed, rank, world, device, dist, row_shards,
                     make_id=_new_unique_id):
    """chains x rows mesh: chain group g = rank // row_shards owns chains [b[g], b[g+1]) and all the
    rows, cut over its `row_shards` ranks; the per-leapfrog all-reduce runs on a communicator
    private to the group"""
    cg, rr, ngroups = mesh_coords(rank, world, row_shards)
    b = chain_bounds(n_chains_total, ngroups)
    local = shard_rows(model, row_shards, rr) if row_shards > 1 else model
    e = Engine(local, spl, b[cg + 1] - b[cg], n_adapts, seed, device=device, chain_offset=b[cg],
               row_shards=row_shards, row_rank=rr, strategy=2)
    e._local_model = local
    e.mesh = (cg, rr, ngroups)
    if row_shards > 1:
        grp, leader = mesh_row_group(dist, rank, world, row_shards)
        uid = broadcast_unique_id(dist, rank, "cuda:%d" % device if device is not None else None,
                                  src=leader, group=grp, make_id=make_id)
        e.comm_init(uid.ctypes.data_as(C.c_void_p), row_shards, rr)
        e._uid = uid
    return e


def combine_diagnostics(dist, rhat_local, ess_local, device=None):
    """chains sharded over ranks are independent: bulk ESS adds over shards; R-hat is reported
    as the worst shard (each shard is itself a many-chain ensemble)"""
    import torch
    r = torch.as_tensor(np.asarray(rhat_local, dtype=np.float64))
    e = torch.as_tensor(np.asarray(ess_local, dtype=np.float64))
    if device is not None:
        r, e = r.to(device), e.to(device)
    dist.all_reduce(r, op=dist.ReduceOp.MAX)
    dist.all_reduce(e, op=dist.ReduceOp.SUM)
    return r.cpu().numpy(), e.cpu().numpy()


def gather_param_moments(dist, draws, device=None):
    """pooled posterior mean / variance per parameter from the shards' draws
    [iters, params, chains_local] via a sum of sufficient statistics"""
    import torch
    x = np.asarray(draws, dtype=np.float64)
    n = float(x.shape[0] * x.shape[2])
    s = np.concatenate([[n], x.sum(axis=(0, 2)), (x * x).sum(axis=(0, 2))])
    t = torch.as_tensor(s)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    t = t.cpu().numpy()
    P = x.shape[1]
    nn, s1, s2 = t[0], t[1:1 + P], t[1 + P:]
    mean = s1 / nn
    var = (s2 - nn * mean * mean) / (nn - 1.0)
    return mean, var
