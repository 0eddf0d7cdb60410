"""GPU: multi-GPU modes.  Chain sharding needs no collective, so its invariant -- the union of
the shards is bit-identical to one engine running all chains -- is checked on ONE GPU with two
engines; the NCCL row-sharded path needs >= 2 GPUs and is skipped otherwise."""
import os
import subprocess
import sys

import numpy as np
import pytest

import turing_jl_b200 as T
from turing_jl_b200 import distributed as D

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("strategy,model_name", [(1, "eight"), (2, "logistic")])
def test_chain_shards_union_equals_single_engine(built, strategy, model_name):
    model = T.EightSchools() if model_name == "eight" else T.models.synthetic_logistic(200, 6, seed=2)
    spl = T.NUTS(20, 0.8)
    nch = 40 if strategy == 1 else 150     # lock-step: more than one 64-chain tile per engine
    full = T.Engine(model, spl, nch, 20, seed=3, strategy=strategy)
    full.init()
    full.run(30, 1, 1)
    a = full.fetch_draws()
    full.close()
    parts = []
    b = D.chain_bounds(nch, 3)
    for r in range(3):
        e = D.make_chain_sharded_engine(model, spl, nch, 20, 3, r, 3, device=0)
        if strategy == 2:
            e.close()
            e = T.Engine(model, spl, b[r + 1] - b[r], 20, 3, device=0, chain_offset=b[r], strategy=2)
        e.init()
        e.run(30, 1, 1)
        parts.append(e.fetch_draws())
        e.close()
    assert np.array_equal(a, np.concatenate(parts, axis=2))


def test_row_sharded_logistic_two_gpus(built):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29611",
         os.path.join(ROOT, "scripts", "multi_gpu_check.py"), "--rows"],
        capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "ROWS-OK" in out.stdout


def test_row_sharded_logistic_tc_wide_two_gpus(built):
    """rows sharded, D = 160: the CTA-pair tcgen05 kernel under the per-leapfrog all-reduce"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29615",
         os.path.join(ROOT, "scripts", "multi_gpu_check.py"), "--rows", "--tc", "--wide"],
        capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "ROWS-OK" in out.stdout


def test_row_sharded_logistic_tc_two_gpus(built):
    """the tcgen05 likelihood on row shards: identical chains on every rank, close to single GPU"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29617",
         os.path.join(ROOT, "scripts", "multi_gpu_check.py"), "--rows", "--tc"],
        capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "ROWS-OK" in out.stdout


def test_pooled_diagnostics_two_gpus(built):
    """chains sharded over 2 GPUs: the NCCL parameter-parallel diagnostics equal the single-engine
    R-hat / bulk ESS of the same chains"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29614",
         os.path.join(ROOT, "scripts", "multi_gpu_check.py"), "--diag"],
        capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DIAG-OK" in out.stdout


def test_chain_by_row_mesh_four_gpus(built):
    """SURVEY 8e "Hybrid": 2 chain groups x 2 row shards, NCCL all-reduce inside each group only"""
    import torch
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4",
         "--master-addr", "127.0.0.1", "--master-port", "29619",
         os.path.join(ROOT, "scripts", "multi_gpu_check.py"), "--mesh"],
        capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MESH-OK" in out.stdout


def test_streamed_slab_fills_a_wider_chain_array(built):
    """tnuts_run_streamed_strided: two engines (same GPU) stream their draws straight into disjoint
    chain slabs of ONE (iterations x columns x chains) array; equals one engine running all chains"""
    model = T.EightSchools()
    spl = T.NUTS(20, 0.8)
    nch, n_tr = 96, 50
    full = T.Engine(model, spl, nch, 20, seed=5)
    full.init()
    full.run(n_tr, 11, 2)
    a = full.fetch_draws()
    full.close()
    value = T.pinned.empty((a.shape[0], a.shape[1], nch))
    value[:] = np.nan
    b = [0, 40, 96]
    for r in range(2):
        e = T.Engine(model, spl, b[r + 1] - b[r], 20, seed=5, chain_offset=b[r])
        e.init()
        assert e.run_streamed_slab(n_tr, 11, 2, value, b[r], chunk_draws=7) == a.shape[0]
        e.close()
    assert np.array_equal(a, value)


def test_sample_over_two_devices_in_one_process(built):
    """sample(..., devices=[0, 1]): each GPU streams into its slab, diagnostics pooled over NCCL;
    same chains and same R-hat / ESS as the single-GPU call"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    model = T.EightSchools()
    one = T.sample(model, T.NUTS(0.8), T.MCMCThreads(), 200, 512, seed=9, diagnostics=True, verbose=False)
    two = T.sample(model, T.NUTS(0.8), T.MCMCThreads(), 200, 512, seed=9, diagnostics=True, verbose=False,
                   devices=[0, 1])
    assert np.array_equal(one.value, two.value)
    assert np.allclose(one.info["rhat"], two.info["rhat"], rtol=1e-9)
    assert np.allclose(one.info["ess_bulk"], two.info["ess_bulk"], rtol=1e-9)
