"""GPU: parity at the shapes BASELINE.json names, i.e. the kernel instantiations and grids the
benchmarked jobs actually run (the other GPU tests use small models):

  configs[2]  logistic N=100000 x D=100, 4096 points: k_logistic_mma<8,1,13,32,2> over all 296 row
              chunks (1e-12 against the oracle) and k_logistic_tc<128> with the full-width row cut
              (stated bounds, incl. the energy-difference bound at N = 1e5)
  configs[3]  hierarchical regression 1000 groups x 100 obs, D = 2005: k_block_run<HierBlock<8>, 8>,
              first transitions of 8 chains against the oracle
  configs[1]  eight schools, 65 536 chains: k_chain_run_dyn (task queue), 64 sampled chain ids against
              the oracle

Oracle work is kept to seconds: it evaluates a sample of the points / chains only.
"""
import numpy as np
import pytest

import oracle as O
import turing_jl_b200 as T

from helpers import oracle_config, oracle_of

pytestmark = pytest.mark.gpu

TC = {"logistic_tc": 1}


@pytest.fixture(scope="module")
def config3():
    model = T.models.config3_logistic()
    rng = np.random.default_rng(5)
    # a cloud around a point near the posterior mode (Newton steps on the host)
    X, y = model.X, model.y
    b = np.zeros(model.D)
    for _ in range(6):
        p = 1.0 / (1.0 + np.exp(-X @ b))
        H = (X * (p * (1 - p))[:, None]).T @ X + np.eye(model.D)
        b = b + np.linalg.solve(H, X.T @ (y - p) - b)
    wide = rng.standard_normal((4096, model.D)) * 0.3           # far from the mode: |eta| up to ~10
    wide[:2048] = b + rng.standard_normal((2048, model.D)) * 0.01   # posterior-width cloud
    idx = np.concatenate([rng.choice(2048, 24, replace=False), 2048 + rng.choice(2048, 24, replace=False)])
    olp, og = oracle_of(model).lpgrad(wide[idx])
    return model, wide, idx, olp, og


def test_logistic_fp64_dmma_at_config3_shape(built, config3):
    model, th, idx, olp, og = config3
    ldf = T.LogDensityFunction(model)
    lp, g = ldf.logdensity_and_gradient(th)          # 4096 points: the bench job's grid
    np.testing.assert_allclose(lp[idx], olp, rtol=1e-12)
    gn = np.linalg.norm(og, axis=1, keepdims=True)
    assert np.max(np.abs(g[idx] - og) / gn) < 1e-12


def test_logistic_tcgen05_at_config3_shape(built, config3):
    """stated bounds of the bf16x3 mode at N = 1e5 (include/tnuts.h):
    |d lp| <= 3e-7 |lp|, |d grad|_2 <= 1e-4 |grad_lik|_2, and for what the sampler consumes --
    lp differences inside a posterior-width cloud -- |d (lp_a - lp_b)| <= 3e-3 (the systematic
    ~1e-7 shrink of eta, a smooth function of theta; energy errors of a NUTS trajectory are O(0.1))"""
    model, th, idx, olp, og = config3
    ldf = T.LogDensityFunction(model, options=TC)
    lp, g = ldf.logdensity_and_gradient(th)
    np.testing.assert_allclose(lp[idx], olp, rtol=3e-7)
    glik = og + th[idx] / model.hyper[0] ** 2
    err = np.linalg.norm(g[idx] - og, axis=1)
    assert np.all(err <= 1e-4 * np.linalg.norm(glik, axis=1)), (err / np.linalg.norm(glik, axis=1)).max()
    ref = T.LogDensityFunction(model).logdensity(th[:2048])
    d = (lp[:2048] - lp[0]) - (ref - ref[0])
    print("tcgen05 energy-difference error over the posterior cloud at N=1e5: %.2e" % np.max(np.abs(d)))
    assert np.max(np.abs(d)) < 3e-3, np.max(np.abs(d))


def test_hier_block_kernel_at_config4_shape(built):
    """k_block_run<HierBlock<8>, 8> (D = 2005): integer tree statistics and positions of the first
    transitions of 8 chains equal the oracle's"""
    model = T.models.config4_hier()
    assert model.D == 2005
    spl = T.NUTS(3, 0.8, max_depth=6)       # depth capped: the first transitions would run 1023 leapfrogs
    eng = T.Engine(model, spl, 8, 3, seed=21)
    eng.set_option("keep_theta", 1)
    eng.init()
    eng.run(5, 1, 1)
    out = eng.fetch(theta=True)
    ref = O.sample_chains(oracle_of(model), oracle_config(spl, 3, 21), 8, 5)
    st = out["stats"].transpose(2, 0, 1)
    for col in (0, 6, 7):
        np.testing.assert_array_equal(st[:, :, col], ref["stats"][:, :, col])
    np.testing.assert_allclose(out["theta"].transpose(2, 0, 1), ref["theta"], atol=1e-7)
    np.testing.assert_allclose(eng.stepsize(), ref["eps"], rtol=1e-8)
    eng.close()


def test_eight_schools_task_queue_kernel_at_config2_shape(built):
    """65 536 chains (k_chain_run_dyn: 512 chain groups on 296 persistent CTAs): 64 sampled chain ids
    against the oracle through warm-up adaptation"""
    model = T.EightSchools()
    spl = T.NUTS(30, 0.8)
    C = 65536
    eng = T.Engine(model, spl, C, 30, seed=9)
    eng.set_option("keep_theta", 1)
    eng.init()
    eng.run(48, 1, 1)
    out = eng.fetch(theta=True)
    rng = np.random.default_rng(1)
    ids = np.sort(rng.choice(C, 64, replace=False))
    ids[0], ids[-1] = 0, C - 1
    ref = O.sample_chains(oracle_of(model), oracle_config(spl, 30, 9), 0, 48, chain_ids=ids)
    st = out["stats"][:, :, ids].transpose(2, 0, 1)
    same = (st[..., 0] == ref["stats"][..., 0]) & (st[..., 6] == ref["stats"][..., 6])
    assert np.all(same[:, :8])
    assert same.mean() > 0.95, same.mean()          # chaotic dynamics: one ulp grows ~5x per transition
    np.testing.assert_allclose(out["theta"][:, :, ids].transpose(2, 0, 1)[:, :8], ref["theta"][:, :8], atol=1e-7)
    eng.close()


def test_state_blob_is_tied_to_sampler_metric_seed_and_shard(built):
    """tnuts_set_state rejects blobs of another sampler / metric / seed / chain shard (the arrays mean
    different things, and the random streams are addressed by seed and global chain id)"""
    m = T.EightSchools()
    a = T.Engine(m, T.NUTS(10, 0.8), 16, 10, seed=3)
    a.init()
    a.run(12, 1, 1)
    blob = a.get_state()
    a.close()
    ok = T.Engine(m, T.NUTS(10, 0.8), 16, 0, seed=3)
    ok.set_state(blob)                      # n_adapts may differ: a resumed run does not adapt
    ok.close()
    for kw, spl in ((dict(seed=4), T.NUTS(10, 0.8)), (dict(seed=3, chain_offset=16), T.NUTS(10, 0.8)),
                    (dict(seed=3), T.HMCDA(10, 0.8, 0.3)), (dict(seed=3), T.NUTS(10, 0.8, metricT="unit"))):
        e = T.Engine(m, spl, 16, 10, **kw)
        with pytest.raises(T._lib.TnutsError):
            e.set_state(blob)
        e.close()


def test_block_steps_from_a_state_blob_are_reproducible(built):
    """bench.py's step unit: restore the blob, run a block; the in-run kernel timer reports the
    gradient kernel's launches and the persistent tail kernel"""
    m = T.models.synthetic_logistic(3000, 20, seed=4)
    e = T.Engine(m, T.NUTS(40, 0.8), 700, 40, seed=6)
    e.set_option("logistic_tc", 1)
    e.set_option("time_kernels", 1)
    e.init()
    e.run(60, 1, 1)
    blob = e.get_state()
    draws = []
    for _ in range(2):
        e.set_state(blob)
        g0 = e.grad_evals
        e.run(10, 1, 1)
        draws.append(e.fetch_draws())
        sec, launches, sampled, tiles, psec, ppumps = e.kernel_times()
        assert launches > 10 and 0 < sampled <= launches and tiles >= launches      # 700 chains: host-pumped first
        assert ppumps > 0 and psec > 0                                              # <= 512 left: persistent kernel
        assert 0 < sec + psec < e.L.tnuts_device_seconds(e.h)
        assert e.grad_evals - g0 >= 10 * 700
    assert np.array_equal(draws[0], draws[1])
    e.close()
