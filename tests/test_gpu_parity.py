"""GPU: the CUDA engine, called through the C ABI, against the oracle on the same seeded inputs.

Tolerances (fp64 everywhere; the reference's HMC is Float64-only, test/floattypes/main.jl:54-57):
  lp / gradient ............ 1e-12 relative (north star)
  leapfrog trajectories .... 1e-10 absolute over 64 steps (libm + FMA contraction differences)
  NUTS chains .............. integer tree statistics identical and |dtheta| < 1e-8 over the first
                             transitions (the dynamics is chaotic: an ulp grows ~5x per transition)
"""
import numpy as np
import pytest

import oracle as O
import turing_jl_b200 as T

from helpers import golden_cases, model_from_golden, oracle_config, oracle_of, small_models

pytestmark = pytest.mark.gpu


def _strategies(model):
    """every execution strategy that supports the model must agree with the oracle"""
    out = [0]
    return out


@pytest.fixture(scope="module")
def golden(built):
    return golden_cases()


def test_lpgrad_golden_and_oracle(golden):
    rng = np.random.default_rng(1)
    for case in golden["cases"]:
        model = model_from_golden(case)
        try:
            ldf = T.LogDensityFunction(model)
        except T._lib.TnutsError as ex:
            if ex.code == T._lib.E_UNSUPPORTED:
                pytest.fail("no CUDA path for family %s: %s" % (case["name"], ex))
            raise
        assert ldf.dimension() == model.D
        th = np.array([p["theta"] for p in case["points"]])
        lp, g = ldf.logdensity_and_gradient(th)
        np.testing.assert_allclose(lp, [p["lp"] for p in case["points"]], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(g, [p["grad"] for p in case["points"]], rtol=1e-11, atol=1e-11)
        th = rng.standard_normal((257, model.D)) * 0.8
        lp, g = ldf.logdensity_and_gradient(th)
        olp, og = oracle_of(model).lpgrad(th)
        np.testing.assert_allclose(lp, olp, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(g, og, rtol=1e-11, atol=1e-10)
        np.testing.assert_allclose(ldf.logdensity(th), lp, rtol=0, atol=0)
        # single point form
        l1, g1 = ldf.logdensity_and_gradient(th[0])
        assert l1 == lp[0] and np.array_equal(g1, g[0])


def test_constrain_and_user_space_logp(golden):
    rng = np.random.default_rng(2)
    for case in golden["cases"]:
        model = model_from_golden(case)
        eng = T.Engine(model, T.NUTS(), 1, 0, 0)
        th = rng.standard_normal((33, model.D)) * 0.7
        p, lp3 = eng.constrain(th)
        om = oracle_of(model)
        for i in range(33):
            op, olp3 = om.constrain(th[i])
            np.testing.assert_allclose(p[i], op, rtol=1e-14)
            np.testing.assert_allclose(lp3[i], olp3, rtol=1e-12, atol=1e-12)
        eng.close()


def test_leapfrog_trajectories(golden):
    rng = np.random.default_rng(3)
    for case in golden["cases"]:
        model = model_from_golden(case)
        om = oracle_of(model)
        eng = T.Engine(model, T.NUTS(), 1, 0, 0)
        n, L = 16, 64
        th = rng.standard_normal((n, model.D)) * 0.5
        r = rng.standard_normal((n, model.D))
        minv = np.exp(rng.standard_normal(model.D) * 0.3)
        tt, tr, tH = eng.leapfrog(th, r, minv, 0.01, L)
        for i in range(n):
            ot, orr, oH = O.leapfrog_trajectory(om, th[i], r[i], minv, 0.01, L)
            np.testing.assert_allclose(tt[i], ot, atol=1e-10, rtol=1e-10)
            np.testing.assert_allclose(tr[i], orr, atol=1e-9, rtol=1e-10)
            np.testing.assert_allclose(tH[i], oH, atol=1e-8, rtol=1e-12)
        eng.close()


CHAIN_CASES = ([(n, 1) for n in ("gdemo", "linreg", "eight_schools")]
               + [(n, 2) for n in ("gdemo", "linreg", "eight_schools", "linreg_n40", "eight_schools_j5",
                                   "std_normal_d7", "logistic_small", "logistic_d130", "hier_small")]
               + [(n, 3) for n in ("hier_small", "hier_g150", "std_normal_d7", "std_normal_d600")])


@pytest.mark.parametrize("name,strategy", CHAIN_CASES)
@pytest.mark.parametrize("n_adapts", [0, 30])
def test_nuts_chains_match_oracle(built, name, strategy, n_adapts):
    """strategy 1 = one thread per chain, 2 = lock-step pump, 3 = one CTA per chain; every
    strategy must reproduce the oracle"""
    model = small_models()[name]
    spl = T.NUTS(n_adapts, 0.8)
    C, n_tr = 96, n_adapts + 12
    eng = T.Engine(model, spl, C, n_adapts, seed=5, strategy=strategy)
    eng.set_option("keep_theta", 1)
    eng.init()
    eng.run(n_tr, 1, 1)
    out = eng.fetch(theta=True)
    ref = O.sample_chains(oracle_of(model), oracle_config(spl, n_adapts, 5), C, n_tr)
    st = out["stats"].transpose(2, 0, 1)
    th = out["theta"].transpose(2, 0, 1)
    # first 8 transitions: exact tree statistics, tight positions
    for col in (0, 6, 7):   # n_steps, tree_depth, numerical_error
        np.testing.assert_array_equal(st[:, :8, col], ref["stats"][:, :8, col])
    np.testing.assert_allclose(th[:, :8], ref["theta"][:, :8], atol=1e-7, rtol=1e-7)
    np.testing.assert_allclose(st[:, :8, 1], ref["stats"][:, :8, 1], atol=1e-7)   # acceptance_rate
    np.testing.assert_allclose(st[:, :8, 8], ref["stats"][:, :8, 8], rtol=1e-9)   # step_size
    # whole run: the overwhelming majority of transitions still has identical trees
    same = (st[..., 0] == ref["stats"][..., 0]) & (st[..., 6] == ref["stats"][..., 6])
    assert same.mean() > 0.9
    # constrained params / log-probs of kept draws are consistent with theta
    p, lp3 = eng.constrain(th[:, -1])
    np.testing.assert_allclose(out["params"].transpose(2, 0, 1)[:, -1], p, rtol=1e-13)
    np.testing.assert_allclose(out["logp"].transpose(2, 0, 1)[:, -1], lp3, rtol=1e-11, atol=1e-10)
    eng.close()


@pytest.mark.parametrize("name,strategy", [("eight_schools", 1), ("eight_schools", 2),
                                           ("logistic_small", 2), ("hier_small", 2)])
def test_init_matches_oracle(built, name, strategy):
    model = small_models()[name]
    spl = T.NUTS(0.8)
    eng = T.Engine(model, spl, 512, 0, seed=77, strategy=strategy)
    eng.init()
    ref = O.sample_chains(oracle_of(model), oracle_config(spl, 0, 77), 512, 0)
    np.testing.assert_allclose(eng.stepsize(), ref["eps"], rtol=1e-9)    # find_good_stepsize
    th = eng.theta()
    assert np.all(np.abs(th) <= 2.0)                                     # InitFromUniform
    assert eng.init_tries_max >= 1
    eng.close()


def test_strategies_agree_bitwise_on_tree_stats(built):
    """the two execution strategies are the same algorithm: identical n_steps / depth for a
    whole short run, positions equal to rounding"""
    model = T.EightSchools()
    spl = T.NUTS(40, 0.8)
    outs = []
    for strat in (1, 2):
        eng = T.Engine(model, spl, 64, 40, seed=21, strategy=strat)
        eng.set_option("keep_theta", 1)
        eng.init()
        eng.run(60, 1, 1)
        outs.append(eng.fetch(theta=True))
        eng.close()
    same = outs[0]["stats"][:, [0, 6, 7], :] == outs[1]["stats"][:, [0, 6, 7], :]
    assert same.mean() > 0.98
    np.testing.assert_allclose(outs[0]["theta"][:10], outs[1]["theta"][:10], atol=1e-8)


@pytest.mark.parametrize("name,strategy", [("gdemo", 1), ("gdemo", 2), ("logistic_small", 2),
                                           ("hier_small", 2), ("hier_small", 3), ("hier_g150", 3)])
def test_hmc_and_hmcda_match_oracle(built, name, strategy):
    """HMC (fixed L) and HMCDA (fixed integration time) on every execution strategy
    (src/mcmc/hmc.jl:59-80, 285-327, 425-434)"""
    model = small_models()[name]
    for spl, nad in ((T.HMC(0.02, 7), 0), (T.HMCDA(40, 0.65, 0.15), 40)):
        C, n_tr = 64, nad + 10
        eng = T.Engine(model, spl, C, nad, seed=9, strategy=strategy)
        eng.set_option("keep_theta", 1)
        eng.init()
        eng.run(n_tr, 1, 1)
        out = eng.fetch(theta=True)
        ref = O.sample_chains(oracle_of(model), oracle_config(spl, nad, 9), C, n_tr)
        th = out["theta"].transpose(2, 0, 1)
        st = out["stats"].transpose(2, 0, 1)
        np.testing.assert_allclose(th[:, :8], ref["theta"][:, :8], atol=1e-7)
        np.testing.assert_array_equal(st[:, :8, 0], ref["stats"][:, :8, 0])      # n_steps
        np.testing.assert_array_equal(st[:, :8, 6], ref["stats"][:, :8, 6])      # is_accept
        np.testing.assert_allclose(st[:, :8, 1], ref["stats"][:, :8, 1], atol=1e-7)
        eng.close()


def test_init_failure_message(built):
    """test/mcmc/hmc.jl:284-296: the error names the initial-parameters URL"""
    model = T.GDemo(float("nan"), 2.0)   # lp is NaN everywhere -> no finite initial point
    with pytest.raises(RuntimeError) as ei:
        T.sample(model, T.NUTS(0.8), T.MCMCThreads(), 10, 4, verbose=False)
    assert "failed to find valid initial parameters in 1000 tries" in str(ei.value)
    assert "https://turinglang.org/docs/uri/initial-parameters" in str(ei.value)


@pytest.mark.parametrize("name,strategy", [("eight_schools", 1), ("eight_schools", 2), ("logistic_small", 2),
                                           ("hier_small", 3)])
def test_unit_metric_matches_oracle(built, name, strategy):
    """metricT = UnitEuclideanMetric (the HMC / HMCDA default, selectable for NUTS): windows and
    dual-averaging restarts as with the diagonal metric, identity mass matrix, on every strategy"""
    model = small_models()[name]
    for spl in (T.NUTS(120, 0.8, metricT="unit"), T.HMCDA(120, 0.65, 0.1)):
        assert spl.metric == 1
        C, n_tr = 64, 130
        eng = T.Engine(model, spl, C, 120, seed=31, strategy=strategy)
        eng.set_option("keep_theta", 1)
        eng.init()
        eng.run(n_tr, 1, 1)
        out = eng.fetch(theta=True)
        assert np.all(eng.inv_metric() == 1.0)
        ref = O.sample_chains(oracle_of(model), oracle_config(spl, 120, 31), C, n_tr)
        st = out["stats"].transpose(2, 0, 1)
        np.testing.assert_array_equal(st[:, :8, 0], ref["stats"][:, :8, 0])          # n_steps
        np.testing.assert_allclose(out["theta"].transpose(2, 0, 1)[:, :8], ref["theta"][:, :8], atol=1e-7)
        # the adapted step sizes agree after all windows (loose: the chains decorrelate by rounding, and
        # the median of 64 chains whose step sizes spread by ~20 % has a standard error of ~3 %)
        np.testing.assert_allclose(np.median(eng.stepsize()), np.median(ref["eps"]), rtol=0.12)
        eng.close()


@pytest.mark.parametrize("name,strategy", [("gdemo", 1), ("gdemo", 2), ("hier_small", 3)])
def test_hmc_with_zero_step_size_searches_one(built, name, strategy):
    """src/mcmc/hmc.jl:189-194: `iszero(spl.eps)` runs find_good_stepsize for HMC as well"""
    model = small_models()[name]
    spl = T.HMC(0.0, 4)
    eng = T.Engine(model, spl, 48, 0, seed=6, strategy=strategy)
    eng.init()
    ref = O.sample_chains(oracle_of(model), oracle_config(spl, 0, 6), 48, 0)
    np.testing.assert_allclose(eng.stepsize(), ref["eps"], rtol=1e-9)
    assert len(np.unique(eng.stepsize())) > 1        # searched per chain, not a constant
    eng.close()


def test_dense_metric_matches_oracle(built):
    """metricT = DenseEuclideanMetric on the thread-per-chain strategy.  (1) every family: identical
    to the diagonal metric until the first window ends, and to the oracle's dense run over the first
    transitions; (2) standard normal (linear, non-chaotic dynamics): the whole run INCLUDING the
    covariance refresh and its Cholesky factor follows the oracle; (3) gdemo known answers."""
    for name in ("gdemo", "linreg", "eight_schools"):
        model = small_models()[name]
        spl = T.NUTS(120, 0.8, metricT="dense")
        assert spl.metric == 2
        C, n_tr = 64, 110
        eng = T.Engine(model, spl, C, 120, seed=41)
        eng.set_option("keep_theta", 1)
        eng.init()
        eng.run(n_tr, 1, 1)
        out = eng.fetch(theta=True)
        ref = O.sample_chains(oracle_of(model), oracle_config(spl, 120, 41), C, n_tr)
        np.testing.assert_array_equal(out["stats"].transpose(2, 0, 1)[:, :8, 0], ref["stats"][:, :8, 0])
        np.testing.assert_allclose(out["theta"].transpose(2, 0, 1)[:, :8], ref["theta"][:, :8], atol=1e-7)
        dm = eng.dense_inv_metric()
        assert dm.shape == (C, model.D, model.D)
        np.testing.assert_allclose(dm, dm.transpose(0, 2, 1), rtol=1e-12, atol=1e-14)      # symmetric
        assert np.all(np.linalg.eigvalsh(dm) > 0)                                         # positive definite
        np.testing.assert_allclose(np.diagonal(dm, axis1=1, axis2=2), eng.inv_metric(), rtol=0, atol=0)
        eng.close()
        diag = T.Engine(model, T.NUTS(120, 0.8), C, 120, seed=41, strategy=1)
        diag.set_option("keep_theta", 1)
        diag.init()
        diag.run(99, 1, 1)                   # the first window ends at iteration 100
        np.testing.assert_array_equal(diag.fetch(theta=True)["theta"], out["theta"][:99])
        diag.close()
    # (2) linear dynamics: compare through the metric refresh (windows 7..36 of 40 warm-up iterations)
    model = T.StdNormal(4)
    spl = T.NUTS(40, 0.8, metricT="dense")
    eng = T.Engine(model, spl, 32, 40, seed=7)
    eng.set_option("keep_theta", 1)
    eng.init()
    eng.run(60, 1, 1)
    out = eng.fetch(theta=True)
    ref = O.sample_chains(oracle_of(model), oracle_config(spl, 40, 7), 32, 60)
    th, rth = out["theta"].transpose(2, 0, 1), ref["theta"]
    err = np.abs(th - rth).max(axis=(1, 2))
    same = err < 1e-6
    # a tree decision that falls within rounding of its threshold flips a chain for good
    assert same.mean() >= 0.8, (same.mean(), np.sort(err)[-8:])
    np.testing.assert_allclose(eng.stepsize()[same], ref["eps"][same], rtol=1e-8)
    np.testing.assert_allclose(eng.inv_metric()[same], ref["minv"][same], rtol=1e-8)
    assert np.any(np.abs(eng.dense_inv_metric()[:, 0, 1]) > 1e-6)       # off-diagonals were estimated
    eng.close()
    # (3) test/mcmc/hmc.jl:138-142 with the dense metric; HMCDA runs through the same pieces
    for spl in (T.NUTS(500, 0.8, metricT="dense"), T.HMCDA(500, 0.65, 0.3, metricT="dense")):
        ch = T.sample(T.gdemo_default(), spl, T.MCMCThreads(), 1000, 512, seed=3, verbose=False,
                      diagnostics=True)
        assert abs(ch["s"].mean() - 49 / 24) < 0.1 and abs(ch["m"].mean() - 7 / 6) < 0.05
        assert np.max(ch.info["rhat"]) < 1.01


def test_dense_metric_on_the_pump_strategy_matches_oracle(built):
    """metricT = DenseEuclideanMetric on the pump strategy (dense_pump.cuh): the state machine runs the
    unit-metric dynamics in whitened coordinates phi = U^-T theta, which is the dense-metric dynamics of
    the oracle up to rounding.  (1) every family on strategy 2, incl. logistic regression with D > 32
    (several dimensions per lane): the oracle's dense run over the first transitions, and the
    diagonal-metric pump until the first window ends; (2) standard normal: through the covariance
    refresh, its Cholesky factor and the re-expression of phi; (3) a metric is estimated, symmetric,
    positive definite, and its diagonal is what tnuts_get_inv_metric reports; (4) resume is bit-exact;
    (5) the reference's gdemo answers with NUTS-dense and HMCDA-dense on this strategy."""
    models = dict(small_models())
    models["logistic_d40"] = T.models.synthetic_logistic(400, 40, seed=21)
    for name in ("gdemo", "eight_schools", "logistic_small", "logistic_d40", "hier_small"):
        model = models[name]
        spl = T.NUTS(120, 0.8, metricT="dense")
        C, n_tr = 24, 110
        eng = T.Engine(model, spl, C, 120, seed=41, strategy=2)
        eng.set_option("keep_theta", 1)
        eng.init()
        eng.run(n_tr, 1, 1)
        out = eng.fetch(theta=True)
        ref = O.sample_chains(oracle_of(model), oracle_config(spl, 120, 41), C, n_tr)
        np.testing.assert_array_equal(out["stats"].transpose(2, 0, 1)[:, :8, 0], ref["stats"][:, :8, 0], err_msg=name)
        np.testing.assert_allclose(out["theta"].transpose(2, 0, 1)[:, :8], ref["theta"][:, :8], atol=1e-7, err_msg=name)
        dm = eng.dense_inv_metric()
        assert dm.shape == (C, model.D, model.D)
        np.testing.assert_allclose(dm, dm.transpose(0, 2, 1), rtol=1e-9, atol=1e-12)
        assert np.all(np.linalg.eigvalsh(dm) > 0)
        assert np.any(np.abs(dm[:, 0, 1]) > 1e-9), name                    # the first window (iteration 100) has ended
        np.testing.assert_allclose(np.diagonal(dm, axis1=1, axis2=2), eng.inv_metric(), rtol=0, atol=0)
        # chains that still follow the oracle after the refresh estimated the same metric
        th, rth = out["theta"].transpose(2, 0, 1), ref["theta"]
        same = np.abs(th - rth).max(axis=(1, 2)) < 1e-5
        if same.any():
            np.testing.assert_allclose(eng.inv_metric()[same], ref["minv"][same], rtol=1e-6, err_msg=name)
        eng.close()
        diag = T.Engine(model, T.NUTS(120, 0.8), C, 120, seed=41, strategy=2)
        diag.set_option("keep_theta", 1)
        diag.init()
        diag.run(99, 1, 1)                   # identity metric until the first window ends at iteration 100
        dth = diag.fetch(theta=True)["theta"]
        np.testing.assert_allclose(dth[:8], out["theta"][:8], atol=1e-9, err_msg=name)
        diag.close()
    # (2) linear dynamics: compare through the metric refresh (windows 7..36 of 40 warm-up iterations)
    model = T.StdNormal(7)
    spl = T.NUTS(40, 0.8, metricT="dense")
    eng = T.Engine(model, spl, 32, 40, seed=7, strategy=2)
    eng.set_option("keep_theta", 1)
    eng.init()
    eng.run(60, 1, 1)
    out = eng.fetch(theta=True)
    ref = O.sample_chains(oracle_of(model), oracle_config(spl, 40, 7), 32, 60)
    th, rth = out["theta"].transpose(2, 0, 1), ref["theta"]
    err = np.abs(th - rth).max(axis=(1, 2))
    same = err < 1e-6
    assert same.mean() >= 0.8, (same.mean(), np.sort(err)[-8:])
    # (the whitened arithmetic rounds differently from the oracle's M^-1 r products: 1e-6, not 1e-8)
    np.testing.assert_allclose(eng.stepsize()[same], ref["eps"][same], rtol=1e-6)
    np.testing.assert_allclose(eng.inv_metric()[same], ref["minv"][same], rtol=1e-6)
    # (4) resume: the blob carries phi, grad_phi, M^-1, its factor and the covariance accumulator
    blob = None
    a = T.Engine(model, spl, 32, 40, seed=7, strategy=2)
    a.init()
    a.run(25, 1, 1)                          # stops inside the second window
    blob = a.get_state()
    a.run(35, 1, 1)
    tail = a.fetch()["params"]
    a.close()
    b = T.Engine(model, spl, 32, 40, seed=7, strategy=2)
    b.set_state(blob)
    b.run(35, 1, 1)
    np.testing.assert_array_equal(b.fetch()["params"], tail)
    b.close()
    eng.close()
    # (5) test/mcmc/hmc.jl:138-142 with the dense metric on the pump strategy
    for spl in (T.NUTS(500, 0.8, metricT="dense"), T.HMCDA(500, 0.65, 0.3, metricT="dense")):
        ch = T.sample(T.gdemo_default(), spl, T.MCMCThreads(), 1000, 256, seed=3, verbose=False,
                      diagnostics=True, strategy=2)
        assert abs(ch["s"].mean() - 49 / 24) < 0.1 and abs(ch["m"].mean() - 7 / 6) < 0.05
        assert np.max(ch.info["rhat"]) < 1.01


def test_dense_metric_unsupported_combinations(built):
    with pytest.raises(T._lib.TnutsError):      # the CTA-per-chain strategy has no dense form
        T.Engine(small_models()["hier_small"], T.NUTS(10, 0.8, metricT="dense"), 8, 10, seed=1, strategy=3)
    with pytest.raises(T._lib.TnutsError):      # D > 256 on the pump
        T.Engine(small_models()["std_normal_d600"], T.NUTS(10, 0.8, metricT="dense"), 8, 10, seed=1)
