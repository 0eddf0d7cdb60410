"""GPU: the public `sample` API -- posterior agreement (4 MCSE, R-hat < 1.01), plumbing the
reference's tests pin (sizes, determinism, initial_params, discard_adapt, resume), diagnostics."""
import warnings

import numpy as np
import pytest

import oracle as O
from oracle import diagnostics as dg
import turing_jl_b200 as T

from helpers import oracle_config, oracle_of

pytestmark = pytest.mark.gpu


def test_demo_models_posterior_means(built):
    """DynamicPPL.TestUtils.DEMO_MODELS as test/mcmc/external_sampler.jl:198-230 samples them (NUTS,
    n_adapts = 1000, discard_initial = 1000, 2000 draws, rtol 0.2): the vector-valued demo models are two
    independent one-observation gdemo blocks, E[s] = [19/8, 8/3], E[m] = [3/4, 1]; on the
    thread-per-chain and the pump strategy"""
    for strategy in (1, 2):
        for x, es, em in ((1.5, 19 / 8, 3 / 4), (2.0, 8 / 3, 1.0)):
            ch = T.sample(T.GDemo([x], []), T.NUTS(1000, 0.8), T.MCMCThreads(), 2000, 128, seed=4,
                          diagnostics=True, verbose=False, strategy=strategy)
            np.testing.assert_allclose([ch.mean("s"), ch.mean("m")], [es, em], rtol=0.05)
            assert ch.rhat().max() < 1.01


def test_gdemo_posterior_known_answer(built):
    """E[s] = 49/24, E[m] = 7/6 (test/mcmc/hmc.jl:138-142, atol 0.2); we also require 4 MCSE"""
    ch = T.sample(T.gdemo_default(), T.NUTS(1000, 0.8), T.MCMCThreads(), 2000, 256, seed=1,
                  diagnostics=True, verbose=False)
    assert ch.size() == (2000, 14, 256)
    x = ch.get_params().transpose(2, 0, 1)[:64]
    s = dg.summarize(x)
    z = (np.array([ch.mean("s"), ch.mean("m")]) - [49 / 24, 7 / 6])
    assert np.abs(z).max() < 0.2
    assert np.abs(z / (s["mcse"] / 2)).max() < 4.0     # 256 chains: MCSE halves vs 64
    assert ch.rhat().max() < 1.01


def test_eight_schools_agrees_with_oracle_posterior(built):
    model = T.EightSchools()
    ch = T.sample(model, T.NUTS(0.8), T.MCMCThreads(), 1000, 1024, seed=3, verbose=False,
                  diagnostics=True)
    spl = T.NUTS(0.8)
    ref = O.sample_chains(oracle_of(model), oracle_config(spl, 500, 1003), 64, 1499, first_keep=500)
    sr = dg.summarize(ref["params"])
    gp = ch.get_params()                                # [iters, P, chains]
    mean = gp.mean(axis=(0, 2))
    var = gp.transpose(1, 0, 2).reshape(10, -1).var(axis=1, ddof=1)
    # oracle's MCSE dominates (64 chains vs 1024)
    assert np.abs((mean - sr["mean"]) / sr["mcse"]).max() < 4.0
    assert np.abs(var / sr["var"] - 1).max() < 0.15
    assert ch.rhat().max() < 1.01 and sr["rhat"].max() < 1.01
    assert ch.ess_bulk().min() > 0.2 * 1024 * 1000


@pytest.mark.parametrize("which", ["logistic", "hier", "linreg"])
def test_big_families_agree_with_oracle_posterior(built, which):
    """posterior means within 4 MCSE and variances within 15 % of the oracle's NUTS, R-hat < 1.01,
    for the families that run on the pump / CTA-per-chain strategies"""
    if which == "logistic":
        model = T.models.synthetic_logistic(400, 6, seed=8)
    elif which == "hier":
        model = T.models.synthetic_hier(12, 30, seed=9)   # informative groups: no funnel
    else:
        model = T.models.config1_linear_regression()
    spl = T.NUTS(0.8)
    ch = T.sample(model, spl, T.MCMCThreads(), 600, 512, seed=13, verbose=False, diagnostics=True)
    ref = O.sample_chains(oracle_of(model), oracle_config(spl, 300, 2013), 64, 899, first_keep=300)
    sr = dg.summarize(ref["params"])
    gp = ch.get_params()
    P = model.D
    mean = gp.mean(axis=(0, 2))
    var = gp.transpose(1, 0, 2).reshape(P, -1).var(axis=1, ddof=1)
    assert ch.rhat().max() < 1.01 and sr["rhat"].max() < 1.02
    assert np.abs((mean - sr["mean"]) / sr["mcse"]).max() < 4.5
    assert np.abs(var / sr["var"] - 1).max() < 0.2
    assert ch.size() == (600, P + 12, 512)
    np.testing.assert_allclose(ch["logjoint"], ch["logprior"] + ch["loglikelihood"], rtol=1e-10)


def test_single_chain_and_odd_counts(built):
    """ragged sizes: 1 chain, chain counts that are not multiples of the CTA / tile sizes"""
    for model in (T.gdemo_default(), T.models.synthetic_logistic(77, 3, seed=1), T.models.synthetic_hier(3, 5)):
        for n in (1, 13):
            ch = T.sample(model, T.NUTS(20, 0.8), T.MCMCThreads(), 30, n, seed=2, verbose=False)
            assert ch.size() == (30, model.D + 12, n)
            assert np.all(np.isfinite(ch.value))
    ch = T.sample(T.gdemo_default(), T.NUTS(20, 0.8), 25, seed=2, verbose=False)
    assert ch.size() == (25, 14, 1)


def test_device_diagnostics_match_numpy(built):
    ch = T.sample(T.EightSchools(), T.NUTS(0.8), T.MCMCThreads(), 400, 48, seed=4, verbose=False,
                  diagnostics=True)
    gp = ch.get_params()
    for p in range(10):
        x = gp[:, p, :]
        assert ch.ess_bulk()[p] == pytest.approx(dg.ess_bulk(x), rel=1e-6)
        assert ch.rhat()[p] == pytest.approx(dg.rhat(x), rel=1e-8)


def test_same_seed_same_chains_any_sharding(built):
    """test/mcmc/hmc.jl:164-171, test/mcmc/chains.jl:92-105: same rng => identical chains,
    different rng => different; and chain k does not depend on how many chains run beside it"""
    m = T.models.config1_linear_regression()
    a = T.sample(m, T.NUTS(), T.MCMCThreads(), 200, 8, seed=42, verbose=False)
    b = T.sample(m, T.NUTS(), T.MCMCThreads(), 200, 8, seed=42, verbose=False)
    c = T.sample(m, T.NUTS(), T.MCMCThreads(), 200, 8, seed=43, verbose=False)
    assert np.array_equal(a.value, b.value)
    assert not np.array_equal(a.value, c.value)
    d = T.sample(m, T.NUTS(), T.MCMCThreads(), 200, 200, seed=42, verbose=False)
    assert np.array_equal(a.value, d.value[:, :, :8])
    rng1, rng2 = np.random.default_rng(7), np.random.default_rng(7)
    e = T.sample(m, T.NUTS(), T.MCMCThreads(), 50, 2, rng=rng1, verbose=False)
    f = T.sample(m, T.NUTS(), T.MCMCThreads(), 50, 2, rng=rng2, verbose=False)
    assert np.array_equal(e.value, f.value)


def test_sizes_discard_adapt_and_initial_params(built):
    m = T.gdemo_default()
    # test/mcmc/hmc.jl:144-152
    ch = T.sample(m, T.NUTS(100, 0.8), 300, seed=1, verbose=False)
    assert ch.size() == (300, 14, 1)
    ch = T.sample(m, T.NUTS(100, 0.8), 300, seed=1, discard_adapt=False, verbose=False)
    assert ch.size() == (300, 14, 1)
    # first un-discarded sample = the initial parameters, with missing stats (hmc.jl:173-197)
    ip = [T.InitFromParams({"s": 4.0, "m": -1.0}), T.InitFromParams({"s": 0.5, "m": 2.0})]
    ch = T.sample(m, T.NUTS(0.8), T.MCMCThreads(), 20, 2, seed=1, discard_adapt=False,
                  initial_params=ip, verbose=False)
    np.testing.assert_allclose(ch["s"][0], [4.0, 0.5], rtol=1e-14)
    np.testing.assert_allclose(ch["m"][0], [-1.0, 2.0], rtol=1e-14)
    assert np.all(np.isnan(ch["tree_depth"][0])) and not np.any(np.isnan(ch["tree_depth"][1:]))
    assert np.all(np.isfinite(ch["logjoint"][0]))
    # thinning
    ch = T.sample(m, T.NUTS(50, 0.8), T.MCMCThreads(), 40, 3, seed=2, thinning=3, verbose=False)
    assert ch.size() == (40, 14, 3)
    # stats keys / user-space log-prob identity (test/test_utils/sampler.jl:14-29)
    np.testing.assert_allclose(ch["logjoint"], ch["logprior"] + ch["loglikelihood"], rtol=1e-12)
    assert ch.info["step_size"].dtype == np.float64                      # getstepsize: Float64
    assert np.all(ch["n_steps"] >= 1) and np.all(ch["tree_depth"] <= 10)


def test_save_state_resume(built):
    """test/mcmc/hmc.jl:199-218: resume with initial_state; here resuming is exactly the
    uninterrupted run because every random draw is addressed by (chain, iteration)"""
    m = T.EightSchools()
    full = T.sample(m, T.NUTS(60, 0.8), T.MCMCThreads(), 100, 16, seed=8, discard_adapt=True,
                    verbose=False)
    first = T.sample(m, T.NUTS(60, 0.8), T.MCMCThreads(), 50, 16, seed=8, save_state=True,
                     verbose=False)
    rest = T.sample(m, T.NUTS(60, 0.8), T.MCMCThreads(), 50, 16, initial_state=T.loadstate(first),
                    verbose=False)
    assert np.array_equal(first.value, full.value[:50])
    assert np.array_equal(rest.value, full.value[50:])
    with pytest.raises(ValueError):
        T.loadstate(full)


def test_divergence_warning(built):
    """post_sample_hook (src/mcmc/hmc.jl:476-486; regex of test/mcmc/abstractmcmc.jl:238)"""
    m = T.EightSchools()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        ch = T.sample(m, T.NUTS(0, 0.8, init_eps=3.0), T.MCMCThreads(), 200, 64, seed=1)
    n = int(np.nansum(ch["numerical_error"]))
    assert n > 0
    assert any("There were %d divergent transitions" % n in str(x.message) for x in w)


@pytest.mark.parametrize("name", ["gdemo", "eight_schools", "logistic_small", "hier_small"])
def test_sghmc_and_sgld_match_oracle(built, name):
    """SGHMC / SGLD (src/mcmc/sghmc.jl:79-105, 223-249) on the batched gradient kernels: same
    Philox addresses as the oracle, so the chains agree draw for draw"""
    from helpers import oracle_config, oracle_of, small_models
    import oracle as O
    model = small_models()[name]
    # the hierarchical model started from U(-2, 2) has gradients of 1e4: keep SGHMC stable there
    lr = 1e-7 if name == "hier_small" else 1e-4
    for spl in (T.SGHMC(learning_rate=lr, momentum_decay=0.3),
                T.SGLD(stepsize=T.PolynomialStepsize(2 * lr, 3.0, 0.6))):
        C, n_tr = 70, 25
        eng = T.Engine(model, spl, C, 0, seed=21)
        eng.set_option("keep_theta", 1)
        eng.init()
        eng.run(n_tr, 1, 1)
        out = eng.fetch(theta=True)
        ref = O.sample_chains(oracle_of(model), oracle_config(spl, 0, 21), C, n_tr)
        np.testing.assert_allclose(out["theta"].transpose(2, 0, 1), ref["theta"], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(out["params"].transpose(2, 0, 1), ref["params"], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(out["logp"].transpose(2, 0, 1), ref["logp"], rtol=1e-9, atol=1e-8)
        st, rst = out["stats"].transpose(2, 0, 1), ref["stats"]
        np.testing.assert_allclose(st[:, :, [0, 2, 8]], rst[:, :, [0, 2, 8]], rtol=1e-9, atol=1e-8)
        assert np.all(np.isnan(st[:, :, [1, 3, 4, 5, 6, 7]]))
        assert eng.grad_evals == ref["n_grad"]
        eng.close()


def test_sgld_through_sample_reference_style(built):
    """test/mcmc/sghmc.jl:40-53 through the public call: first sample = initial parameters,
    SGLD_stepsize column, step-size weighted means (median over chains: no accept step, an early
    overshoot of a single chain is never forgotten)"""
    chain = T.sample(T.gdemo_default(), T.SGLD(stepsize=T.PolynomialStepsize(0.5)), T.MCMCThreads(),
                     10000, 64, seed=1, verbose=False)
    ss = chain["SGLD_stepsize"]
    assert chain.value.shape[0] == 10000 and np.all(np.isnan(ss[0]))       # the init step's sample
    np.testing.assert_allclose(ss[1:4, 0], 0.5 / np.arange(1, 4) ** 0.55, rtol=1e-13)
    w = ss[1:] / ss[1:].sum(axis=0)
    assert abs(np.median((chain["s"][1:] * w).sum(axis=0)) - 49 / 24) < 0.2
    assert abs(np.median((chain["m"][1:] * w).sum(axis=0)) - 7 / 6) < 0.2
    np.testing.assert_allclose(chain["logjoint"], chain["logprior"] + chain["loglikelihood"], rtol=1e-12)


def test_dynamic_scheduling_is_bit_identical(built):
    """thread-per-chain strategy with more chain groups than resident CTAs: the task-queue kernel
    (k_chain_run_dyn: sub-ranges of the transitions, state handed over through global memory, any
    SM) must reproduce the static grid bit for bit, warm-up adaptation included"""
    outs = []
    for dyn in (1, 0):
        eng = T.Engine(T.EightSchools(), T.NUTS(40, 0.8), 40000, 40, seed=17, strategy=1)
        eng.set_option("dyn_sched", dyn)
        eng.init()
        eng.run(40 + 59, 40, 1)          # 99 transitions: warm-up windows and 60 kept draws
        outs.append((eng.fetch_draws(), eng.stepsize(), eng.inv_metric(), eng.grad_evals))
        eng.close()
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1], outs[1][1])
    np.testing.assert_array_equal(outs[0][2], outs[1][2])
    assert outs[0][3] == outs[1][3]


def test_sghmc_resume_carries_the_velocity(built):
    """SGHMCState(logdensity, params, velocity) (src/mcmc/sghmc.jl:49-53): the state blob holds the
    velocity, so a resumed SGHMC run continues the uninterrupted one bit for bit"""
    m = T.models.synthetic_logistic(300, 5, seed=4)
    spl = T.SGHMC(learning_rate=1e-4, momentum_decay=0.2)
    full = T.sample(m, spl, T.MCMCThreads(), 41, 24, seed=5, verbose=False)
    first = T.sample(m, spl, T.MCMCThreads(), 21, 24, seed=5, save_state=True, verbose=False)
    rest = T.sample(m, spl, T.MCMCThreads(), 20, 24, initial_state=T.loadstate(first), verbose=False)
    np.testing.assert_array_equal(first.value[1:], full.value[1:21])   # NaN stats columns compare equal
    np.testing.assert_array_equal(rest.value, full.value[21:])


def test_warns_after_ten_failed_initial_draws(built):
    """src/mcmc/abstractmcmc.jl:62-69: a warning at the tenth failed attempt, then the error with the
    initial-parameters URL once 1000 attempts are spent (test/mcmc/hmc.jl:284-296)"""
    model = T.GDemo(float("nan"), 2.0)          # lp is NaN everywhere
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        with pytest.raises(RuntimeError, match="initial-parameters"):
            T.sample(model, T.NUTS(5, 0.8), T.MCMCThreads(), 4, 8, seed=3, verbose=False)
    assert any("failed to find valid initial parameters in 10 tries" in str(x.message) for x in w)
    # a healthy model does not warn
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        T.sample(T.gdemo_default(), T.NUTS(5, 0.8), T.MCMCThreads(), 4, 8, seed=3, verbose=False)
    assert not any("initial parameters" in str(x.message) for x in w)


def test_reference_posterior_pins_of_the_constrained_support_models(built):
    """the reference's own HMC tests on its own data (test/mcmc/hmc.jl:24-62), every chain held to the
    reference's tolerance: Beta-Bernoulli mean(p) = 10/14 (atol 0.1) with HMC(1.5, 3); Dirichlet-
    Categorical mean(ps) = [5/16, 11/16] (atol 0.015) with HMC(0.75, 2); 1000 draws each.  The simplex
    model has D = 4 linked dimensions and 6 chain columns."""
    import oracle as O
    from helpers import oracle_config, oracle_of
    m1 = T.BetaBernoulli([0, 1, 0, 1, 1, 1, 1, 1, 1, 1])
    ch = T.sample(m1, T.HMC(1.5, 3), T.MCMCThreads(), 1000, 64, seed=123, verbose=False)
    assert ch.value.shape == (1000, 1 + 12, 64) and ch.names[0] == "p"
    assert np.all(np.abs(ch["p"][1:].mean(axis=0) - 10 / 14) < 0.1)
    ref = O.sample_chains(oracle_of(m1), oracle_config(T.HMC(1.5, 3), 0, 123), 64, 999)
    np.testing.assert_allclose(ch["p"][1:41].T, ref["params"][:, :40, 0], rtol=1e-9)      # draw for draw
    m2 = T.DirichletCategorical([1, 2, 1, 2, 2, 2, 2, 2, 2, 2])
    assert (m2.D, m2.P) == (4, 6)
    ch = T.sample(m2, T.HMC(0.75, 2), T.MCMCThreads(), 1000, 64, seed=123, verbose=False)
    assert ch.value.shape == (1000, 6 + 12, 64) and ch.names[:3] == ("ps[1]", "ps[2]", "pd[1]")
    means = ch.get_params()[1:].mean(axis=0)                                              # [6, chains]
    assert np.all(np.abs(means[0] - 5 / 16) < 0.015) and np.all(np.abs(means[1] - 11 / 16) < 0.015)
    np.testing.assert_allclose(ch["ps[1]"] + ch["ps[2]"], 1.0, rtol=1e-12)
    np.testing.assert_allclose(ch.get_params()[:, 2:].sum(axis=1), 1.0, rtol=1e-12)
    ref = O.sample_chains(oracle_of(m2), oracle_config(T.HMC(0.75, 2), 0, 123), 64, 999)
    np.testing.assert_allclose(ch.get_params()[1:41].transpose(2, 0, 1), ref["params"][:, :40], rtol=1e-8)
    # NUTS with the default diagonal metric on the simplex model, initial_params in user space
    ch = T.sample(m2, T.NUTS(200, 0.8), T.MCMCThreads(), 500, 32, seed=5, verbose=False,
                  initial_params=[T.InitFromParams({"ps": [0.5, 0.5], "pd": [0.25] * 4})] * 32)
    assert abs(ch["ps[1]"].mean() - 5 / 16) < 0.01


def test_abstractmcmc_step_interface_matches_sample(built):
    """src/mcmc/hmc.jl:156-264: step(rng, model, spl; ...) gives the initial point and a state with
    i = 0, step(rng, model, spl, state; ...) one transition; iterating it reproduces sample()'s chain"""
    model = T.EightSchools()
    spl = T.NUTS(20, 0.8)
    ch = T.sample(model, spl, 30, seed=77, nadapts=20, discard_initial=0, verbose=False)
    tr, st = T.step(None, model, spl, seed=77, nadapts=20)
    assert st.i == 0 and set(tr.params) == set(model.param_names)
    assert np.isnan(tr.stats["n_steps"]) and np.isfinite(tr.stats["logjoint"])
    rows = [tr]
    for _ in range(29):
        tr, st = T.step(None, model, spl, st, nadapts=20)
        rows.append(tr)
    assert st.i == 29
    for n in ("μ", "τ", "θ̃[3]"):
        np.testing.assert_array_equal(np.array([r.params[n] for r in rows]), ch[n][:, 0])
    np.testing.assert_array_equal(np.array([r.stats["n_steps"] for r in rows[1:]]), ch["n_steps"][1:, 0])
    p, s, e = T.params_with_stats(model, spl, rows[-1], st, stats=True)
    assert [k for k, _ in p] == list(model.param_names) and "acceptance_rate" in s and e == {}
    st.engine.close()


def test_callback_sees_every_kept_iteration(built):
    """AbstractMCMC.mcmcsample's `callback(rng, model, sampler, transition, state, iteration)`; the chain is
    the one sample() returns without a callback"""
    model = T.gdemo_default()
    spl = T.NUTS(30, 0.65)
    seen = []
    ch = T.sample(model, spl, T.MCMCThreads(), 40, 3, seed=5, verbose=False, chain_type=T.VNChain,
                  callback=lambda rng, m, s, tr, st, i: seen.append((i, tr.params["m"].copy(), st.i)))
    ref = T.sample(model, spl, T.MCMCThreads(), 40, 3, seed=5, verbose=False)
    assert [i for i, _, _ in seen] == list(range(1, 41))
    np.testing.assert_array_equal(np.stack([m for _, m, _ in seen]), ref["m"])
    np.testing.assert_array_equal(ch["m"], ref["m"])
    assert isinstance(ch, T.VNChain) and ch.parameters == ("s", "m")
