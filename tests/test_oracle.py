"""CPU: pins the oracle (oracle/tnuts_oracle.c) against the golden vectors (mpmath, independent
of both implementations), against the known answers the reference's own tests hold for this
path, and against independent evaluators (torch fp64 autograd, finite differences)."""
import numpy as np
import pytest

import oracle as O
import turing_jl_b200 as T
from oracle import diagnostics as dg

from helpers import golden_cases, model_from_golden, oracle_config, oracle_of


@pytest.fixture(scope="module")
def golden(built):
    return golden_cases()


def test_lpgrad_matches_golden(golden):
    for case in golden["cases"]:
        om = oracle_of(model_from_golden(case))
        for pt in case["points"]:
            lp, g = om.lpgrad(pt["theta"])
            assert lp == pytest.approx(pt["lp"], rel=1e-13, abs=1e-13), case["name"]
            np.testing.assert_allclose(g, pt["grad"], rtol=1e-12, atol=1e-12, err_msg=case["name"])


def test_reference_known_answers(golden):
    ka = golden["reference_known_answers"]
    gd = oracle_of(model_from_golden(golden["cases"][0]))
    # survey 8c seeded values (verified against the gdemo definition test/test_utils/models.jl:15-21)
    assert gd.lpgrad([0.0, 0.0])[0] == pytest.approx(ka["gdemo_lp_at_0_0"], abs=1e-14)
    lp, g = gd.lpgrad([0.3, -0.4])
    assert lp == pytest.approx(-7.362044505871338, abs=1e-13)
    np.testing.assert_allclose(g, [2.2524534835935, 3.4818456372041], rtol=1e-12)
    # logdensity(ldf, [0]) == logpdf(Normal(), 0)   (test/main.jl:43-45)
    sn = O.OracleModel(O.STD_NORMAL, D=1)
    assert sn.lpgrad([0.0])[0] == pytest.approx(ka["std_normal_lp_at_0"], abs=1e-15)

    # MAP / MLE of gdemo in USER space pin the log-joint surface
    # (test/optimisation/Optimisation.jl:246-248, 299-301)
    def user(fn_idx):
        def f(s, m):
            return gd.constrain([np.log(s), m])[1][fn_idx]
        return f
    for idx, key in ((2, "gdemo_map_user_space"), (1, "gdemo_mle")):
        f, s0, m0 = user(idx), ka[key]["s"], ka[key]["m"]
        h = 1e-5
        ds = (f(s0 + h, m0) - f(s0 - h, m0)) / (2 * h)
        dm = (f(s0, m0 + h) - f(s0, m0 - h)) / (2 * h)
        assert abs(ds) < 1e-6 and abs(dm) < 1e-6, key
        assert f(s0, m0) > max(f(s0 * 1.01, m0), f(s0 * 0.99, m0), f(s0, m0 + 0.01), f(s0, m0 - 0.01))


def test_user_space_logp_identities(golden):
    """logjoint = logprior + loglikelihood, and the linked lp = logjoint + log|J|
    (test/test_utils/sampler.jl:14-29)"""
    for case in golden["cases"]:
        m = model_from_golden(case)
        om = oracle_of(m)
        for pt in case["points"]:
            th = np.array(pt["theta"])
            p, lp3 = om.constrain(th)
            assert lp3[2] == pytest.approx(lp3[0] + lp3[1], rel=1e-14)
            logjac = m.logabsdetjac(th)       # log link: sum of the thetas; logit / stick breaking: Bijectors
            assert om.lpgrad(th)[0] == pytest.approx(lp3[2] + logjac, rel=1e-12, abs=1e-12)
            np.testing.assert_allclose(p, m.invlink(th), rtol=1e-14)


def test_gradient_vs_torch_autograd(golden):
    """independent fp64 evaluator: torch autograd on a plain-torch restatement of logistic regression"""
    import torch
    case = [c for c in golden["cases"] if c["name"] == "logistic"][0]
    m = model_from_golden(case)
    om = oracle_of(m)
    X = torch.tensor(m.X, dtype=torch.float64)
    y = torch.tensor(m.y, dtype=torch.float64)
    for pt in case["points"]:
        th = torch.tensor(pt["theta"], dtype=torch.float64, requires_grad=True)
        eta = X @ th
        lp = (torch.distributions.Normal(0.0, 1.0).log_prob(th).sum()
              - torch.nn.functional.binary_cross_entropy_with_logits(eta, y, reduction="sum"))
        lp.backward()
        olp, og = om.lpgrad(pt["theta"])
        assert olp == pytest.approx(lp.item(), rel=1e-13)
        np.testing.assert_allclose(og, th.grad.numpy(), rtol=1e-11, atol=1e-12)


def test_philox_known_answer(built):
    # Random123 known-answer test vectors for philox4x32-10
    assert O.philox(0, 0, 0, 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    key = 0xffffffffffffffff
    assert O.philox(key, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == (
        0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    key = 0xa4093822 | (0x299f31d0 << 32)
    assert O.philox(key, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344) == (
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    u = np.array([O.uniform2(7, c, 1, 0, 0) for c in range(2000)])
    assert 0 < u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.02
    z = np.array([O.normal2(7, c, 1, 1, 0) for c in range(4000)]).ravel()
    assert abs(z.mean()) < 0.05 and abs(z.std() - 1) < 0.05


def test_window_schedule(built):
    # Stan's schedule for 1000 / 500 warm-up iterations
    assert O.window_splits(1000) == (76, 950, [100, 150, 250, 450, 950])
    assert O.window_splits(500) == (76, 450, [100, 150, 250, 450])
    ws, we, sp = O.window_splits(100)
    assert (ws, we, sp) == (16, 90, [90])
    assert O.window_splits(10)[2] == []


def test_leapfrog_reversible_and_energy(built):
    m = O.OracleModel(O.EIGHT_SCHOOLS, D=10, N=8, y=[28, 8, -3, 7, -1, 1, 18, 12],
                      sigma=[15, 10, 16, 11, 9, 11, 10, 18], hyper=(5.0, 5.0))
    rng = np.random.default_rng(3)
    th, r, minv = rng.standard_normal(10) * 0.5, rng.standard_normal(10), np.exp(rng.standard_normal(10) * 0.3)
    tt, tr, tH = O.leapfrog_trajectory(m, th, r, minv, 0.05, 40)
    assert np.abs(tH - tH[0]).max() < 0.05          # symplectic: bounded energy error
    bt, br, _ = O.leapfrog_trajectory(m, tt[-1], -tr[-1], minv, 0.05, 40)
    np.testing.assert_allclose(bt[-1], th, atol=1e-10)  # time reversible


def test_posterior_known_answers_gdemo(built):
    """E[s] = 49/24, E[m] = 7/6 (test/mcmc/hmc.jl:138-142 uses atol 0.2; we require 4 MCSE)"""
    m = O.OracleModel(O.GDEMO, D=2, N=2, y=[1.5, 2.0], hyper=(2.0, 3.0))
    for mode in (0, 1):
        cfg = O.make_config(delta=0.8, n_adapts=1000, seed=2 + mode, sampler_mode=mode)
        r = O.sample_chains(m, cfg, 32, 4000, first_keep=1001)
        s = dg.summarize(r["params"])
        z = (s["mean"] - np.array([49 / 24, 7 / 6])) / s["mcse"]
        assert np.abs(z).max() < 4.0, (mode, z)
        assert s["rhat"].max() < 1.01
        assert abs(s["mean"][0] - 49 / 24) < 0.2 and abs(s["mean"][1] - 7 / 6) < 0.2


def test_posterior_known_answers_demo_models(built):
    """The analytic posterior means the reference's `DEMO_MODELS` sweep is held to
    (test/mcmc/external_sampler.jl:198-230: NUTS with n_adapts = 1000, discard_initial = 1000, 2000
    draws, rtol 0.2 against DynamicPPL.TestUtils.posterior_mean).  Every demo model is the gdemo prior
    s ~ InverseGamma(2, 3), m ~ Normal(0, sqrt(s)); the vector-valued ones observe 1.5 through
    (s[1], m[1]) and 2.0 through (s[2], m[2]), i.e. two independent one-observation blocks with
    E[s] = (3 + x^2 / 4) / 1.5 = [19/8, 8/3] and E[m] = x / 2 = [3/4, 1]; the scalar ones are gdemo
    itself (49/24, 7/6, above)."""
    for x, es, em in ((1.5, 19 / 8, 3 / 4), (2.0, 8 / 3, 1.0)):
        m = O.OracleModel(O.GDEMO, D=2, N=1, y=[x], hyper=(2.0, 3.0))
        cfg = O.make_config(delta=0.8, n_adapts=1000, seed=5, sampler_mode=1)
        r = O.sample_chains(m, cfg, 32, 3000, first_keep=1001)
        s = dg.summarize(r["params"])
        np.testing.assert_allclose(s["mean"], [es, em], rtol=0.2)          # the reference's tolerance
        assert np.abs((s["mean"] - np.array([es, em])) / s["mcse"]).max() < 4.0
        assert s["rhat"].max() < 1.01


def test_sampler_modes_same_tree(built):
    """AdvancedHMC merge sampling (mode 0) and the streaming form (mode 1) build the SAME tree:
    n_steps, depth, acceptance statistic and divergence flag of the first transition agree."""
    m = O.OracleModel(O.EIGHT_SCHOOLS, D=10, N=8, y=[28, 8, -3, 7, -1, 1, 18, 12],
                      sigma=[15, 10, 16, 11, 9, 11, 10, 18], hyper=(5.0, 5.0))
    a = O.sample_chains(m, O.make_config(delta=0.8, seed=5, sampler_mode=0), 64, 1)
    b = O.sample_chains(m, O.make_config(delta=0.8, seed=5, sampler_mode=1), 64, 1)
    for col in (0, 1, 5, 6, 7, 8):
        np.testing.assert_array_equal(a["stats"][:, 0, col], b["stats"][:, 0, col])


def test_candidate_law_equivalence(built):
    """the two candidate-selection forms pick leaves with the same probabilities: compare the
    distribution of the first transition's log_density over many seeds (two-sample KS)"""
    from scipy import stats
    m = O.OracleModel(O.GDEMO, D=2, N=2, y=[1.5, 2.0], hyper=(2.0, 3.0))
    th0 = np.tile([0.2, 1.0], (4000, 1))
    out = []
    for mode in (0, 1):
        cfg = O.make_config(delta=0.8, seed=99, sampler_mode=mode, init_eps=0.4)
        out.append(O.sample_chains(m, cfg, 4000, 1, theta0=th0)["stats"][:, 0, 2])
    assert stats.ks_2samp(out[0], out[1]).pvalue > 1e-3


def test_diagnostics_against_known(built):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1000, 8))
    e = dg.ess_bulk(x)
    assert 6500 < e < 9500           # iid draws: ESS ~ S
    assert abs(dg.rhat(x) - 1.0) < 0.01
    # AR(1) with rho = 0.9: ESS/S ~ (1-rho)/(1+rho)
    y = np.zeros((4000, 4))
    for t in range(1, 4000):
        y[t] = 0.9 * y[t - 1] + rng.standard_normal(4) * np.sqrt(1 - 0.81)
    e = dg.ess_bulk(y)
    assert 0.5 < e / (16000 * 0.1 / 1.9) < 1.6
    # shifted chains are detected
    x[:, 0] += 1.0
    assert dg.rhat(x) > 1.05


def test_sghmc_and_sgld_oracle_semantics():
    """src/mcmc/sghmc.jl:79-105, 223-249 restated.  The reference pins them on gdemo
    (test/mcmc/sghmc.jl:23-53: means 49/24 and 7/6, atol 0.1 for SGHMC(0.02, 0.5) on 10 000 samples,
    atol 0.2 for the step-size weighted SGLD(PolynomialStepsize(0.5)) means on 20 000) for ONE seed.
    Neither sampler has an accept step: an SGLD chain whose first (largest, most heavily weighted)
    steps overshoot carries that forever, so the check here is the median over 16 chains; SGHMC has a
    discretisation bias of about +0.15 on s, hence atol 0.25 on the pooled mean."""
    m = O.OracleModel(0, D=2, N=2, y=[1.5, 2.0], hyper=(2.0, 3.0))
    cfg = O.make_config(algorithm=4, init_eps=0.5, delta=0.0, lambda_=0.55, seed=3)
    r = O.sample_chains(m, cfg, 16, 10000, first_keep=1)
    eps = r["stats"][0, :, 8]
    np.testing.assert_allclose(eps[:4], 0.5 / np.arange(1, 5) ** 0.55, rtol=1e-14)   # SGLD_stepsize
    w = eps / eps.sum()
    s_w = np.median((r["params"][:, :, 0] * w).sum(axis=1))    # weighted by the step sizes
    m_w = np.median((r["params"][:, :, 1] * w).sum(axis=1))
    assert abs(s_w - 49 / 24) < 0.2 and abs(m_w - 7 / 6) < 0.2, (s_w, m_w)
    assert r["n_grad"] == 16 * (10000 + 1)
    assert np.all(r["stats"][:, :, 0] == 1.0) and np.all(np.isnan(r["stats"][:, :, 1]))
    cfg = O.make_config(algorithm=3, init_eps=0.02, lambda_=0.5, seed=4)
    r = O.sample_chains(m, cfg, 16, 10000, first_keep=1)
    assert abs(r["params"][:, :, 0].mean() - 49 / 24) < 0.25
    assert abs(r["params"][:, :, 1].mean() - 7 / 6) < 0.1
    # the user-space log-prob identity holds for these samplers too (test_chain_logp_metadata)
    np.testing.assert_allclose(r["logp"][:, :, 2], r["logp"][:, :, 0] + r["logp"][:, :, 1], rtol=1e-13)


def test_unit_metric_keeps_identity_mass_matrix():
    """HMCDA / HMC default to UnitEuclideanMetric (src/mcmc/hmc.jl:76, 308, 323): the Stan windows and
    the dual-averaging restarts run as for NUTS, the mass matrix never moves"""
    m = O.OracleModel(O.EIGHT_SCHOOLS, D=10, N=8, y=[28, 8, -3, 7, -1, 1, 18, 12],
                      sigma=[15, 10, 16, 11, 9, 11, 10, 18], hyper=(5.0, 5.0))
    unit = O.sample_chains(m, O.make_config(delta=0.8, n_adapts=150, seed=2, metric=1), 4, 160)
    diag = O.sample_chains(m, O.make_config(delta=0.8, n_adapts=150, seed=2, metric=0), 4, 160)
    assert np.all(unit["minv"] == 1.0)
    assert np.all(diag["minv"] != 1.0)
    # before the first window ends (iteration 100 of 150: windows 76-100) the two runs are identical
    np.testing.assert_array_equal(unit["theta"][:, :99], diag["theta"][:, :99])
    assert not np.array_equal(unit["theta"][:, 120:], diag["theta"][:, 120:])


def test_dense_metric_oracle():
    """metricT = DenseEuclideanMetric restated (momentum through the Cholesky factor of M^-1, kinetic
    energy and U-turn products with M^-1 r, Welford covariance with the Stan regulariser).  With the
    identity it is the diagonal metric bit for bit; after adaptation its diagonal is close to the
    diagonal metric's variances; gdemo's known posterior means hold (test/mcmc/hmc.jl:138-142)."""
    es = O.OracleModel(O.EIGHT_SCHOOLS, D=10, N=8, y=[28, 8, -3, 7, -1, 1, 18, 12],
                       sigma=[15, 10, 16, 11, 9, 11, 10, 18], hyper=(5.0, 5.0))
    dense = O.sample_chains(es, O.make_config(delta=0.8, n_adapts=150, seed=2, metric=2), 6, 160)
    diag = O.sample_chains(es, O.make_config(delta=0.8, n_adapts=150, seed=2, metric=0), 6, 160)
    np.testing.assert_array_equal(dense["theta"][:, :99], diag["theta"][:, :99])   # identity until window end
    assert not np.array_equal(dense["theta"][:, 120:], diag["theta"][:, 120:])
    ratio = dense["minv"] / diag["minv"]
    assert np.all(ratio > 0.3) and np.all(ratio < 3.0)
    gd = O.OracleModel(0, D=2, N=2, y=[1.5, 2.0], hyper=(2.0, 3.0))
    r = O.sample_chains(gd, O.make_config(delta=0.8, n_adapts=500, seed=5, metric=2), 8, 2500, first_keep=501)
    s = dg.summarize(r["params"])
    assert abs(r["params"][:, :, 0].mean() - 49 / 24) < 0.15 and abs(r["params"][:, :, 1].mean() - 7 / 6) < 0.1
    assert s["rhat"].max() < 1.02
    # HMCDA with the dense metric runs through the same pieces
    r = O.sample_chains(gd, O.make_config(delta=0.65, n_adapts=300, seed=6, metric=2, algorithm=2, lambda_=0.3),
                        8, 1500, first_keep=301)
    assert abs(r["params"][:, :, 1].mean() - 7 / 6) < 0.15


def test_resume_continues_bit_for_bit_and_chain_ids_address_the_streams():
    """orc_resume_chains (bench.py's CPU step unit) == the uninterrupted run; orc_sample_chain_ids
    runs any subset of the global chain ids"""
    m = T.models.synthetic_logistic(300, 7, seed=2)
    om = oracle_of(m)
    cfg = O.make_config(delta=0.8, n_adapts=30, seed=5)
    a = O.sample_chains(om, cfg, 6, 60)
    b = O.sample_chains(om, cfg, 6, 40)
    r = O.resume_chains(om, cfg, b["theta"][:, -1], b["eps"], b["minv"], 40, 20)
    assert np.array_equal(r["theta"], a["theta"][:, 40:])
    assert np.array_equal(r["stats"][..., 0], a["stats"][:, 40:, 0])
    assert b["n_grad"] + r["n_grad"] == a["n_grad"] + 6      # one re-evaluation per restored chain
    sub = O.sample_chains(om, cfg, 0, 60, chain_ids=[4, 1])
    assert np.array_equal(sub["theta"][0], a["theta"][4]) and np.array_equal(sub["theta"][1], a["theta"][1])


def test_reference_posterior_pins_of_the_constrained_support_models():
    """test/mcmc/hmc.jl:24-43: p ~ Beta(2,2), ten Bernoulli observations, HMC(1.5, 3), 1000 draws ->
    mean(p) = 10/14 (atol 0.1); test/mcmc/hmc.jl:45-62: ps ~ Dirichlet(2,3), pd ~ Dirichlet(4,1),
    ten Categorical observations, HMC(0.75, 2), 1000 draws -> mean(ps) = [5/16, 11/16] (atol 0.015).
    The reference runs one chain with one seed; here every one of 32 seeded chains must pass."""
    m1 = T.BetaBernoulli([0, 1, 0, 1, 1, 1, 1, 1, 1, 1])
    r = O.sample_chains(oracle_of(m1), oracle_config(T.HMC(1.5, 3), 0, 123), 32, 1000)
    assert np.all(np.abs(r["params"][:, :, 0].mean(axis=1) - 10 / 14) < 0.1)
    m2 = T.DirichletCategorical([1, 2, 1, 2, 2, 2, 2, 2, 2, 2])
    r = O.sample_chains(oracle_of(m2), oracle_config(T.HMC(0.75, 2), 0, 123), 32, 1000)
    means = r["params"].mean(axis=1)
    assert np.all(np.abs(means[:, :2] - [5 / 16, 11 / 16]) < 0.015)
    np.testing.assert_allclose(r["params"][:, :, :2].sum(axis=2), 1.0, rtol=1e-12)
    np.testing.assert_allclose(means[:, 2:].mean(axis=0), 0.25, atol=0.01)      # pd: the uniform prior
