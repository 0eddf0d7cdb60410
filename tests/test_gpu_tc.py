"""GPU: the tcgen05 (bf16x3) logistic-regression likelihood, engine option "logistic_tc"
(turing.jl_b200/csrc/logistic_tc.cuh), through the C ABI.

This mode trades the 1e-12 fp64 parity of the default DMMA path for tensor-core throughput; the
north star allows "a stated bound in fp32".  The bounds stated (include/tnuts.h) and tested here:
  log-likelihood ........ |d lp| <= 3e-7 * |lp|           (fp64 accumulation of fp32 terms; the tensor
                                                          core's fp32 accumulation rounds |eta| down by
                                                          ~1e-7 relative, a same-sign bias equivalent to
                                                          rescaling X by 1 - 1e-7)
  energy differences .... |d (lp_a - lp_b)| <= 1e-3 over a posterior-width cloud at N = 20000 (measured
                          3e-4; it is the same 1e-7 shrink of eta, i.e. c * theta . grad_lik, a smooth
                          function of theta and not noise)
  gradient .............. |d g|_2 <= 1e-4 * |g_lik|_2     (measured ~1e-6 .. 2e-5: fp32 accumulation
                                                          of the row sums in tensor memory)
  sampling .............. posterior means within 4 MCSE of the fp64 path, R-hat < 1.01; the same
                          seed gives bit-identical chains (the kernel is deterministic)
"""
import numpy as np
import pytest

import turing_jl_b200 as T

from helpers import oracle_of, small_models

pytestmark = pytest.mark.gpu

TC = {"logistic_tc": 1}


@pytest.mark.parametrize("N,D,n", [(70, 100, 5), (5017, 100, 200), (3001, 128, 130), (2000, 64, 257),
                                   (999, 20, 64), (300, 5, 33), (1, 3, 2), (64, 1, 1),
                                   # 128 < D <= 256: the CTA-pair kernel (k_logistic_tc_wide), incl. a second
                                   # half of one feature, ragged halves and several row chunks / chain tiles
                                   (3001, 256, 130), (777, 129, 5), (20011, 200, 257), (64, 160, 1), (5000, 145, 300)])
def test_tc_lpgrad_against_oracle(built, N, D, n):
    model = T.models.synthetic_logistic(N, D, seed=N + D)
    ldf = T.LogDensityFunction(model, options=TC)
    rng = np.random.default_rng(N)
    th = rng.standard_normal((n, D)) * 0.5
    lp, g = ldf.logdensity_and_gradient(th)
    olp, og = oracle_of(model).lpgrad(th)
    # the prior part is evaluated in fp64 by both; bounds are relative to the likelihood part
    prior_g = -th / model.hyper[0] ** 2
    glik = og - prior_g
    np.testing.assert_allclose(lp, olp, rtol=3e-7, atol=1e-6)
    err = np.linalg.norm(g - og, axis=1)
    assert np.all(err <= 1e-4 * np.linalg.norm(glik, axis=1) + 1e-6), err.max()
    # value-only entry and single-point form agree with the batched one
    np.testing.assert_array_equal(ldf.logdensity(th), lp)
    l1, g1 = ldf.logdensity_and_gradient(th[0])
    np.testing.assert_allclose(l1, lp[0], rtol=3e-7)


def test_tc_energy_differences_are_tight(built):
    """what the sampler consumes are lp DIFFERENCES along a trajectory"""
    N, D = 20000, 100
    model = T.models.synthetic_logistic(N, D, seed=3)
    ldf = T.LogDensityFunction(model, options=TC)
    ref = T.LogDensityFunction(model)
    rng = np.random.default_rng(0)
    centre = rng.standard_normal(D) * 0.2
    th = centre + rng.standard_normal((256, D)) * 0.02      # a posterior-width cloud
    lp = ldf.logdensity(th)
    rlp = ref.logdensity(th)
    d = (lp - lp[0]) - (rlp - rlp[0])
    assert np.max(np.abs(d)) < 1e-3, np.max(np.abs(d))


def test_tc_rejects_models_past_the_family_limit(built):
    """D <= 256 is the logistic family's limit on every path (fp64 DMMA and tcgen05)"""
    model = T.models.synthetic_logistic(300, 257, seed=5)
    for opts in (TC, None):
        with pytest.raises(T._lib.TnutsError):
            T.LogDensityFunction(model, options=opts).logdensity(np.zeros((2, 257)))


def test_tc_wide_matches_fp64_path_at_config5_width(built):
    """D = 256 (BASELINE configs[4]'s feature count), 256 chains, enough rows for the full row cut:
    the CTA-pair kernel against the fp64 DMMA kernel of the same engine, and bit-identical when
    evaluated twice (the partial-eta exchange between the two CTAs is deterministic)"""
    N, D, C = 150000, 256, 256
    model = T.models.synthetic_logistic(N, D, seed=9)
    rng = np.random.default_rng(1)
    th = rng.standard_normal((C, D)) * 0.1
    lp, g = T.LogDensityFunction(model, options=TC).logdensity_and_gradient(th)
    lp2, g2 = T.LogDensityFunction(model, options=TC).logdensity_and_gradient(th)
    np.testing.assert_array_equal(lp, lp2)
    np.testing.assert_array_equal(g, g2)
    rlp, rg = T.LogDensityFunction(model).logdensity_and_gradient(th)
    np.testing.assert_allclose(lp, rlp, rtol=3e-7)
    glik = rg + th / model.hyper[0] ** 2
    err = np.linalg.norm(g - rg, axis=1)
    assert np.all(err <= 1e-4 * np.linalg.norm(glik, axis=1)), (err / np.linalg.norm(glik, axis=1)).max()


def test_tc_wide_sampling_follows_the_fp64_chains(built):
    """the pump strategy on the CTA-pair kernel (work list, shrinking row cut): same seeds, nearly the same
    arithmetic, so the first transitions build the same trees as the fp64 path"""
    model = T.models.synthetic_logistic(900, 140, seed=12)
    outs = []
    for tc in (1, 0):
        eng = T.Engine(model, T.NUTS(20, 0.8), 200, 20, seed=3)
        eng.set_option("logistic_tc", tc)
        eng.set_option("keep_theta", 1)
        eng.init()
        eng.run(30, 1, 1)
        outs.append(eng.fetch(theta=True))
        eng.close()
    a, b = outs
    assert np.mean(a["stats"][:3, 0] == b["stats"][:3, 0]) > 0.95
    np.testing.assert_allclose(a["theta"][:2], b["theta"][:2], atol=1e-3)


def test_tc_sampling_is_deterministic_and_matches_fp64_posterior(built):
    model = T.models.synthetic_logistic(4000, 24, seed=8)
    kw = dict(seed=11, verbose=False, diagnostics=True)
    a = T.sample(model, T.NUTS(0.8), T.MCMCThreads(), 300, 256, engine_options=TC, **kw)
    b = T.sample(model, T.NUTS(0.8), T.MCMCThreads(), 300, 256, engine_options=TC, **kw)
    np.testing.assert_array_equal(a.value, b.value)
    ref = T.sample(model, T.NUTS(0.8), T.MCMCThreads(), 300, 256, **kw)
    assert np.max(a.info["rhat"]) < 1.01 and np.max(ref.info["rhat"]) < 1.01
    P = model.D
    ma, mr = a.value[:, :P].mean(axis=(0, 2)), ref.value[:, :P].mean(axis=(0, 2))
    sd = ref.value[:, :P].std(axis=(0, 2))
    mcse = sd / np.sqrt(np.minimum(a.info["ess_bulk"], ref.info["ess_bulk"]))
    assert np.all(np.abs(ma - mr) < 4 * np.sqrt(2) * mcse), np.max(np.abs(ma - mr) / mcse)
    va, vr = a.value[:, :P].var(axis=(0, 2)), ref.value[:, :P].var(axis=(0, 2))
    np.testing.assert_allclose(va, vr, rtol=0.05)
    # same sampler behaviour: step sizes and tree depths of the two arithmetic modes agree closely
    np.testing.assert_allclose(np.mean(a.info["step_size"]), np.mean(ref.info["step_size"]), rtol=0.02)


def test_tc_pump_work_list(built):
    """chains finish at different pumps: the tcgen05 kernel reads the compacted work list"""
    model = T.models.synthetic_logistic(700, 10, seed=2)
    eng = T.Engine(model, T.NUTS(20, 0.8), 200, 20, seed=3)
    eng.set_option("logistic_tc", 1)
    eng.set_option("keep_theta", 1)
    eng.init()
    eng.run(40, 1, 1)
    out = eng.fetch(theta=True)
    ref = T.Engine(model, T.NUTS(20, 0.8), 200, 20, seed=3)
    ref.set_option("keep_theta", 1)
    ref.init()
    ref.run(40, 1, 1)
    rout = ref.fetch(theta=True)
    # identical seeds and nearly identical arithmetic: the first transitions build the same trees
    st, rst = out["stats"], rout["stats"]
    same = np.mean(st[:3, 0] == rst[:3, 0])
    assert same > 0.95, same
    np.testing.assert_allclose(out["theta"][:2], rout["theta"][:2], atol=1e-3)
    eng.close()
    ref.close()


def test_tc_serves_hmc_hmcda_sghmc_and_resume(built):
    """every sampler of the pump strategy calls the same batched gradient entry, so the tcgen05
    kernel serves them all; a resumed run continues bit for bit (deterministic kernel, state blob)"""
    m = T.models.synthetic_logistic(900, 12, seed=6)
    ref_kw = dict(seed=4, verbose=False)
    for spl in (T.HMC(0.01, 5), T.HMCDA(30, 0.65, 0.1), T.SGHMC(learning_rate=1e-4, momentum_decay=0.2)):
        a = T.sample(m, spl, T.MCMCThreads(), 40, 130, engine_options=TC, **ref_kw)
        b = T.sample(m, spl, T.MCMCThreads(), 40, 130, **ref_kw)
        assert np.all(np.isfinite(a.value[:, :m.D]))
        if not spl.adaptive:
            # same seeds, nearly identical arithmetic: the first transitions after the initial point
            # agree closely with the fp64 path (after a warm-up phase the chains have decorrelated)
            np.testing.assert_allclose(a.value[1:3, :m.D], b.value[1:3, :m.D], rtol=1e-3, atol=1e-4)
        else:
            np.testing.assert_allclose(a.value[:, :m.D].mean(axis=(0, 2)), b.value[:, :m.D].mean(axis=(0, 2)),
                                       atol=0.05)
    spl = T.NUTS(20, 0.8)
    full = T.sample(m, spl, T.MCMCThreads(), 30, 130, seed=8, engine_options=TC, verbose=False)
    first = T.sample(m, spl, T.MCMCThreads(), 15, 130, seed=8, save_state=True, engine_options=TC, verbose=False)
    rest = T.sample(m, spl, T.MCMCThreads(), 15, 130, initial_state=T.loadstate(first), engine_options=TC,
                    verbose=False)
    assert np.array_equal(first.value, full.value[:15])
    assert np.array_equal(rest.value, full.value[15:])


@pytest.mark.parametrize("N,D,C,spl", [
    (700, 10, 200, T.NUTS(20, 0.8)),        # <= 512 chains: persistent from the first pump
    (20000, 100, 300, T.NUTS(15, 0.8)),     # 313 row tiles: the cut uses the whole grid
    (5000, 33, 1300, T.NUTS(20, 0.8)),      # host-pumped until <= 512 chains are left, then persistent
    (3000, 64, 130, T.HMC(0.01, 7)),        # static trajectories
    (900, 12, 520, T.HMCDA(30, 0.65, 0.1)),
])
def test_persistent_tail_kernel_is_bit_identical_to_the_host_pumped_path(built, N, D, C, spl):
    """tc_persist.cuh: gradient + reduction + NUTS state machine per leapfrog inside one cooperative
    kernel, same row cut, same summation order, same refresh schedule -> the very same chains"""
    model = T.models.synthetic_logistic(N, D, seed=N % 97)
    nad = spl.n_adapts if spl.adaptive else 0
    res = []
    for persist in (1, 0):
        e = T.Engine(model, spl, C, nad, seed=5)
        e.set_option("logistic_tc", 1)
        e.set_option("persist", persist)
        e.set_option("time_kernels", 1)
        e.init()
        e.run(nad + 12, 1, 1)
        res.append((e.fetch_draws(), e.get_state(), e.grad_evals, e.kernel_times()))
        e.close()
    (da, sa, ga, ka), (db, sb, gb, kb) = res
    assert ka[5] > 0 and kb[5] == 0          # the persistent kernel really ran / really did not
    assert ga == gb
    assert np.array_equal(da, db, equal_nan=True)
    assert np.array_equal(sa, sb)


@pytest.mark.parametrize("family,C,tc,persist", [
    ("logistic", 1300, 1, 1),     # host-pumped views 1300 -> 640 .. -> persistent views 512 -> 256 -> 128
    ("logistic", 1300, 1, 0),     # every view host-pumped
    ("logistic", 600, 0, 0),      # fp64 DMMA likelihood
    ("hier", 520, 0, 0),          # a chain-local family on the pump
])
def test_progressive_compaction_is_bit_identical(built, family, C, tc, persist):
    """lockstep.cu: as chains finish, the unfinished ones are gathered into narrower state arrays (column
    k of a view = chain gid[k]).  Per-chain arithmetic does not depend on the column, the gather points
    follow the refresh schedule -> identical draws, state blob and gradient count with option compact=0"""
    if family == "logistic":
        model = T.models.synthetic_logistic(3000, 33, seed=4)
    else:
        model = small_models()["hier_small"]
    spl = T.NUTS(20, 0.8)
    res = []
    for compact in (1, 0):
        e = T.Engine(model, spl, C, 20, seed=9, strategy=2)
        if tc:
            e.set_option("logistic_tc", 1)
        e.set_option("persist", persist)
        e.set_option("compact", compact)
        e.set_option("keep_theta", 1)
        e.init()
        e.run(32, 1, 1)
        out = e.fetch(theta=True)
        res.append((e.fetch_draws(), out["theta"], e.get_state(), e.grad_evals, e.launches))
        e.close()
    (da, ta, sa, ga, la), (db, tb, sb, gb, lb) = res
    assert ga == gb
    if tc:
        # the tcgen05 mode reads its finished-chain counter at fixed pump indices, so the gathers run at
        # fixed pumps; the other modes poll a mapped counter and gather when the host happens to see it
        # (a host far ahead of the GPU may issue the whole short run before any chain finishes)
        assert la != lb                   # the gathers really ran
    assert np.array_equal(da, db, equal_nan=True)
    assert np.array_equal(ta, tb, equal_nan=True)
    assert np.array_equal(sa, sb)
