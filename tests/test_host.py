"""CPU: host-side mirror of the reference interface and the C-ABI surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

import turing_jl_b200 as T
from turing_jl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "tnuts.h")).read()
    declared = set(re.findall(r"\b(tnuts_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libtnuts.so does not export " + name
    # the binding declares a prototype for each of them too
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert _lib.load().tnuts_abi_version() == 1


def test_struct_layout_matches_header(built):
    # tnuts_config / tnuts_model are passed by pointer: sizes must agree with the C compiler
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "tnuts.h"\nint main(){printf("%zu %zu\\n", sizeof(tnuts_model), sizeof(tnuts_config));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", os.path.join(d, "s")])
        a, b = map(int, subprocess.check_output([os.path.join(d, "s")]).split())
    assert a == ctypes.sizeof(_lib.TnutsModel)
    assert b == ctypes.sizeof(_lib.TnutsConfig)


def test_no_gpu_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.TnutsError) as ei:
        T.sample(T.gdemo_default(), T.NUTS(0.8), T.MCMCThreads(), 10, 2)
    assert "no CPU fallback" in str(ei.value)


def test_nuts_constructors():
    # src/mcmc/hmc.jl:377-402
    s = T.NUTS()
    assert (s.n_adapts, s.delta, s.max_depth, s.Delta_max, s.eps) == (-1, 0.65, 10, 1000.0, 0.0)
    s = T.NUTS(0.8)
    assert (s.n_adapts, s.delta) == (-1, 0.8)
    s = T.NUTS(1000, 0.8, max_depth=7, init_eps=0.1)
    assert (s.n_adapts, s.delta, s.max_depth, s.eps) == (1000, 0.8, 7, 0.1)
    with pytest.raises(TypeError):
        T.NUTS(1)          # NUTS(::Int) does not exist in the reference either
    h = T.HMC(0.05, 10)
    assert (h.eps, h.n_leapfrog, h.adaptive) == (0.05, 10, False)
    d = T.HMCDA(200, 0.65, 0.3)
    assert (d.n_adapts, d.delta, d.lam, d.adaptive) == (200, 0.65, 0.3, True)


def test_model_descriptors_and_link():
    m = T.EightSchools()
    assert m.D == 10 and m.param_names[:3] == ("μ", "τ", "θ̃[1]")
    th = np.arange(10.0) / 10
    np.testing.assert_allclose(m.link(m.invlink(th)), th)
    assert m.invlink(th)[1] == pytest.approx(np.exp(0.1))
    with pytest.raises(ValueError):
        m.link(-np.ones(10))
    lr = T.models.synthetic_logistic(50, 4)
    assert lr.D == 4 and lr.X.shape == (50, 4) and set(np.unique(lr.y)) <= {0.0, 1.0}
    h = T.models.synthetic_hier(3, 5)
    assert h.D == 11 and h.param_names[5] == "a[1]" and h.param_names[8] == "b[1]"
    cs = m.c_struct()
    assert (cs.family, cs.D, cs.N) == (2, 10, 8) and cs.hyper[0] == 5.0


def test_sample_argument_validation(built):
    m = T.gdemo_default()
    with pytest.raises(TypeError):
        T.sample(m, T.NUTS(0.8))
    with pytest.raises(ValueError) as ei:   # src/mcmc/abstractmcmc.jl:147-150
        T.sample(m, T.NUTS(0.8), T.MCMCThreads(), 10, 4, initial_params=[None, None])
    assert "one element per chain" in str(ei.value)

    class Disc(T.GDemo):
        def has_discrete_variables(self):
            return True
    with pytest.raises(ValueError):          # src/common.jl:31-46
        T.sample(Disc(), T.NUTS(0.8), T.MCMCThreads(), 10, 2)


def test_initial_params_resolution():
    from turing_jl_b200.inference import _resolve_initial_params
    m = T.gdemo_default()
    th = _resolve_initial_params(m, [T.InitFromParams({"s": 2.0, "m": 0.5}), [1.0, -1.0]], 2)
    np.testing.assert_allclose(th, [[np.log(2.0), 0.5], [0.0, -1.0]])
    assert _resolve_initial_params(m, None, 3) is None
    es = T.EightSchools()
    th = _resolve_initial_params(es, [{"μ": 1.0, "τ": 2.0, "θ̃": np.zeros(8)}], 1)
    assert th.shape == (1, 10) and th[0, 1] == pytest.approx(np.log(2.0))


def test_chains_container():
    from turing_jl_b200.chains import INTERNALS
    v = np.random.default_rng(0).standard_normal((5, 2 + len(INTERNALS), 3))
    c = T.Chains(v, ("s", "m"))
    assert c.size() == (5, 14, 3) and c["m"].shape == (5, 3)
    assert c.names[:5] == ("s", "m", "logprior", "loglikelihood", "logjoint")
    for k in ("tree_depth", "n_steps", "acceptance_rate", "hamiltonian_energy", "numerical_error"):
        assert k in c.names       # test/mcmc/callbacks.jl:61-64, src/mcmc/hmc.jl:480
    np.testing.assert_array_equal(c["nom_step_size"], c["step_size"])
    assert np.all(c["is_accept"] == 1.0)
    cc = T.chainscat([c, c])
    assert cc.size() == (5, 14, 6)


def test_sghmc_sgld_constructors():
    """test/mcmc/sghmc.jl:18-21, 35-38; src/mcmc/sghmc.jl:134-141 (decay rate check)"""
    import turing_jl_b200 as T
    a = T.SGHMC(learning_rate=0.01, momentum_decay=0.1)
    assert isinstance(a, T.SGHMC) and a.algorithm == 3 and not a.adaptive
    assert (a.eps, a.lam) == (0.01, 0.1)
    g = T.SGLD(stepsize=T.PolynomialStepsize(0.25))
    assert isinstance(g, T.SGLD) and g.algorithm == 4
    assert (g.eps, g.delta, g.lam) == (0.25, 0.0, 0.55)
    assert T.SGLD().stepsize.a == 0.01                      # default PolynomialStepsize(0.01)
    f = T.PolynomialStepsize(0.5, 2.0, 0.6)
    assert f(3) == pytest.approx(0.5 / 5.0 ** 0.6)
    for bad in (0.5, 1.01, 0.0):
        with pytest.raises(ValueError, match="decay rate"):
            T.PolynomialStepsize(0.1, 0.0, bad)
    with pytest.raises(TypeError):
        T.SGHMC(0.01, 0.1)                                   # keyword-only, like the reference


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver's reference arm) runs without a GPU and prints exactly
    ONE JSON line on stdout with the contract's keys (eight schools here: the default configs[2]
    workload takes minutes of host time per chain)"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--workload", "eight_schools", "--ref-quick", "--no-cache"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "grad_evals_per_sec" and d["unit"] == "grad_evals/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert "configs[1]" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "grad_evals/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}


def test_bench_cpu_block_steps_continue_the_full_job(tmp_path, monkeypatch):
    """the CPU arm of the configs[2] bench: whole schedule on `cores` chains (cached for the other arm),
    then steps that continue the adapted chains -- here on a small logistic model"""
    import importlib.util
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    monkeypatch.setattr(b.tempfile, "gettempdir", lambda: str(tmp_path))
    model = T.models.synthetic_logistic(300, 6, seed=3)
    model.N = 300
    spl = T.NUTS(0.8)
    # make the small model take the "heavy" path (cores chains, block steps)
    monkeypatch.setattr(b, "host_cores", lambda: 2)
    full, st = b.cpu_full_job("logistic_test", model, spl, 1000, quick=True)
    assert full["cached"] is False and full["chains"] == 256 and full["grad_evals"] > 0
    full2, st2 = b.cpu_full_job("logistic_test", model, spl, 1000, quick=True)
    assert isinstance(full2["cached"], str) and np.array_equal(st["theta"], st2["theta"])
    g, dt = b.cpu_block(model, spl, full, st, 3)
    assert g > 3 * full["chains"] and dt > 0
    cb = b.cpu_baseline_object(full, g / dt, "blocks")
    assert cb["kind"] == "port" and cb["whole_job_grad_evals_per_sec"] == full["value"]


def test_vnchain_groups_flat_columns_by_variable():
    """chain_type = VNChain (the reference's default, src/mcmc/Inference.jl:65): whole variables by
    name, elements by indexed name, sampler columns as extras, the saved state as last_sampler_state"""
    import turing_jl_b200 as T
    from turing_jl_b200.chains import INTERNALS
    names = ("μ", "τ") + tuple("θ̃[%d]" % (j + 1) for j in range(3))
    rng = np.random.default_rng(0)
    v = rng.standard_normal((5, len(names) + len(INTERNALS), 2))
    ch = T.VNChain(v, names, INTERNALS, {"samplerstate": {"blobs": [b"x"], "seed": 1}})
    assert ch.parameters == ("μ", "τ", "θ̃")
    assert ch["θ̃"].shape == (5, 2, 3)
    np.testing.assert_array_equal(ch["θ̃"][:, :, 1], v[:, 3, :])
    np.testing.assert_array_equal(ch["θ̃[2]"], v[:, 3, :])
    np.testing.assert_array_equal(ch["μ"], v[:, 0, :])
    np.testing.assert_array_equal(ch["logjoint"], v[:, len(names) + 2, :])
    assert "n_steps" in ch.extras and "is_accept" in ch and "θ̃" in ch
    assert ch.last_sampler_state["seed"] == 1
    assert ch.to_chains().shape == v.shape
    assert ch.mean("θ̃").shape == (3,)


def test_params_with_stats_overload():
    """src/mcmc/Inference.jl:88-101: parameters from the transition as (string, value) pairs; stats only
    on request; extras always empty"""
    import turing_jl_b200 as T
    tr = T.ParamsWithStats({"s": 2.0, "m": -1.0}, {"logjoint": -3.0, "n_steps": 7.0})
    p, s, e = T.params_with_stats(None, None, tr, None)
    assert p == [("s", 2.0), ("m", -1.0)] and s == {} and e == {}
    p, s, e = T.params_with_stats(None, None, tr, None, params=False, stats=True, extras=True)
    assert p is None and s["n_steps"] == 7.0 and e == {}
