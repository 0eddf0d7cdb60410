"""`sample(model, NUTS(delta), MCMCThreads(), N, n_chains)` -- host mirror of the reference's
entry points, driving the CUDA engine through the C ABI.

Mirrors, with the same names, argument meaning and error behaviour:
  sample(rng, model, ::AdaptiveHamiltonian, N; ...)            src/mcmc/hmc.jl:85-154
  sample(rng, model, spl, ensemble, N, n_chains; ...)          src/mcmc/abstractmcmc.jl:133-166
  AbstractMCMC.mcmcsample loop (discard_initial, thinning)     SURVEY.md App. C
  post_sample_hook (divergence warning)                        src/mcmc/hmc.jl:476-486
  loadstate / initial_state resume                             src/mcmc/hmc.jl:135-152
The difference: all chains advance inside one engine call instead of one task per chain.
"""
import ctypes as C
import threading
import time
import warnings

import numpy as np

from . import _lib, pinned
from .chains import INTERNALS, Chains, ParamsWithStats
from .engine import Engine
from .samplers import (Hamiltonian, InitFromParams, InitFromUniform, MCMCThreads)

_DEFAULT_RNG = np.random.default_rng()


def _check_model(model, spl):
    """Turing._check_model (src/common.jl:31-46): HMC-family samplers reject discrete variables."""
    if isinstance(spl, Hamiltonian) and model.has_discrete_variables():
        raise ValueError("Discrete variables are not supported by HMC/NUTS samplers")


def _seed_from(rng, seed):
    if seed is not None:
        return int(seed)
    if rng is None:
        rng = _DEFAULT_RNG
    if isinstance(rng, (int, np.integer)):
        return int(rng)
    return int(rng.integers(0, 2 ** 63 - 1))


def _resolve_initial_params(model, initial_params, n_chains):
    """-> [n_chains, D] linked-space array, or None for InitFromUniform."""
    if initial_params is None or isinstance(initial_params, InitFromUniform):
        return None
    if not isinstance(initial_params, (list, tuple)) or len(initial_params) != n_chains:
        raise ValueError("`initial_params` must be an AbstractVector of length `n_chains`; "
                         "one element per chain")
    if all(ip is None or isinstance(ip, InitFromUniform) for ip in initial_params):
        return None
    rows = []
    for ip in initial_params:
        if isinstance(ip, InitFromParams):
            ip = ip.params
        if isinstance(ip, dict):
            vec = []
            for n in model.param_names:
                base = n.split("[")[0]
                if n in ip:
                    vec.append(ip[n])
                elif base in ip:
                    vec.append(np.ravel(ip[base])[int(n.split("[")[1][:-1]) - 1])
                else:
                    raise ValueError("initial_params is missing a value for %s" % n)
            ip = vec
        if ip is None or isinstance(ip, InitFromUniform):
            raise ValueError("mixing InitFromUniform with explicit initial_params is not supported")
        rows.append(model.link(np.asarray(ip, dtype=np.float64)))
    return np.stack(rows)


def sample(model, spl, *args, rng=None, seed=None, chain_type=Chains, initial_params=None,
           initial_state=None, nadapts=None, discard_adapt=True, discard_initial=-1, thinning=1,
           save_state=False, progress=False, verbose=True, check_model=True, devices=None,
           diagnostics=False, strategy=0, keep_engine=False, chain_offset=0, chunk_draws=None,
           engine_options=None, callback=None):
    """sample(model, spl, N) | sample(model, spl, ensemble, N, n_chains).

    engine_options: {name: value} passed to tnuts_set_option before init, e.g.
    {"logistic_tc": 1} (tcgen05 likelihood for logistic regression, include/tnuts.h).
    callback: `callback(rng, model, sampler, transition, state, iteration)` after every kept iteration
    (AbstractMCMC.mcmcsample's `callback` keyword); the engine then runs one kept draw per call."""
    if len(args) == 1:
        N, n_chains, single = int(args[0]), 1, True
        if initial_params is not None and not isinstance(initial_params, (list, tuple)):
            initial_params = [initial_params]
        elif initial_params is not None and not isinstance(initial_params[0], (InitFromParams, dict, list, tuple, np.ndarray)):
            initial_params = [initial_params]
    elif len(args) == 3:
        if not isinstance(args[0], MCMCThreads):
            raise TypeError("sample(model, spl, ensemble, N, n_chains): bad ensemble")
        N, n_chains, single = int(args[1]), int(args[2]), False
    else:
        raise TypeError("sample(model, spl, N) or sample(model, spl, ensemble, N, n_chains)")
    if check_model:
        _check_model(model, spl)
    if N < 1:
        raise ValueError("N must be >= 1")

    # --- src/mcmc/hmc.jl:104-152 ---
    if initial_state is None:
        na = spl.n_adapts if nadapts is None else nadapts
        if spl.adaptive:
            _nadapts = min(1000, N // 2) if na == -1 else int(na)
        else:
            _nadapts = 0
        if spl.adaptive and discard_initial == -1:
            _discard = _nadapts if discard_adapt else 0
        else:
            _discard = max(0, discard_initial)
    else:
        _nadapts, _discard = 0, 0
    theta0 = _resolve_initial_params(model, initial_params, n_chains) if initial_state is None else None
    seed64 = _seed_from(rng, seed)
    if initial_state is not None:
        seed64 = initial_state["seed"]

    # AHMCAdaptor (src/mcmc/hmc.jl:447-465): `iszero(alg.n_adapts)` gives NoAdaptation whatever
    # `nadapts` the call passes; the discard schedule above still follows the keyword
    _engine_nadapts = 0 if (spl.adaptive and spl.n_adapts == 0) else _nadapts
    devices = list(devices) if devices else [0]
    ndev = len(devices)
    if n_chains < ndev:
        devices, ndev = devices[:n_chains], n_chains
    bounds = [n_chains * i // ndev for i in range(ndev + 1)]

    if chunk_draws is None:
        # the kept draws leave in segments while the next one computes; after warm-up the copy is the
        # slower of the two (PCIe), so small segments only matter for how early the first copy starts
        chunk_draws = max(1, -(-N // 40))
    t_start = time.perf_counter()
    tm = {}
    engines = []
    for i, dev in enumerate(devices):
        engines.append(Engine(model, spl, bounds[i + 1] - bounds[i], _engine_nadapts, seed64, device=dev,
                              chain_offset=chain_offset + bounds[i], strategy=strategy))
        for k, v in (engine_options or {}).items():
            engines[-1].set_option(k, v)
    tm["create"] = time.perf_counter() - t_start
    results = [None] * ndev
    errors = [None] * ndev
    # bundle_samples target: (iterations x names x chains); the engine's device layout is
    # identical, so a single-GPU run copies straight into it (pinned host memory)
    P = model.P          # chain columns: D, except K per K-simplex
    nint = len(INTERNALS)
    value = pinned.empty((N, P + nint, n_chains))
    if _discard == 0 and initial_state is None:
        value[0, P + 3:, :] = np.nan
    tm["alloc"] = time.perf_counter() - t_start - tm["create"]

    uid = None
    ok = [False] * ndev
    gate = threading.Barrier(ndev)
    if diagnostics and ndev > 1:
        from .distributed import _new_unique_id
        uid = _new_unique_id()

    def work(i):
        try:
            e = engines[i]
            lo, hi = bounds[i], bounds[i + 1]
            first = None
            if uid is not None:
                # one in-process NCCL communicator over the engines (joined before anything can fail)
                e.comm_init(uid.ctypes.data_as(C.c_void_p), ndev, i)
            if initial_state is not None:
                e.set_state(initial_state["blobs"][i])
                # mcmcsample keeps the very first step from the restored state, then thins
                n_tr, first_keep = 1 + (N - 1) * thinning, 1
            else:
                try:
                    e.init(None if theta0 is None else theta0[lo:hi])
                finally:
                    if theta0 is None and e.init_tries_max >= 10:
                        # src/mcmc/abstractmcmc.jl:62-63 (emitted at the tenth failed attempt)
                        warnings.warn("failed to find valid initial parameters in 10 tries; consider "
                                      "providing a different initialisation strategy with the "
                                      "`initial_params` keyword")
                if _discard == 0:
                    # first sample = initial parameters, no sampler stats (test/mcmc/hmc.jl:173-197)
                    th = e.theta()
                    p, lp = e.constrain(th)
                    first = (p, lp)
                    n_tr, first_keep = (N - 1) * thinning, thinning
                else:
                    n_tr, first_keep = _discard + (N - 1) * thinning, _discard
            row = 0 if first is None else 1
            if first is not None:
                value[0, :P, lo:hi] = first[0].T
                value[0, P:P + 3, lo:hi] = first[1].T
            if n_tr > 0 and callback is not None:
                # mcmcsample calls the callback after every kept iteration: one kept draw per engine call
                if first is not None:
                    callback(rng, model, spl, transition_of(model, value[0, :, lo:hi], single), HMCState(e, 0), 1)
                it_done = 0
                for k in range(N - row):
                    n_k = first_keep if k == 0 else thinning
                    e.run(n_k, n_k, 1)
                    it_done += n_k
                    value[row + k, :, lo:hi] = e.fetch_draws()[0]
                    callback(rng, model, spl, transition_of(model, value[row + k, :, lo:hi], single),
                             HMCState(e, it_done), row + k + 1)
                e.n_kept = 0    # the engine's draw buffer holds the last draw only: no device diagnostics
            elif n_tr > 0:
                if ndev == 1:
                    # straight into the chain array, D2H overlapped with the computation
                    e.run_streamed(n_tr, first_keep, thinning, value[row:], chunk_draws)
                else:
                    # this GPU's chains are a slab [lo, hi) of the chain axis: strided D2H copies,
                    # overlapped with the computation like the single-GPU path, no host reshuffle
                    e.run_streamed_slab(n_tr, first_keep, thinning, value[row:], lo, chunk_draws)
            else:
                e.n_kept = 0
            pooled = True
            if uid is not None:
                # pooled over every GPU's chains (a collective): only if every engine got this far
                ok[i] = True
                gate.wait()
                pooled = all(ok)
            diag = e.diagnostics() if (diagnostics and pooled and e.n_kept >= 4) else None
            ndiv = e.divergences if e.n_kept else 0
            results[i] = (diag, e.grad_evals, e.device_seconds, e.stepsize(),
                          e.get_state() if save_state else None, e.launches, ndiv)
        except Exception as ex:  # re-raised on the caller's thread
            errors[i] = ex
            if uid is not None and not ok[i]:
                try:
                    gate.wait(timeout=600)
                except threading.BrokenBarrierError:
                    pass

    if ndev == 1:
        work(0)
    else:
        ths = [threading.Thread(target=work, args=(i,)) for i in range(ndev)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    for ex in errors:
        if ex is not None:
            for e in engines:
                e.close()
            if isinstance(ex, _lib.TnutsError) and ex.code == _lib.E_INIT:
                raise RuntimeError(ex.message) from None
            raise ex
    wall = time.perf_counter() - t_start
    tm["work"] = wall - tm["create"] - tm["alloc"]

    info = {
        "sampling_time": np.full(n_chains, wall),
        "wall_seconds": wall,
        "device_seconds": max(r[2] for r in results),
        "grad_evals": int(sum(r[1] for r in results)),
        "gpu_launches": int(sum(r[5] for r in results)),
        "step_size": np.concatenate([r[3] for r in results]),
        "algorithm": spl.algorithm, "host_timing": tm,
        "n_divergent": int(sum(r[6] for r in results)),
        "nadapts": _nadapts, "discard_initial": _discard, "thinning": thinning, "seed": seed64,
    }
    if results[0][0] is not None:
        info["rhat"], info["ess_bulk"] = results[0][0]
    if save_state:
        info["samplerstate"] = {"blobs": [r[4] for r in results], "seed": seed64,
                                "bounds": bounds, "devices": devices}
    chain = chain_type(value, model.param_names, INTERNALS, info)
    t_c = time.perf_counter()
    if keep_engine:
        chain.info["engines"] = engines
    else:
        for e in engines:
            e.close()
    tm["close"] = time.perf_counter() - t_c
    post_sample_hook(chain, spl, verbose=verbose)
    return chain


class HMCState:
    """`HMCState` (src/mcmc/hmc.jl:1-12: i, kernel, hamiltonian, z, adaptor).  Here the engine holds
    all of it on the device (position, gradient, step size, metric, dual-averaging and Welford state);
    `i` is the iteration counter, `blob()` the serialised state (`save_state` / `initial_state`)."""

    def __init__(self, engine, i):
        self.engine, self.i = engine, int(i)

    def blob(self):
        return self.engine.get_state()

    def stepsize(self):
        return self.engine.stepsize()


def transition_of(model, cols, single):
    """one kept draw ([ncol][chains] columns of the chain array) as a ParamsWithStats in user space"""
    P = len(model.param_names)
    pick = (lambda v: float(v[0])) if single else (lambda v: np.array(v))
    params = {n: pick(cols[i]) for i, n in enumerate(model.param_names)}
    stats = {n: pick(cols[P + i]) for i, n in enumerate(INTERNALS)}
    return ParamsWithStats(params, stats)


def step(rng, model, spl, state=None, *, initial_params=None, nadapts=0, n_chains=1, seed=None, device=0,
         strategy=0, engine_options=None):
    """`AbstractMCMC.step` for the Hamiltonian samplers on an engine of `n_chains` chains (default one).
    step(rng, model, spl; initial_params, nadapts)  -- src/mcmc/hmc.jl:156-207: initial parameters (InitFromUniform
        with <= 1000 tries, or the given ones), step-size search when `spl.eps == 0`, adaptor; the
        transition is the initial point (logp columns, no sampler stats), the state has i = 0
    step(rng, model, spl, state; nadapts)           -- src/mcmc/hmc.jl:209-264: one transition + adaptation
        while i <= nadapts; the transition carries the AdvancedHMC statistics.
    Returns (ParamsWithStats, HMCState)."""
    single = n_chains == 1
    P = len(model.param_names)
    if state is None:
        _check_model(model, spl)
        na = int(nadapts) if spl.adaptive and spl.n_adapts != 0 else 0
        e = Engine(model, spl, n_chains, na, _seed_from(rng, seed), device=device, strategy=strategy)
        for k, v in (engine_options or {}).items():
            e.set_option(k, v)
        if initial_params is not None and not isinstance(initial_params, (list, tuple)):
            initial_params = [initial_params] * n_chains
        theta0 = _resolve_initial_params(model, initial_params, n_chains)
        try:
            e.init(theta0)
        except _lib.TnutsError as ex:
            e.close()
            if ex.code == _lib.E_INIT:
                raise RuntimeError(ex.message) from None
            raise
        p, lp = e.constrain(e.theta())
        cols = np.full((P + len(INTERNALS), n_chains), np.nan)
        cols[:P] = p.T
        cols[P:P + 3] = lp.T
        return transition_of(model, cols, single), HMCState(e, 0)
    e = state.engine
    e.run(1, 1, 1)
    return transition_of(model, e.fetch_draws()[0], single), HMCState(e, state.i + 1)


def post_sample_hook(chain, spl, verbose=True):
    """src/mcmc/hmc.jl:476-486"""
    n_div = chain.info.get("n_divergent")
    if n_div is None:  # chains not produced by the engine: count on the host like the reference
        n_div = int(np.nansum(chain["numerical_error"]))
    if verbose and n_div > 0:
        warnings.warn("There were %d divergent transitions. Consider reparameterising your model or "
                      "using a smaller step size. For adaptive samplers such as NUTS and HMCDA, "
                      "consider increasing `target_accept`." % n_div)


def loadstate(chain):
    """Turing.Inference.loadstate (src/mcmc/Inference.jl:75-77)"""
    st = chain.info.get("samplerstate")
    if st is None:
        raise ValueError("the chain object does not contain the final state of the sampler; "
                         "sample with `save_state=true`")
    return st


class LogDensityFunction:
    """Engine-backed object with the LogDensityProblems interface (dimension, logdensity,
    logdensity_and_gradient) so the host can cross-check it against
    `DynamicPPL.LogDensityFunction(model, getlogjoint_internal, LinkAll())`
    (src/mcmc/hmc.jl:170-182)."""

    def __init__(self, model, device=0, strategy=0, options=None):
        from .samplers import NUTS
        self.model = model
        self._e = Engine(model, NUTS(), 1, 0, 0, device=device, strategy=strategy)
        for k, v in (options or {}).items():
            self._e.set_option(k, v)

    def dimension(self):
        return self._e.D

    def logdensity(self, theta):
        lp = self._e.logdensity(theta)
        return lp[0] if np.ndim(theta) == 1 else lp

    def logdensity_and_gradient(self, theta):
        lp, g = self._e.logdensity_and_gradient(theta)
        if np.ndim(theta) == 1:
            return lp[0], g[0]
        return lp, g

    def capabilities(self):
        return 1  # LogDensityOrder{1}: gradient available
