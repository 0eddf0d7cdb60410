"""ctypes binding of oracle/libtnuts_oracle.so (test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtnuts_oracle.so")

GDEMO, LINREG, EIGHT_SCHOOLS, LOGISTIC, HIER_LINREG, STD_NORMAL, BETA_BERNOULLI, DIRICHLET_CATEGORICAL = range(8)
NSTAT = 9
STAT_NAMES = ("n_steps", "acceptance_rate", "log_density", "hamiltonian_energy",
              "hamiltonian_energy_error", "max_hamiltonian_energy_error", "tree_depth",
              "numerical_error", "step_size")
SLOT_INIT, SLOT_MOMENTUM, SLOT_DIRECTION, SLOT_MH, SLOT_LEAF, SLOT_MERGE, SLOT_EPS_MOMENTUM = range(7)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


class _Model(C.Structure):
    _fields_ = [("family", C.c_int), ("D", C.c_int), ("N", C.c_int64), ("G", C.c_int),
                ("X", _dp), ("y", _dp), ("sigma", _dp), ("group", _ip), ("hyper", C.c_double * 8)]


class _Config(C.Structure):
    _fields_ = [("delta", C.c_double), ("max_depth", C.c_int), ("delta_max", C.c_double),
                ("init_eps", C.c_double), ("n_adapts", C.c_int64), ("sampler_mode", C.c_int),
                ("seed", C.c_uint64), ("algorithm", C.c_int), ("n_leapfrog", C.c_int),
                ("lambda_", C.c_double), ("metric", C.c_int)]


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "tnuts_oracle.c")
    hdr = os.path.join(_HERE, "tnuts_oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libtnuts_oracle.so"],
                          stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_logdensity_and_gradient.restype = C.c_double
        L.orc_logdensity_and_gradient.argtypes = [C.POINTER(_Model), _dp, _dp]
        L.orc_num_params.restype = C.c_int
        L.orc_num_params.argtypes = [C.POINTER(_Model)]
        L.orc_constrain_and_logp.restype = None
        L.orc_constrain_and_logp.argtypes = [C.POINTER(_Model), _dp, _dp, _dp]
        L.orc_philox4x32_10.restype = None
        L.orc_philox4x32_10.argtypes = [C.c_uint64] + [C.c_uint32] * 4 + [C.POINTER(C.c_uint32)]
        L.orc_uniform2.restype = None
        L.orc_uniform2.argtypes = [C.c_uint64] + [C.c_uint32] * 4 + [_dp]
        L.orc_normal2.restype = None
        L.orc_normal2.argtypes = [C.c_uint64] + [C.c_uint32] * 4 + [_dp]
        L.orc_sample_chains.restype = C.c_int64
        L.orc_sample_chains.argtypes = [C.POINTER(_Model), C.POINTER(_Config), C.c_int, C.c_int,
                                        _dp, C.c_int64, C.c_int64, C.c_int64, _dp, _dp, _dp, _dp,
                                        _dp, _dp]
        L.orc_sample_chain_ids.restype = C.c_int64
        L.orc_sample_chain_ids.argtypes = [C.POINTER(_Model), C.POINTER(_Config), C.c_int, C.c_int,
                                           C.POINTER(C.c_uint32), _dp, C.c_int64, C.c_int64, C.c_int64,
                                           _dp, _dp, _dp, _dp, _dp, _dp]
        L.orc_resume_chains.restype = C.c_int64
        L.orc_resume_chains.argtypes = [C.POINTER(_Model), C.POINTER(_Config), C.c_int, C.c_int,
                                        C.c_uint32, _dp, _dp, _dp, C.c_int64, C.c_int64, C.c_int64,
                                        C.c_int64, _dp, _dp, _dp, _dp, _dp]
        L.orc_leapfrog_trajectory.restype = None
        L.orc_leapfrog_trajectory.argtypes = [C.POINTER(_Model), _dp, _dp, _dp, C.c_double, C.c_int,
                                              _dp, _dp, _dp]
        L.orc_window_splits.restype = C.c_int
        L.orc_window_splits.argtypes = [C.c_int64, _lp, _lp, _lp, C.c_int]
        L.orc_find_good_stepsize.restype = C.c_double
        L.orc_find_good_stepsize.argtypes = [C.POINTER(_Model), C.POINTER(_Config), C.c_uint32,
                                             _dp, _dp, _lp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


class OracleModel:
    """family id + data arrays + prior hyper-parameters."""

    def __init__(self, family, D, N=0, G=0, X=None, y=None, sigma=None, group=None, hyper=()):
        self.family, self.D, self.N, self.G = int(family), int(D), int(N), int(G)
        self.X = None if X is None else np.ascontiguousarray(X, dtype=np.float64)
        self.y = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        self.sigma = None if sigma is None else np.ascontiguousarray(sigma, dtype=np.float64)
        self.group = None if group is None else np.ascontiguousarray(group, dtype=np.int32)
        self.hyper = tuple(float(h) for h in hyper)
        m = _Model()
        m.family, m.D, m.N, m.G = self.family, self.D, self.N, self.G
        m.X, m.y, m.sigma = _d(self.X), _d(self.y), _d(self.sigma)
        m.group = self.group.ctypes.data_as(_ip) if self.group is not None else None
        for i, h in enumerate(self.hyper):
            m.hyper[i] = h
        self._c = m

    def lpgrad(self, theta):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        if theta.ndim == 1:
            g = np.empty(self.D)
            lp = lib().orc_logdensity_and_gradient(C.byref(self._c), _d(theta), _d(g))
            return lp, g
        lps = np.empty(theta.shape[0])
        gs = np.empty_like(theta)
        for i in range(theta.shape[0]):
            lps[i] = lib().orc_logdensity_and_gradient(C.byref(self._c), _d(theta[i]), _d(gs[i]))
        return lps, gs

    @property
    def P(self):
        """user-space parameters (D, except K per K-simplex)"""
        return lib().orc_num_params(C.byref(self._c))

    def constrain(self, theta):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        p = np.empty(self.P)
        lp3 = np.empty(3)
        lib().orc_constrain_and_logp(C.byref(self._c), _d(theta), _d(p), _d(lp3))
        return p, lp3


def make_config(delta=0.65, max_depth=10, delta_max=1000.0, init_eps=0.0, n_adapts=0,
                sampler_mode=1, seed=0, algorithm=0, n_leapfrog=0, lambda_=0.0, metric=0):
    c = _Config()
    c.delta, c.max_depth, c.delta_max, c.init_eps = delta, max_depth, delta_max, init_eps
    c.n_adapts, c.sampler_mode, c.seed = n_adapts, sampler_mode, seed
    c.algorithm, c.n_leapfrog, c.lambda_ = algorithm, n_leapfrog, lambda_
    c.metric = metric
    return c


def philox(seed, c0, c1, c2, c3):
    out = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(seed, c0, c1, c2, c3, out)
    return tuple(out)


def uniform2(seed, chain, it, slot, sub):
    u = np.empty(2)
    lib().orc_uniform2(seed, chain, it, slot, sub, _d(u))
    return u


def normal2(seed, chain, it, slot, sub):
    u = np.empty(2)
    lib().orc_normal2(seed, chain, it, slot, sub, _d(u))
    return u


def sample_chains(model, cfg, n_chains, n_transitions, first_keep=1, thin=1, theta0=None,
                  n_threads=0, chain_ids=None):
    """Returns dict(theta[C,K,D], params[C,K,P], logp[C,K,3], stats[C,K,NSTAT], eps[C], minv[C,D],
    n_grad).  chain_ids: global ids of the chains to run (default 0..n_chains-1)."""
    ids = None
    if chain_ids is not None:
        ids = np.ascontiguousarray(chain_ids, dtype=np.uint32)
        n_chains = ids.size
    D = model.D
    nkeep = (n_transitions - first_keep) // thin + 1 if n_transitions >= first_keep >= 1 else 0
    th = np.zeros((n_chains, nkeep, D))
    pa = np.zeros((n_chains, nkeep, model.P))
    lp = np.zeros((n_chains, nkeep, 3))
    st = np.zeros((n_chains, nkeep, NSTAT))
    eps = np.zeros(n_chains)
    minv = np.zeros((n_chains, D))
    t0 = None if theta0 is None else np.ascontiguousarray(theta0, dtype=np.float64)
    n = lib().orc_sample_chain_ids(C.byref(model._c), C.byref(cfg), n_chains, n_threads,
                                   ids.ctypes.data_as(C.POINTER(C.c_uint32)) if ids is not None else None,
                                   _d(t0), n_transitions, first_keep, thin, _d(th), _d(pa), _d(lp),
                                   _d(st), _d(eps), _d(minv))
    if n < 0:
        raise RuntimeError("oracle: failed to find valid initial parameters")
    return dict(theta=th, params=pa, logp=lp, stats=st, eps=eps, minv=minv, n_grad=int(n))


def resume_chains(model, cfg, theta, eps, minv, iter0, n_transitions, first_keep=1, thin=1,
                  n_threads=0, chain_id0=0):
    """continue adapted chains: theta [C,D], eps [C], minv [C,D]; returns the same dict as
    sample_chains plus final_theta [C,D]"""
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    eps = np.ascontiguousarray(eps, dtype=np.float64)
    minv = np.ascontiguousarray(minv, dtype=np.float64)
    n_chains, D = theta.shape
    nkeep = (n_transitions - first_keep) // thin + 1 if n_transitions >= first_keep >= 1 else 0
    th = np.zeros((n_chains, nkeep, D))
    pa = np.zeros((n_chains, nkeep, model.P))
    lp = np.zeros((n_chains, nkeep, 3))
    st = np.zeros((n_chains, nkeep, NSTAT))
    fin = np.zeros((n_chains, D))
    n = lib().orc_resume_chains(C.byref(model._c), C.byref(cfg), n_chains, n_threads, chain_id0,
                                _d(theta), _d(eps), _d(minv), iter0, n_transitions, first_keep, thin,
                                _d(th), _d(pa), _d(lp), _d(st), _d(fin))
    return dict(theta=th, params=pa, logp=lp, stats=st, final_theta=fin, n_grad=int(n))


def leapfrog_trajectory(model, theta, r, minv, eps, L):
    D = model.D
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    r = np.ascontiguousarray(r, dtype=np.float64)
    minv = np.ascontiguousarray(minv, dtype=np.float64)
    tt = np.empty((L + 1, D))
    tr = np.empty((L + 1, D))
    tH = np.empty(L + 1)
    lib().orc_leapfrog_trajectory(C.byref(model._c), _d(theta), _d(r), _d(minv), eps, L, _d(tt),
                                  _d(tr), _d(tH))
    return tt, tr, tH


def window_splits(n_adapts):
    ws, we = C.c_int64(), C.c_int64()
    sp = (C.c_int64 * 64)()
    k = lib().orc_window_splits(n_adapts, C.byref(ws), C.byref(we), sp, 64)
    return ws.value, we.value, [sp[i] for i in range(k)]


def find_good_stepsize(model, cfg, chain, theta, minv):
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    minv = np.ascontiguousarray(minv, dtype=np.float64)
    ng = C.c_int64(0)
    return lib().orc_find_good_stepsize(C.byref(model._c), C.byref(cfg), chain, _d(theta), _d(minv),
                                        C.byref(ng))
