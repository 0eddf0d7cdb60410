"""Thin object wrapper over the C ABI handle (one engine == one GPU)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class Engine:
    def __init__(self, model, spl, n_chains, n_adapts, seed, device=0, chain_offset=0,
                 strategy=0, row_shards=1, row_rank=0):
        self.L = _lib.load()
        self.model = model
        cfg = _lib.TnutsConfig()
        cfg.algorithm = spl.algorithm
        cfg.delta = spl.delta
        cfg.max_depth = spl.max_depth
        cfg.delta_max = spl.Delta_max
        cfg.init_eps = spl.eps
        cfg.n_adapts = int(n_adapts)
        cfg.n_leapfrog = getattr(spl, "n_leapfrog", 0)
        cfg.lambda_ = getattr(spl, "lam", 0.0)
        cfg.metric = getattr(spl, "metric", 0)
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.n_chains = int(n_chains)
        cfg.chain_offset = int(chain_offset)
        cfg.device = int(device)
        cfg.row_shards, cfg.row_rank, cfg.strategy = int(row_shards), int(row_rank), int(strategy)
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.L.tnuts_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise _lib.TnutsError(rc, (self.L.tnuts_last_error(None) or b"").decode())
        self._cm = model.c_struct()
        try:
            if getattr(model, "on_device", False):
                check(self.h, self.L.tnuts_set_model_device(self.h, C.byref(self._cm)))
            else:
                check(self.h, self.L.tnuts_set_model(self.h, C.byref(self._cm)))
        except Exception:
            self.close()
            raise
        self.C, self.D, self.P = int(n_chains), model.D, self.L.tnuts_num_params(self.h)
        self.n_kept = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.tnuts_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- sampling ---
    def init(self, theta0=None):
        t0 = None if theta0 is None else np.ascontiguousarray(theta0, dtype=np.float64)
        if t0 is not None and t0.shape != (self.C, self.D):
            raise ValueError("theta0 must be [n_chains, D]")
        check(self.h, self.L.tnuts_init(self.h, _dp(t0)))

    def run(self, n_transitions, first_keep=1, thin=1):
        k = C.c_int64(0)
        check(self.h, self.L.tnuts_run(self.h, int(n_transitions), int(first_keep), int(thin),
                                       C.byref(k)))
        self.n_kept = k.value
        return self.n_kept

    def run_streamed(self, n_transitions, first_keep, thin, out, chunk_draws=100):
        """run + overlapped D2H of the kept draws straight into `out` [n_kept, ncol, C]"""
        k = C.c_int64(0)
        assert out.flags.c_contiguous and out.dtype == np.float64
        check(self.h, self.L.tnuts_run_streamed(self.h, int(n_transitions), int(first_keep),
                                                int(thin), _dp(out), int(chunk_draws), C.byref(k)))
        self.n_kept = k.value
        assert out.shape[0] >= self.n_kept
        return self.n_kept

    def run_streamed_slab(self, n_transitions, first_keep, thin, out, lo, chunk_draws=100):
        """run + overlapped D2H into chains [lo, lo + C) of a wider chain array `out`
        [n_kept, ncol, n_chains_total] (several engines, one per GPU, fill one array)"""
        k = C.c_int64(0)
        assert out.flags.c_contiguous and out.dtype == np.float64 and out.ndim == 3
        assert out.shape[1] == self.ncol and 0 <= lo and lo + self.C <= out.shape[2]
        base = C.cast(out.ctypes.data + 8 * lo, C.POINTER(C.c_double))
        check(self.h, self.L.tnuts_run_streamed_strided(self.h, int(n_transitions), int(first_keep), int(thin),
                                                        base, int(out.shape[2]), int(chunk_draws), C.byref(k)))
        self.n_kept = k.value
        assert out.shape[0] >= self.n_kept
        return self.n_kept

    def set_option(self, name, value):
        check(self.h, self.L.tnuts_set_option(self.h, name.encode(), float(value)))

    @property
    def ncol(self):
        return self.L.tnuts_num_columns(self.h)

    def fetch_draws(self, out=None):
        """the whole kept array [n_kept, ncol, n_chains] in one contiguous copy"""
        if out is None:
            out = np.empty((self.n_kept, self.ncol, self.C))
        assert out.shape == (self.n_kept, self.ncol, self.C) and out.flags.c_contiguous
        if self.n_kept:
            check(self.h, self.L.tnuts_fetch_draws(self.h, _dp(out)))
        return out

    def fetch(self, params=True, logp=True, stats=True, theta=False):
        K, Cn = self.n_kept, self.C
        out = {}
        bufs = {}
        for name, want, w in (("params", params, self.P), ("logp", logp, 3),
                              ("stats", stats, _lib.NSTAT), ("theta", theta, self.D)):
            bufs[name] = np.empty((K, w, Cn)) if want else None
        check(self.h, self.L.tnuts_fetch(self.h, _dp(bufs["params"]), _dp(bufs["logp"]),
                                         _dp(bufs["stats"]), _dp(bufs["theta"])))
        for k_, v in bufs.items():
            if v is not None:
                out[k_] = v
        return out

    def theta(self):
        t = np.empty((self.C, self.D))
        check(self.h, self.L.tnuts_get_theta(self.h, _dp(t)))
        return t

    def stepsize(self):
        e = np.empty(self.C)
        check(self.h, self.L.tnuts_get_stepsize(self.h, _dp(e)))
        return e

    def inv_metric(self):
        m = np.empty((self.C, self.D))
        check(self.h, self.L.tnuts_get_inv_metric(self.h, _dp(m)))
        return m

    def dense_inv_metric(self):
        """[n_chains, D, D]: the full M^-1 of a DenseEuclideanMetric engine"""
        m = np.empty((self.C, self.D, self.D))
        check(self.h, self.L.tnuts_get_dense_inv_metric(self.h, _dp(m)))
        return m

    def constrain(self, theta):
        th = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        n = th.shape[0]
        p = np.empty((n, self.P))
        lp = np.empty((n, 3))
        check(self.h, self.L.tnuts_constrain(self.h, _dp(th), n, _dp(p), _dp(lp)))
        return p, lp

    # --- LogDensityProblems interface ---
    def logdensity_and_gradient(self, theta):
        th = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        n = th.shape[0]
        lp = np.empty(n)
        g = np.empty((n, self.D))
        check(self.h, self.L.tnuts_logdensity_and_gradient(self.h, _dp(th), n, _dp(lp), _dp(g)))
        return lp, g

    def logdensity(self, theta):
        th = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        lp = np.empty(th.shape[0])
        check(self.h, self.L.tnuts_logdensity(self.h, _dp(th), th.shape[0], _dp(lp)))
        return lp

    def leapfrog(self, theta, r, minv, eps, L):
        th = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        rr = np.ascontiguousarray(np.atleast_2d(r), dtype=np.float64)
        mi = np.ascontiguousarray(minv, dtype=np.float64)
        n = th.shape[0]
        tt = np.empty((n, L + 1, self.D))
        tr = np.empty((n, L + 1, self.D))
        tH = np.empty((n, L + 1))
        check(self.h, self.L.tnuts_leapfrog(self.h, _dp(th), _dp(rr), _dp(mi), float(eps), int(L),
                                            n, _dp(tt), _dp(tr), _dp(tH)))
        return tt, tr, tH

    # --- state blob ---
    def get_state(self):
        n = C.c_size_t(0)
        check(self.h, self.L.tnuts_get_state(self.h, None, C.byref(n)))
        buf = np.empty(n.value, dtype=np.uint8)
        check(self.h, self.L.tnuts_get_state(self.h, buf.ctypes.data_as(C.c_void_p), C.byref(n)))
        return buf

    def set_state(self, blob):
        b = np.ascontiguousarray(blob, dtype=np.uint8)
        check(self.h, self.L.tnuts_set_state(self.h, b.ctypes.data_as(C.c_void_p), b.size))

    # --- diagnostics / counters ---
    def diagnostics(self):
        r = np.empty(self.P)
        e = np.empty(self.P)
        check(self.h, self.L.tnuts_diagnostics(self.h, _dp(r), _dp(e)))
        return r, e

    def bench_gradient(self, reps=5):
        s = C.c_double(0.0)
        check(self.h, self.L.tnuts_bench_gradient(self.h, int(reps), C.byref(s)))
        return s.value / reps

    def kernel_times(self):
        """(seconds, launches, sampled, tile_launches, persist_seconds, persist_pumps): the gradient
        kernel's launches and the persistent tail kernel inside the last run (option "time_kernels")"""
        s, n, k, t = C.c_double(0.0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
        ps, pp = C.c_double(0.0), C.c_int64(0)
        check(self.h, self.L.tnuts_kernel_times(self.h, C.byref(s), C.byref(n), C.byref(k), C.byref(t),
                                                C.byref(ps), C.byref(pp)))
        return s.value, n.value, k.value, t.value, ps.value, pp.value

    def comm_init(self, unique_id, world, rank):
        check(self.h, self.L.tnuts_comm_init(self.h, unique_id, world, rank))

    @property
    def grad_evals(self):
        return self.L.tnuts_num_grad_evals(self.h)

    @property
    def divergences(self):
        return self.L.tnuts_num_divergences(self.h)

    @property
    def launches(self):
        return self.L.tnuts_num_kernel_launches(self.h)

    @property
    def device_seconds(self):
        """CUDA-event seconds of the last init + the last run"""
        return self.L.tnuts_device_seconds(self.h) + self.L.tnuts_init_device_seconds(self.h)

    @property
    def init_tries_max(self):
        return self.L.tnuts_init_tries_max(self.h)
