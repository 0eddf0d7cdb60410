"""Registered model families: the host-side descriptor the engine needs instead of Turing's
opaque `DynamicPPL.Model` closure (SURVEY.md section 7.3 "black-box model vs registered family").

Each class mirrors a Turing `@model` (quoted in its docstring), carries the observed data,
names its parameters in model statement order (vector parameters flattened as `x[1], x[2]`,
test/mcmc/chains.jl:233-251) and knows the link transform (`Bijectors`: positive support
links by log) so that user-space `initial_params` can be mapped to linked space.
"""
import ctypes as C

import numpy as np

from . import _lib

GDEMO, LINREG, EIGHT_SCHOOLS, LOGISTIC, HIER_LINREG, STD_NORMAL, BETA_BERNOULLI, DIRICHLET_CATEGORICAL = range(8)


class Model:
    family = -1
    #: names of parameters whose support is (0, inf) -> linked by log
    positive = ()

    def __init__(self):
        self.X = self.y = self.sigma = self.group = None
        self.N = 0
        self.G = 0
        self.hyper = ()

    # --- interface used by sample() ---
    @property
    def D(self):
        """LogDensityProblems.dimension: length of the linked (unconstrained) vector"""
        return len(self.param_names)

    @property
    def P(self):
        """number of user-space parameters (columns of the chain): D, except K per K-simplex"""
        return len(self.param_names)

    def has_discrete_variables(self):
        """Turing._check_model (src/common.jl:31-46) rejects discrete variables for HMC."""
        return False

    def link(self, params):
        """user-space (constrained) vector -> linked space."""
        p = np.array(params, dtype=np.float64, copy=True)
        for i, n in enumerate(self.param_names):
            if n in self.positive:
                if not p[..., i].min() > 0:
                    raise ValueError("initial value of %s must be positive" % n)
                p[..., i] = np.log(p[..., i])
        return p

    def invlink(self, theta):
        t = np.array(theta, dtype=np.float64, copy=True)
        for i, n in enumerate(self.param_names):
            if n in self.positive:
                t[..., i] = np.exp(t[..., i])
        return t

    def logabsdetjac(self, theta):
        """log |det J| of the inverse link at theta: what the linked log density adds to the user-space
        log joint (Bijectors; positive support links by log, so it is the sum of those thetas)"""
        t = np.asarray(theta, dtype=np.float64)
        return float(sum(t[i] for i, n in enumerate(self.param_names) if n in self.positive))

    def c_struct(self):
        """tnuts_model with host pointers into this object's arrays (kept alive by self)."""
        m = _lib.TnutsModel()
        m.family, m.D, m.N, m.G = self.family, self.D, int(self.N), int(self.G)
        dp = C.POINTER(C.c_double)
        m.X = self.X.ctypes.data_as(dp) if self.X is not None else None
        m.y = self.y.ctypes.data_as(dp) if self.y is not None else None
        m.sigma = self.sigma.ctypes.data_as(dp) if self.sigma is not None else None
        m.group = (self.group.ctypes.data_as(C.POINTER(C.c_int32))
                   if self.group is not None else None)
        for i, h in enumerate(self.hyper):
            m.hyper[i] = float(h)
        return m

    def data_nbytes(self):
        return sum(a.nbytes for a in (self.X, self.y, self.sigma, self.group) if a is not None)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class GDemo(Model):
    """@model function gdemo(x, y)
           s ~ InverseGamma(2, 3); m ~ Normal(0, sqrt(s))
           x ~ Normal(m, sqrt(s)); y ~ Normal(m, sqrt(s))
       end                                   (test/test_utils/models.jl:15-21)"""
    family = GDEMO
    positive = ("s",)
    param_names = ("s", "m")

    def __init__(self, x=1.5, y=2.0, alpha=2.0, beta=3.0):
        super().__init__()
        self.y = _f64([x, y] if np.isscalar(x) else np.concatenate([np.ravel(x), np.ravel(y)]))
        self.N = self.y.size
        self.hyper = (alpha, beta)


def gdemo_default():
    return GDemo(1.5, 2.0)


class LinearRegression(Model):
    """@model function linear_regression(x)
           alpha ~ Normal(0, 1); beta ~ Normal(0, 1)
           sigma2 ~ truncated(Cauchy(0, 3); lower=0)
           y ~ MvNormal(alpha .+ beta .* x, sigma2 * I)
       end                                   (README.md:38-47)"""
    family = LINREG
    positive = ("σ²",)
    param_names = ("α", "β", "σ²")

    def __init__(self, x, y, prior_sd=1.0, cauchy_scale=3.0):
        super().__init__()
        self.X, self.y = _f64(x), _f64(y)
        self.N = self.y.size
        self.hyper = (prior_sd, cauchy_scale)


class EightSchools(Model):
    """@model function eight_schools_nc(y, sigma)
           mu ~ Normal(0, 5); tau ~ truncated(Cauchy(0, 5); lower=0)
           theta_t ~ MvNormal(zeros(J), I)
           y ~ MvNormal(mu .+ tau .* theta_t, Diagonal(sigma .^ 2))
       end                                   (BASELINE.json configs[1])"""
    family = EIGHT_SCHOOLS
    positive = ("τ",)
    Y = (28.0, 8.0, -3.0, 7.0, -1.0, 1.0, 18.0, 12.0)
    SIGMA = (15.0, 10.0, 16.0, 11.0, 9.0, 11.0, 10.0, 18.0)

    def __init__(self, y=Y, sigma=SIGMA, mu_sd=5.0, cauchy_scale=5.0):
        super().__init__()
        self.y, self.sigma = _f64(y), _f64(sigma)
        self.N = self.y.size
        self.hyper = (mu_sd, cauchy_scale)
        self.param_names = ("μ", "τ") + tuple("θ̃[%d]" % (j + 1) for j in range(self.N))


class LogisticRegression(Model):
    """@model function logreg(X, y)
           beta ~ MvNormal(zeros(D), sd^2 * I)
           y .~ BernoulliLogit.(X * beta)
       end                                   (BASELINE.json configs[2], configs[4])"""
    family = LOGISTIC

    def __init__(self, X, y, prior_sd=1.0):
        super().__init__()
        self.X, self.y = _f64(X), _f64(y)
        self.N, d = self.X.shape
        self.hyper = (prior_sd,)
        self.param_names = tuple("β[%d]" % (j + 1) for j in range(d))


class HierLinearRegression(Model):
    """@model function hier_linreg(x, y, g, G)
           mu_a ~ Normal(0, 10); mu_b ~ Normal(0, 10)
           sigma_a ~ truncated(Normal(0, 5); lower=0); sigma_b ~ ...; sigma_y ~ ...
           a ~ filldist(Normal(mu_a, sigma_a), G); b ~ filldist(Normal(mu_b, sigma_b), G)
           y ~ MvNormal(a[g] .+ b[g] .* x, sigma_y^2 * I)
       end                                   (BASELINE.json configs[3])"""
    family = HIER_LINREG
    positive = ("σ_a", "σ_b", "σ_y")

    def __init__(self, x, y, group, n_groups, mu_sd=10.0, sigma_sd=5.0):
        super().__init__()
        self.X, self.y = _f64(x), _f64(y)
        self.group = np.ascontiguousarray(group, dtype=np.int32)
        self.N, self.G = self.y.size, int(n_groups)
        self.hyper = (mu_sd, sigma_sd)
        self.param_names = (("μ_a", "μ_b", "σ_a", "σ_b", "σ_y")
                            + tuple("a[%d]" % (g + 1) for g in range(self.G))
                            + tuple("b[%d]" % (g + 1) for g in range(self.G)))


class StdNormal(Model):
    """@model f() = x ~ MvNormal(zeros(D), I)     (test/main.jl:39-48 for D = 1)"""
    family = STD_NORMAL

    def __init__(self, D=1):
        super().__init__()
        self.param_names = ("x",) if D == 1 else tuple("x[%d]" % (i + 1) for i in range(D))


class BetaBernoulli(Model):
    """@model function constrained_test(obs)
           p ~ Beta(2, 2)
           for i in 1:length(obs); obs[i] ~ Bernoulli(p); end
       end                                   (test/mcmc/hmc.jl:24-43)
    p in (0, 1) links by logit (Bijectors)."""
    family = BETA_BERNOULLI
    param_names = ("p",)

    def __init__(self, obs, a=2.0, b=2.0):
        super().__init__()
        self.y = _f64(obs)
        self.N = self.y.size
        self.hyper = (a, b)

    def link(self, params):
        p = np.asarray(params, dtype=np.float64)
        if not (p.min() > 0 and p.max() < 1):
            raise ValueError("initial value of p must be in (0, 1)")
        return np.log(p) - np.log1p(-p)

    def invlink(self, theta):
        return 1.0 / (1.0 + np.exp(-np.asarray(theta, dtype=np.float64)))

    def logabsdetjac(self, theta):
        t = np.asarray(theta, dtype=np.float64)
        return float(np.sum(_log_sigmoid(t) + _log_sigmoid(-t)))     # log p + log(1 - p)


def _log_sigmoid(u):
    u = np.asarray(u, dtype=np.float64)
    return np.where(u >= 0, -np.log1p(np.exp(-np.abs(u))), u - np.log1p(np.exp(-np.abs(u))))


def _stick_link(x):
    """Bijectors' SimplexBijector (the Stan stick-breaking transform): K-simplex -> K - 1 reals"""
    x = np.asarray(x, dtype=np.float64)
    K = x.shape[-1]
    rem = 1.0 - np.concatenate([np.zeros(x.shape[:-1] + (1,)), np.cumsum(x[..., :-2], axis=-1)], axis=-1)
    z = x[..., :-1] / rem
    return np.log(z) - np.log1p(-z) + np.log(np.arange(K - 1, 0, -1))


def _stick_logabsdetjac(y):
    y = np.asarray(y, dtype=np.float64)
    K = y.shape[-1] + 1
    u = y - np.log(np.arange(K - 1, 0, -1))
    lz, l1z = _log_sigmoid(u), _log_sigmoid(-u)
    lrem = np.concatenate([[0.0], np.cumsum(l1z)[:-1]])
    return float(np.sum(lz + l1z + lrem))


def _stick_invlink(y):
    y = np.asarray(y, dtype=np.float64)
    K = y.shape[-1] + 1
    z = 1.0 / (1.0 + np.exp(-(y - np.log(np.arange(K - 1, 0, -1)))))
    x = np.empty(y.shape[:-1] + (K,))
    rem = np.ones(y.shape[:-1])
    for k in range(K - 1):
        x[..., k] = z[..., k] * rem
        rem = rem * (1.0 - z[..., k])
    x[..., K - 1] = rem
    return x


class DirichletCategorical(Model):
    """@model function constrained_simplex_test(obs12)
           ps ~ Dirichlet(2, 3)
           pd ~ Dirichlet(4, 1)
           for i in 1:length(obs12); obs12[i] ~ Categorical(ps); end
       end                                   (test/mcmc/hmc.jl:45-62)
    Each K-simplex links to K - 1 reals by stick breaking, so D = K1 - 1 + K2 - 1 while the chain
    has P = K1 + K2 parameter columns."""
    family = DIRICHLET_CATEGORICAL

    def __init__(self, obs, K1=2, alpha1=3.0, K2=4, alpha2=1.0):
        super().__init__()
        self.y = _f64(obs)
        self.N = self.y.size
        if self.N and not (self.y.min() >= 1 and self.y.max() <= K1):
            raise ValueError("observed categories must be in 1..K1")
        self.K1, self.K2 = int(K1), int(K2)
        self.hyper = (float(K1), alpha1, float(K2), alpha2)
        self.param_names = (tuple("ps[%d]" % (k + 1) for k in range(self.K1))
                            + tuple("pd[%d]" % (k + 1) for k in range(self.K2)))

    @property
    def D(self):
        return self.K1 - 1 + max(self.K2 - 1, 0)

    def link(self, params):
        p = np.asarray(params, dtype=np.float64)
        parts = [_stick_link(p[..., :self.K1])]
        if self.K2:
            parts.append(_stick_link(p[..., self.K1:]))
        return np.concatenate(parts, axis=-1)

    def invlink(self, theta):
        t = np.asarray(theta, dtype=np.float64)
        parts = [_stick_invlink(t[..., :self.K1 - 1])]
        if self.K2:
            parts.append(_stick_invlink(t[..., self.K1 - 1:]))
        return np.concatenate(parts, axis=-1)

    def logabsdetjac(self, theta):
        t = np.asarray(theta, dtype=np.float64)
        return _stick_logabsdetjac(t[:self.K1 - 1]) + (_stick_logabsdetjac(t[self.K1 - 1:]) if self.K2 else 0.0)


# ---- synthetic data of the BASELINE.json configs ----
def config1_linear_regression(seed=0):
    rng = np.random.default_rng(seed)
    x, y = rng.random(10), rng.random(10)
    return LinearRegression(x, y)


def config2_eight_schools():
    return EightSchools()


def synthetic_logistic(N, D, seed=1):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D))
    beta = rng.standard_normal(D) * (2.0 / np.sqrt(D))
    y = (rng.random(N) < 1.0 / (1.0 + np.exp(-(X @ beta)))).astype(np.float64)
    return LogisticRegression(X, y)


def config3_logistic(N=100_000, D=100, seed=1):
    return synthetic_logistic(N, D, seed)


def synthetic_hier(G, n_per, seed=2):
    rng = np.random.default_rng(seed)
    a = 1.0 + 0.7 * rng.standard_normal(G)
    b = -0.5 + 0.4 * rng.standard_normal(G)
    group = np.repeat(np.arange(G, dtype=np.int32), n_per)
    x = rng.standard_normal(G * n_per)
    y = a[group] + b[group] * x + 0.8 * rng.standard_normal(G * n_per)
    return HierLinearRegression(x, y, group, G)


def config4_hier(G=1000, n_per=100, seed=2):
    return synthetic_hier(G, n_per, seed)


class DeviceLogisticRegression(Model):
    """Logistic regression whose X [N, D] (row-major fp64) and y [N] already live on the GPU
    (torch CUDA tensors): used for row shards that are generated on the device
    (BASELINE.json configs[4]: 6.25 M rows x 256 per GPU = 12.8 GB, never staged on the host)."""
    family = LOGISTIC
    on_device = True

    def __init__(self, X_cuda, y_cuda, prior_sd=1.0):
        super().__init__()
        assert X_cuda.is_cuda and X_cuda.is_contiguous() and str(X_cuda.dtype) == "torch.float64"
        assert y_cuda.is_cuda and y_cuda.is_contiguous() and str(y_cuda.dtype) == "torch.float64"
        self._X, self._y = X_cuda, y_cuda
        self.N, d = X_cuda.shape
        self.hyper = (prior_sd,)
        self.param_names = tuple("β[%d]" % (j + 1) for j in range(d))

    def c_struct(self):
        m = _lib.TnutsModel()
        m.family, m.D, m.N, m.G = self.family, self.D, int(self.N), 0
        dp = C.POINTER(C.c_double)
        m.X = C.cast(self._X.data_ptr(), dp)
        m.y = C.cast(self._y.data_ptr(), dp)
        m.hyper[0] = float(self.hyper[0])
        return m

    def data_nbytes(self):
        return self._X.numel() * 8 + self._y.numel() * 8
