"""Sampler structs and ensemble markers with the reference's names, defaults and argument
meaning (src/mcmc/hmc.jl:59-80 HMC, :285-327 HMCDA, :352-402 NUTS; src/mcmc/sghmc.jl SGHMC, SGLD,
PolynomialStepsize; src/Turing.jl:105-107)."""


METRICS = {"diag": 0, "DiagEuclideanMetric": 0, "unit": 1, "UnitEuclideanMetric": 1,
           "dense": 2, "DenseEuclideanMetric": 2}


def _metric_id(metricT):
    """metricT of the reference's constructors (AHMC.DiagEuclideanMetric / UnitEuclideanMetric /
    DenseEuclideanMetric; the dense one runs on the thread-per-chain strategy and, for D <= 256, on the
    pump strategy)"""
    try:
        return METRICS[metricT]
    except KeyError:
        raise ValueError("unknown metricT %r" % (metricT,)) from None


class Hamiltonian:
    """`allow_discrete_variables = false` (src/mcmc/hmc.jl:4); InitFromUniform (:82)."""
    algorithm = -1
    adaptive = False


class NUTS(Hamiltonian):
    """NUTS()                     -> n_adapts=-1, delta=0.65          (hmc.jl:400-402)
    NUTS(delta)                   -> n_adapts=-1                       (hmc.jl:389-398)
    NUTS(n_adapts, delta; max_depth=10, Delta_max=1000.0, init_eps=0.0) (hmc.jl:377-387)"""
    algorithm = 0
    adaptive = True

    def __init__(self, *args, max_depth=10, Delta_max=1000.0, init_eps=0.0, metricT="diag"):
        if len(args) == 0:
            n_adapts, delta = -1, 0.65
        elif len(args) == 1:
            if not isinstance(args[0], float):
                raise TypeError("NUTS(delta): delta must be a Float64 (use NUTS(n_adapts, delta))")
            n_adapts, delta = -1, args[0]
        elif len(args) == 2:
            n_adapts, delta = int(args[0]), float(args[1])
        else:
            raise TypeError("NUTS takes at most (n_adapts, delta)")
        self.metric = _metric_id(metricT)
        self.n_adapts, self.delta = n_adapts, float(delta)
        self.max_depth, self.Delta_max, self.eps = int(max_depth), float(Delta_max), float(init_eps)


class HMC(Hamiltonian):
    """HMC(eps, n_leapfrog): static trajectory, no adaptation (hmc.jl:59-80)."""
    algorithm = 1
    adaptive = False

    def __init__(self, eps, n_leapfrog, metricT="unit"):
        self.metric = _metric_id(metricT)     # no adaptation at all: the metric stays as created
        self.eps, self.n_leapfrog = float(eps), int(n_leapfrog)
        self.n_adapts, self.delta = 0, 0.0
        self.max_depth, self.Delta_max = 10, 1000.0


class HMCDA(Hamiltonian):
    """HMCDA(n_adapts, delta, lambda; init_eps=0.0) / HMCDA(delta, lambda) (hmc.jl:285-327)."""
    algorithm = 2
    adaptive = True

    def __init__(self, *args, init_eps=0.0, metricT="unit"):
        self.metric = _metric_id(metricT)     # default UnitEuclideanMetric (hmc.jl:308, 323)
        if len(args) == 2:
            n_adapts, delta, lam = -1, args[0], args[1]
        elif len(args) == 3:
            n_adapts, delta, lam = args
        else:
            raise TypeError("HMCDA(n_adapts, delta, lambda) or HMCDA(delta, lambda)")
        self.n_adapts, self.delta, self.lam = int(n_adapts), float(delta), float(lam)
        self.eps = float(init_eps)
        self.max_depth, self.Delta_max = 10, 1000.0


class SGHMC(Hamiltonian):
    """SGHMC(; learning_rate, momentum_decay) (src/mcmc/sghmc.jl:15-47): theta += v;
    v = (1 - alpha) v + eta grad + sqrt(2 eta alpha) randn -- eq. (15) of Chen et al. (2014)."""
    algorithm = 3
    adaptive = False

    def __init__(self, *, learning_rate, momentum_decay):
        self.learning_rate, self.momentum_decay = float(learning_rate), float(momentum_decay)
        # engine config fields (include/tnuts.h): init_eps = learning_rate, lambda = momentum_decay
        self.eps, self.lam, self.delta = self.learning_rate, self.momentum_decay, 0.0
        self.n_adapts, self.max_depth, self.Delta_max = 0, 10, 1000.0


class PolynomialStepsize:
    """PolynomialStepsize(a[, b=0, gamma=0.55]): step size a (b + t)^-gamma at iteration t
    (src/mcmc/sghmc.jl:124-161)."""

    def __init__(self, a, b=0.0, gamma=0.55):
        if not (0.5 < gamma <= 1):
            raise ValueError("the decay rate `γ` has to be in (0.5, 1]")
        self.a, self.b, self.gamma = float(a), float(b), float(gamma)

    def __call__(self, t):
        return self.a / (t + self.b) ** self.gamma


class SGLD(Hamiltonian):
    """SGLD(; stepsize = PolynomialStepsize(0.01)) (src/mcmc/sghmc.jl:163-186):
    theta += eps_t / 2 grad + sqrt(eps_t) randn; the chain carries the `SGLD_stepsize` column."""
    algorithm = 4
    adaptive = False

    def __init__(self, *, stepsize=None):
        self.stepsize = stepsize if stepsize is not None else PolynomialStepsize(0.01)
        if not isinstance(self.stepsize, PolynomialStepsize):
            raise TypeError("the engine evaluates the step size on the device: it must be a PolynomialStepsize")
        # engine config fields (include/tnuts.h): init_eps = a, delta = b, lambda = gamma
        self.eps, self.delta, self.lam = self.stepsize.a, self.stepsize.b, self.stepsize.gamma
        self.n_adapts, self.max_depth, self.Delta_max = 0, 10, 1000.0


class MCMCThreads:
    """all chains advance together on the GPU(s)"""


class MCMCSerial(MCMCThreads):
    """`sample(model, spl, MCMCSerial(), N, n_chains)`: the reference runs the chains one after the
    other on one task (AbstractMCMC.mcmcsample); the result is defined to be the same
    chains as MCMCThreads gives (each chain has its own RNG stream), so the engine serves it with the
    same batched launch -- same draws, only the schedule differs."""


class MCMCDistributed(MCMCThreads):
    """`MCMCDistributed()`: one chain per worker process upstream.  Here the counterpart is the
    one-process-per-GPU launch (`turing_jl_b200.distributed`, chains sharded over ranks with global
    Philox chain ids); inside one process it is served like MCMCThreads."""


class InitFromUniform:
    """theta ~ U(-2, 2) in linked space (HISTORY.md:689; src/mcmc/hmc.jl:82)"""


class InitFromParams:
    """user-space (constrained) initial values: mapping name -> value, or a vector in model order"""

    def __init__(self, params):
        self.params = params
