"""Generates tests/golden/lpgrad_golden.json: log joint density in linked space and its gradient
for every registered family at fixed points, evaluated INDEPENDENTLY of oracle/ and of the CUDA
engine with mpmath at 50 digits, straight from the distribution definitions
(Distributions.jl logpdf formulas; Bijectors log link + log-Jacobian).
Gradients are mpmath numerical derivatives at 50 digits (error << 1e-20).

The reference (Turing.jl) cannot run in this image (no Julia), so these vectors stand in for
`LogDensityProblems.logdensity_and_gradient(ldf, theta)` (src/mcmc/hmc.jl:181) on the same
inputs.  Also records the known answers the reference's own tests pin for gdemo.

Run:  python tests/golden/make_golden.py
"""
import json
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 50
PI = mp.pi


def normal_lpdf(x, mu, sd):
    return -mp.log(sd) - mp.mpf("0.5") * mp.log(2 * PI) - (x - mu) ** 2 / (2 * sd ** 2)


def invgamma_lpdf(x, a, b):
    return a * mp.log(b) - mp.loggamma(a) - (a + 1) * mp.log(x) - b / x


def halfcauchy_lpdf(x, s):  # truncated(Cauchy(0, s); lower=0)
    return mp.log(2) - mp.log(PI * s) - mp.log(1 + (x / s) ** 2)


def halfnormal_lpdf(x, s):  # truncated(Normal(0, s); lower=0)
    return mp.log(2) + normal_lpdf(x, 0, s)


def bernoulli_logit_lpmf(y, eta):
    return y * eta - mp.log(1 + mp.exp(eta))


def lj_gdemo(th, data):
    s, m = mp.exp(th[0]), th[1]
    lp = invgamma_lpdf(s, data["hyper"][0], data["hyper"][1]) + normal_lpdf(m, 0, mp.sqrt(s))
    for d in data["y"]:
        lp += normal_lpdf(d, m, mp.sqrt(s))
    return lp + th[0]


def lj_linreg(th, data):
    a, b, v = th[0], th[1], mp.exp(th[2])
    sd, c = data["hyper"]
    lp = normal_lpdf(a, 0, sd) + normal_lpdf(b, 0, sd) + halfcauchy_lpdf(v, c)
    for x, y in zip(data["X"], data["y"]):
        lp += normal_lpdf(y, a + b * x, mp.sqrt(v))
    return lp + th[2]


def lj_eight(th, data):
    mu, tau = th[0], mp.exp(th[1])
    sd, c = data["hyper"]
    lp = normal_lpdf(mu, 0, sd) + halfcauchy_lpdf(tau, c)
    for j, (y, s) in enumerate(zip(data["y"], data["sigma"])):
        lp += normal_lpdf(th[2 + j], 0, 1) + normal_lpdf(y, mu + tau * th[2 + j], s)
    return lp + th[1]


def lj_logistic(th, data):
    sd = data["hyper"][0]
    lp = mp.mpf(0)
    for t in th:
        lp += normal_lpdf(t, 0, sd)
    for row, y in zip(data["X"], data["y"]):
        eta = mp.fsum(x * t for x, t in zip(row, th))
        lp += bernoulli_logit_lpmf(y, eta)
    return lp


def lj_hier(th, data):
    sdm, s0 = data["hyper"]
    G = data["G"]
    mua, mub = th[0], th[1]
    sa, sb, sy = mp.exp(th[2]), mp.exp(th[3]), mp.exp(th[4])
    lp = normal_lpdf(mua, 0, sdm) + normal_lpdf(mub, 0, sdm)
    lp += halfnormal_lpdf(sa, s0) + halfnormal_lpdf(sb, s0) + halfnormal_lpdf(sy, s0)
    for g in range(G):
        lp += normal_lpdf(th[5 + g], mua, sa) + normal_lpdf(th[5 + G + g], mub, sb)
    for x, y, g in zip(data["X"], data["y"], data["group"]):
        lp += normal_lpdf(y, th[5 + g] + th[5 + G + g] * x, sy)
    return lp + th[2] + th[3] + th[4]


def lj_stdnormal(th, data):
    return mp.fsum(normal_lpdf(t, 0, 1) for t in th)


def beta_lpdf(x, a, b):
    return (a - 1) * mp.log(x) + (b - 1) * mp.log(1 - x) - (mp.loggamma(a) + mp.loggamma(b) - mp.loggamma(a + b))


def dirichlet_lpdf(x, al):
    K = len(x)
    return mp.loggamma(K * al) - K * mp.loggamma(al) + (al - 1) * mp.fsum(mp.log(v) for v in x)


def simplex_invlink(y):
    """Bijectors' SimplexBijector (the Stan stick-breaking transform): K - 1 reals -> K-simplex"""
    K = len(y) + 1
    x, rem = [], mp.mpf(1)
    for k, yk in enumerate(y):
        z = 1 / (1 + mp.exp(-(yk - mp.log(K - 1 - k))))
        x.append(z * rem)
        rem = rem * (1 - z)
    return x + [rem]


def simplex_logabsdetjac(y):
    """log |det d x_{1..K-1} / d y| from the DEFINITION: numerical Jacobian at 50 digits"""
    n = len(y)
    J = mp.zeros(n, n)
    for j in range(n):
        def col(t, j=j):
            v = list(y)
            v[j] = t
            return simplex_invlink(v)
        for i in range(n):
            J[i, j] = mp.diff(lambda t, i=i, j=j: col(t)[i], y[j])
    return mp.log(abs(mp.det(J)))


def lj_beta_bernoulli(th, data):
    """test/mcmc/hmc.jl:24-43: p ~ Beta(a, b); obs[i] ~ Bernoulli(p); logit link"""
    a, b = data["hyper"]
    p = 1 / (1 + mp.exp(-th[0]))
    lp = beta_lpdf(p, a, b)
    for o in data["y"]:
        lp += mp.log(p) if o != 0 else mp.log(1 - p)
    return lp + mp.log(p) + mp.log(1 - p)      # log |d p / d theta|


def lj_dirichlet_cat(th, data):
    """test/mcmc/hmc.jl:45-62: ps ~ Dirichlet(K1, al1); pd ~ Dirichlet(K2, al2); obs[i] ~ Categorical(ps)"""
    K1, al1, K2, al2 = data["hyper"]
    K1, K2 = int(K1), int(K2)
    y1, y2 = list(th[:K1 - 1]), list(th[K1 - 1:])
    ps = simplex_invlink(y1)
    lp = dirichlet_lpdf(ps, al1) + simplex_logabsdetjac(y1)
    for o in data["y"]:
        lp += mp.log(ps[int(o) - 1])
    if K2 > 0:
        lp += dirichlet_lpdf(simplex_invlink(y2), al2) + simplex_logabsdetjac(y2)
    return lp


FAMS = {"beta_bernoulli": (6, lj_beta_bernoulli), "dirichlet_categorical": (7, lj_dirichlet_cat),
        "gdemo": (0, lj_gdemo), "linreg": (1, lj_linreg), "eight_schools": (2, lj_eight),
        "logistic": (3, lj_logistic), "hier": (4, lj_hier), "std_normal": (5, lj_stdnormal)}


def grad(f, th, data):
    g = []
    for i in range(len(th)):
        def fi(t, i=i):
            v = list(th)
            v[i] = t
            return f(v, data)
        g.append(mp.diff(fi, th[i]))
    return g


def main():
    rng = np.random.default_rng(20261019)
    cases = []

    def add(name, data, thetas):
        fam, f = FAMS[name]
        mpdata = {k: ([mp.mpf(float(x)) for x in v] if k in ("y", "sigma", "hyper") else v)
                  for k, v in data.items()}
        if "X" in data:
            X = np.asarray(data["X"], dtype=np.float64)
            mpdata["X"] = ([[mp.mpf(float(x)) for x in row] for row in X] if X.ndim == 2
                           else [mp.mpf(float(x)) for x in X])
        pts = []
        for th in thetas:
            mth = [mp.mpf(float(t)) for t in th]
            lp = f(mth, mpdata)
            g = grad(f, mth, mpdata)
            pts.append({"theta": [float(t) for t in th], "lp": float(lp), "grad": [float(x) for x in g]})
        cases.append({"name": name, "family": fam,
                      "data": {k: (np.asarray(v).tolist()) for k, v in data.items()}, "points": pts})

    add("gdemo", {"y": [1.5, 2.0], "hyper": [2.0, 3.0]},
        [[0.0, 0.0], [0.3, -0.4]] + rng.standard_normal((4, 2)).tolist())
    x, y = np.random.default_rng(0).random(10), np.random.default_rng(0).random(10)
    r0 = np.random.default_rng(0)
    x, y = r0.random(10), r0.random(10)
    add("linreg", {"X": x, "y": y, "hyper": [1.0, 3.0]}, rng.standard_normal((5, 3)).tolist())
    add("eight_schools", {"y": [28, 8, -3, 7, -1, 1, 18, 12], "sigma": [15, 10, 16, 11, 9, 11, 10, 18],
                          "hyper": [5.0, 5.0]}, (rng.standard_normal((5, 10)) * 1.2).tolist())
    X = rng.standard_normal((40, 7))
    yb = (rng.random(40) < 0.5).astype(float)
    add("logistic", {"X": X, "y": yb, "hyper": [1.0]}, (rng.standard_normal((4, 7)) * 1.5).tolist()
        + [(np.ones(7) * 12.0).tolist()])  # large |eta|: log1pexp branches
    G = 4
    grp = np.repeat(np.arange(G), 6)
    add("hier", {"X": rng.standard_normal(G * 6), "y": rng.standard_normal(G * 6), "group": grp, "G": G,
                 "hyper": [10.0, 5.0]}, (rng.standard_normal((4, 5 + 2 * G)) * 0.8).tolist())
    add("std_normal", {}, [[0.0], [1.25]])
    # the constrained-support models of the reference's HMC tests (their own data); a separate
    # generator so that the cases above keep their points
    r2 = np.random.default_rng(7)
    add("beta_bernoulli", {"y": [0, 1, 0, 1, 1, 1, 1, 1, 1, 1], "hyper": [2.0, 2.0]},
        [[0.0], [0.9], [-3.5], [25.0]])
    add("dirichlet_categorical", {"y": [1, 2, 1, 2, 2, 2, 2, 2, 2, 2], "hyper": [2.0, 3.0, 4.0, 1.0]},
        [[0.0, 0.0, 0.0, 0.0]] + (r2.standard_normal((4, 4)) * 1.5).tolist())
    add("dirichlet_categorical", {"y": [3, 1, 3, 2, 3], "hyper": [3.0, 0.7, 0.0, 0.0]},
        (r2.standard_normal((3, 2)) * 1.2).tolist())
    out = {
        "generator": "tests/golden/make_golden.py (mpmath, 50 digits)",
        "reference_known_answers": {
            "gdemo_lp_at_0_0": -6.6845910222777984,
            "gdemo_posterior_mean": {"s": 49 / 24, "m": 7 / 6,
                                     "src": "test/test_utils/numerical_tests.jl:21-25"},
            "gdemo_map_user_space": {"s": 49 / 54, "m": 7 / 6,
                                     "src": "test/optimisation/Optimisation.jl:299-301"},
            "gdemo_mle": {"s": 0.0625, "m": 1.75, "src": "test/optimisation/Optimisation.jl:246-248"},
            "std_normal_lp_at_0": float(normal_lpdf(mp.mpf(0), 0, 1)),
            "beta_bernoulli_posterior_mean": {"p": 10 / 14, "atol": 0.1, "sampler": "HMC(1.5, 3), 1000 draws",
                                              "src": "test/mcmc/hmc.jl:24-43"},
            "dirichlet_categorical_posterior_mean": {"ps": [5 / 16, 11 / 16], "atol": 0.015,
                                                     "sampler": "HMC(0.75, 2), 1000 draws",
                                                     "src": "test/mcmc/hmc.jl:45-62"},
        },
        "cases": cases,
    }
    with open(os.path.join(os.path.dirname(__file__), "lpgrad_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
