"""shared test helpers: the same data as a host model descriptor and as an oracle model"""
import json
import os

import numpy as np

import oracle as O
import turing_jl_b200 as T

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "lpgrad_golden.json")


def oracle_of(model):
    return O.OracleModel(model.family, D=model.D, N=model.N, G=model.G, X=model.X, y=model.y,
                         sigma=model.sigma, group=model.group, hyper=model.hyper)


def model_from_golden(case):
    d, name = case["data"], case["name"]
    if name == "gdemo":
        return T.GDemo(d["y"][0], d["y"][1], *d["hyper"])
    if name == "linreg":
        return T.LinearRegression(d["X"], d["y"], *d["hyper"])
    if name == "eight_schools":
        return T.EightSchools(d["y"], d["sigma"], *d["hyper"])
    if name == "logistic":
        return T.LogisticRegression(np.array(d["X"]), d["y"], *d["hyper"])
    if name == "hier":
        return T.HierLinearRegression(d["X"], d["y"], d["group"], d["G"], *d["hyper"])
    if name == "beta_bernoulli":
        return T.BetaBernoulli(d["y"], *d["hyper"])
    if name == "dirichlet_categorical":
        K1, a1, K2, a2 = d["hyper"]
        return T.DirichletCategorical(d["y"], int(K1), a1, int(K2), a2)
    if name == "std_normal":
        return T.StdNormal(len(case["points"][0]["theta"]))
    raise KeyError(name)


def golden_cases():
    with open(GOLDEN) as f:
        return json.load(f)


def small_models():
    rng = np.random.default_rng(11)
    return {
        "gdemo": T.gdemo_default(),
        "linreg": T.models.config1_linear_regression(),
        "eight_schools": T.EightSchools(),
        # families / sizes that only the lock-step strategy serves
        "linreg_n40": T.LinearRegression(rng.random(40), rng.random(40)),
        "eight_schools_j5": T.EightSchools([3.0, -1.0, 8.0, 2.0, 0.5], [4.0, 6.0, 5.0, 3.0, 7.0]),
        "std_normal_d7": T.StdNormal(7),
        "logistic_small": T.models.synthetic_logistic(300, 5, seed=4),
        "logistic_d130": T.models.synthetic_logistic(257, 130, seed=5),
        "hier_small": T.models.synthetic_hier(7, 11, seed=6),
        "hier_g150": T.models.synthetic_hier(150, 9, seed=7),     # D = 305: two dims per thread
        "std_normal_d600": T.StdNormal(600),
    }


THREAD_MODELS = ("gdemo", "linreg", "eight_schools")


def oracle_config(spl, n_adapts, seed, mode=1):
    return O.make_config(delta=spl.delta, max_depth=spl.max_depth, delta_max=spl.Delta_max,
                         init_eps=spl.eps, n_adapts=n_adapts, seed=seed, sampler_mode=mode,
                         algorithm=spl.algorithm, n_leapfrog=getattr(spl, "n_leapfrog", 0),
                         lambda_=getattr(spl, "lam", 0.0), metric=getattr(spl, "metric", 0))
