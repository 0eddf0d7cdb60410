"""turing.jl_b200 -- host side of the B200-native many-chain HMC/NUTS engine.

Holds only what the hot path needs: `csrc/` (CUDA kernels + the C ABI of include/tnuts.h)
and a Python mirror of the reference's sampler interface for that path
(`sample(model, NUTS(delta), MCMCThreads(), N, n_chains)`).  The directory name contains a
dot, so import it through the repo-root shim:  `import turing_jl_b200`.
"""
from . import _lib, build, distributed, models  # noqa: F401
from .chains import Chains, ParamsWithStats, VNChain, chainscat, params_with_stats  # noqa: F401
from .engine import Engine  # noqa: F401
from .inference import HMCState, LogDensityFunction, loadstate, post_sample_hook, sample, step  # noqa: F401
from .models import (BetaBernoulli, DirichletCategorical, EightSchools, GDemo, HierLinearRegression, LinearRegression,  # noqa: F401
                     LogisticRegression, StdNormal, gdemo_default)
from .samplers import (HMC, HMCDA, NUTS, SGHMC, SGLD, InitFromParams, InitFromUniform,  # noqa: F401
                       MCMCDistributed, MCMCSerial, MCMCThreads, PolynomialStepsize)

__all__ = ["sample", "NUTS", "HMC", "HMCDA", "SGHMC", "SGLD", "PolynomialStepsize", "MCMCThreads", "MCMCSerial", "MCMCDistributed",
           "Chains", "VNChain", "ParamsWithStats", "params_with_stats", "HMCState", "step", "chainscat", "loadstate", "LogDensityFunction", "Engine", "InitFromParams",
           "InitFromUniform", "GDemo", "gdemo_default", "LinearRegression", "EightSchools",
           "LogisticRegression", "HierLinearRegression", "StdNormal", "BetaBernoulli", "DirichletCategorical",
           "models"]
