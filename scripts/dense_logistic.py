"""DenseEuclideanMetric where it matters: Bayesian logistic regression with correlated features on the
pump strategy (tcgen05 likelihood), dense against diagonal metric: leapfrogs per draw, min bulk-ESS"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import turing_jl_b200 as T
N, D, C = 100_000, 100, 256
WARM = int(os.environ.get("WARM", "1000"))
rng = np.random.default_rng(5)
# correlated design: features share a few latent factors
F = rng.standard_normal((N, 6))
X = 0.7 * F @ rng.standard_normal((6, D)) + 0.5 * rng.standard_normal((N, D))
beta = rng.standard_normal(D) * 0.15
y = (rng.random(N) < 1 / (1 + np.exp(-X @ beta))).astype(np.float64)
model = T.LogisticRegression(X, y)
for metric in ("diag", "dense"):
    t0 = time.time()
    ch = T.sample(model, T.NUTS(WARM, 0.8, metricT=metric), T.MCMCThreads(), 300, C, seed=2, verbose=False,
                  diagnostics=True, engine_options={"logistic_tc": 1})
    dt = time.time() - t0
    print("%s metric: %.1f s, n_steps mean %.1f, tree depth mean %.2f, step size mean %.3f, rhat max %.4f, min bulk ESS %.0f (%.0f / s)"
          % (metric, dt, np.nanmean(ch["n_steps"]),
             np.nanmean(ch["tree_depth"]),
             np.mean(ch.info["step_size"]), np.max(ch.info["rhat"]), np.min(ch.info["ess_bulk"]), np.min(ch.info["ess_bulk"]) / dt), flush=True)
