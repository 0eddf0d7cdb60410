"""Per-chain cost spread of a many-chain NUTS run of BASELINE configs[2]: adapted step sizes and
leapfrogs per transition over the chains (what bounds the asynchronous pump's tail)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import turing_jl_b200 as T

C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
model = T.models.config3_logistic()
ch = T.sample(model, T.NUTS(0.8), T.MCMCThreads(), 1000, C, seed=20261019, verbose=False,
              engine_options={"logistic_tc": 1}, discard_adapt=False)
ns = np.asarray(ch["n_steps"])      # [iters, chains]
td = np.asarray(ch["tree_depth"])
eps = np.asarray(ch["step_size"])
print("shape", ns.shape, "device s", ch.info["device_seconds"])
q = [0, 0.001, 0.01, 0.1, 0.5, 0.9, 0.99, 0.999, 1]
nad = 500
tot = ns.sum(0)
print("total leapfrogs per chain quantiles", np.quantile(tot, q), "mean", tot.mean())
print("warm-up leapfrogs per chain quantiles", np.quantile(ns[:nad].sum(0), q), "mean", ns[:nad].sum(0).mean())
post = ns[nad:].sum(0)
print("post-warm-up leapfrogs per chain quantiles", np.quantile(post, q), "mean", post.mean())
print("final eps quantiles", np.quantile(eps[-1], q))
worst = np.argsort(post)[-8:]
print("worst chains", worst, "eps", eps[-1, worst], "post/transition", post[worst] / (ns.shape[0] - nad),
      "depth mean", td[nad:, worst].mean(0))
for lo, hi in ((0, 25), (25, 75), (75, 100), (100, 150), (150, 250), (250, 450), (450, 500), (500, 1500)):
    s = ns[lo:hi].sum(0)
    print("transitions %4d-%4d: leapfrogs per chain mean %8.0f max %8.0f  max/mean %.2f" % (lo, hi, s.mean(), s.max(), s.max() / s.mean()))
