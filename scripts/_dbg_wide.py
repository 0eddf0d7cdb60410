import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import turing_jl_b200 as T
np.set_printoptions(linewidth=250, precision=3, suppress=True)
for (N, D, n) in ((64, 256, 4), (64, 224, 4), (64, 200, 4), (3001, 256, 130)):
    model = T.models.synthetic_logistic(N, D, seed=N + D)
    th = np.random.default_rng(N).standard_normal((n, D)) * 0.5
    lp, g = T.LogDensityFunction(model, options={"logistic_tc": 1}).logdensity_and_gradient(th)
    rlp, rg = T.LogDensityFunction(model).logdensity_and_gradient(th)
    print("N %d D %d n %d  lp relerr %.2e" % (N, D, n, np.max(np.abs(lp - rlp) / np.abs(rlp))))
    for c in (0, n - 1):
        e = np.abs(g[c] - rg[c])
        nb = (D + 15) // 16
        print("  chain %d |err| per 16-feature block / |g| block:" % c,
              np.array([np.linalg.norm(e[16 * b:16 * b + 16]) / (np.linalg.norm(rg[c][16 * b:16 * b + 16]) + 1e-30) for b in range(nb)]))
    if N == 64:
        print("  tc  g[0, 186:200]", g[0, 186:200])
        print("  ref g[0, 186:200]", rg[0, 186:200])
        print("  ref g[0, 58:72]  ", rg[0, 58:72], " (features - 128)")
        print("  ref g[0, 122:136]  ", rg[0, 122:136], " (features - 64)")
