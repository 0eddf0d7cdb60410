"""first GPU check of the thread-per-chain strategy (dev script)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O
from oracle import diagnostics as dg
import turing_jl_b200 as T
import __graft_entry__ as G

G.smoke()

def om_of(model):
    fam = model.family
    return O.OracleModel(fam, D=model.D, N=model.N, G=model.G, X=model.X, y=model.y, sigma=model.sigma, group=model.group, hyper=model.hyper)

for model in (T.gdemo_default(), T.models.config1_linear_regression(), T.EightSchools()):
    om = om_of(model)
    for nad in (0, 30, 200):
        spl = T.NUTS(nad, 0.8)
        eng = T.Engine(model, spl, n_chains=64, n_adapts=nad, seed=5)
        eng.init(); eng.run(nad + 50, 1, 1); out = eng.fetch(theta=True)
        cfg = O.make_config(delta=0.8, n_adapts=nad, seed=5, sampler_mode=1)
        ref = O.sample_chains(om, cfg, 64, nad + 50)
        st = out["stats"].transpose(2, 0, 1); th = out["theta"].transpose(2, 0, 1)
        same = (st[..., 0] == ref["stats"][..., 0]) & (st[..., 6] == ref["stats"][..., 6])
        # per chain: first transition index where stats differ
        firstbad = [int(np.argmin(s)) if not s.all() else -1 for s in same]
        d = np.abs(th - ref["theta"])
        print(type(model).__name__, "nadapt", nad, "same-frac %.4f" % same.mean(), "chains fully identical:", sum(f == -1 for f in firstbad),
              "max dtheta first10 %.2e" % d[:, :10].max(), "eps diff %.2e" % np.abs(eng.stepsize() - ref["eps"]).max())
        eng.close()

# throughput: eight schools
for C in (4096, 65536):
    model = T.EightSchools()
    t = time.time()
    ch = T.sample(model, T.NUTS(0.8), T.MCMCThreads(), 1000, C, seed=1, keep_engine=False)
    wall = time.time() - t
    print("eight schools C=%d wall %.2fs device %.3fs grad evals %.3e  -> %.3e grad/s (device)" % (C, wall, ch.info["device_seconds"], ch.info["grad_evals"], ch.info["grad_evals"] / ch.info["device_seconds"]))
    print(" mean depth %.2f nsteps %.2f acc %.3f div %d" % (np.nanmean(ch["tree_depth"]), np.nanmean(ch["n_steps"]), np.nanmean(ch["acceptance_rate"]), np.nansum(ch["numerical_error"])))
    print(" mu mean %.3f tau mean %.3f" % (ch.mean("μ"), ch.mean("τ")))
    if C == 4096:
        s = dg.summarize(ch.get_params().transpose(2, 0, 1)[:64])
        print(" (64 chains) ess", s["ess_bulk"].min(), "rhat", s["rhat"].max())
