"""lock-step strategy: quick GPU check + timings (dev script)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O
from oracle import diagnostics as dg
import turing_jl_b200 as T

def om_of(model):
    return O.OracleModel(model.family, D=model.D, N=model.N, G=model.G, X=model.X, y=model.y, sigma=model.sigma, group=model.group, hyper=model.hyper)

# throughput: logistic at growing sizes
for (N, D, C, ndraw) in ((10000, 100, 1024, 40), (100000, 100, 4096, 20)):
    model = T.models.synthetic_logistic(N, D)
    t = time.time()
    eng = T.Engine(model, T.NUTS(0.8), C, ndraw // 2, seed=1)
    t_up = time.time() - t
    t = time.time(); eng.init(); t_init = time.time() - t
    g0 = eng.grad_evals
    t = time.time(); eng.run(ndraw, 1, 1); t_run = time.time() - t
    g1 = eng.grad_evals
    out = eng.fetch()
    st = out["stats"]
    print("logistic N=%d D=%d C=%d: upload %.2fs init %.2fs (%d grads) run %.2fs dev %.3fs, %d grads -> %.3e grad/s; %.3e flop/s; launches %d" % (
        N, D, C, t_up, t_init, g0, t_run, eng.device_seconds, g1 - g0, (g1 - g0) / t_run, (g1 - g0) / t_run * 4 * N * D, eng.launches))
    print("   depth mean %.2f max %d  nsteps mean %.1f  acc %.3f eps %.4f" % (st[:, 6].mean(), st[:, 6].max(), st[:, 0].mean(), st[:, 1].mean(), eng.stepsize().mean()))
    eng.close()

# hier
for (G, n_per, C, ndraw) in ((100, 20, 256, 60), (1000, 100, 1024, 30)):
    model = T.models.synthetic_hier(G, n_per)
    eng = T.Engine(model, T.NUTS(0.8), C, ndraw // 2, seed=1)
    t = time.time(); eng.init(); t_init = time.time() - t
    g0 = eng.grad_evals
    t = time.time(); eng.run(ndraw, 1, 1); t_run = time.time() - t
    g1 = eng.grad_evals
    st = eng.fetch()["stats"]
    print("hier G=%d D=%d C=%d: init %.2fs run %.2fs, %d grads -> %.3e grad/s; bytes/s (6*D*8) %.3e; launches %d" % (
        G, model.D, C, t_init, t_run, g1 - g0, (g1 - g0) / t_run, (g1 - g0) / t_run * 6 * model.D * 8, eng.launches))
    print("   depth mean %.2f max %d nsteps mean %.1f acc %.3f" % (st[:, 6].mean(), st[:, 6].max(), st[:, 0].mean(), st[:, 1].mean()))
    eng.close()
