"""times the logistic gradient kernels (fp64 DMMA path) for several shapes; profiles/ cites it"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import turing_jl_b200 as T
for (N, D, C) in ((100000, 100, 4096), (100000, 100, 1024), (100000, 100, 256), (100000, 100, 64), (200000, 256, 256), (20000, 32, 4096)):
    model = T.models.synthetic_logistic(N, D)
    eng = T.Engine(model, T.NUTS(0.8, init_eps=0.01), C, 0, seed=1)
    eng.init()
    dt = eng.bench_gradient(5)
    print("logistic gradient N=%d D=%d C=%d: %.3f ms -> %.2f TFLOP/s (4ND per chain), %.3e grad/s" % (N, D, C, dt * 1e3, 4.0 * N * D * C / dt / 1e12, C / dt))
    eng.close()
for (G, n_per, C) in ((1000, 100, 1024), (1000, 100, 8192)):
    model = T.models.synthetic_hier(G, n_per)
    eng = T.Engine(model, T.NUTS(0.8, init_eps=0.01), C, 0, seed=1)
    eng.init()
    dt = eng.bench_gradient(20)
    print("hier gradient G=%d D=%d C=%d: %.3f ms -> %.1f GB/s (2*D*8 B per chain)" % (G, model.D, C, dt * 1e3, 2 * model.D * 8 * C / dt / 1e9))
    eng.close()
