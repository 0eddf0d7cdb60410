import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import turing_jl_b200 as T
model = T.EightSchools(); spl = T.NUTS(0.8); C = 65536
for K in (4, 8, 16, 24, 32):
    e = T.Engine(model, spl, C, 500, 1)
    e.set_option("flat", 1); e.set_option("flat_batch", K)
    for rep in range(2):
        e.init(); e.run(1499, 500, 1)
    dt = e.L.tnuts_device_seconds(e.h)
    print("flat batch %d: run %.4f s, %.3e grad/s" % (K, dt, e.grad_evals / dt))
    e.close()
