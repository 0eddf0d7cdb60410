import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import turing_jl_b200 as T
model = T.models.config3_logistic()
for C in (128, 512):
    e = T.Engine(model, T.NUTS(30, 0.8), C, 30, seed=3)
    e.set_option("logistic_tc", 1)
    e.set_option("time_kernels", 1)
    e.init()
    for persist in (1, 0, 1):
        e.set_option("persist", persist)
        t = time.perf_counter(); e.run(6, 1, 1); dt = time.perf_counter() - t
        print("C", C, "persist", persist, "device s %.3f" % e.L.tnuts_device_seconds(e.h), e.kernel_times(), flush=True)
    e.close()
