import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import turing_jl_b200 as T
model = T.EightSchools(); spl = T.NUTS(0.8); C = 65536
e = T.Engine(model, spl, C, 500, 1)
for rep in range(3):
    e.init(); e.run(1499, 500, 1)
    print("run %.4f s, %.3e grad/s" % (e.L.tnuts_device_seconds(e.h), e.grad_evals / e.L.tnuts_device_seconds(e.h)))
e.close()
