"""profile target: the tcgen05 logistic likelihood kernel at BASELINE configs[2]'s full width
(N = 100 000 rows, D = 100, 4096 chains, 37 row chunks x 32 chain tiles):
  ncu --set full -k regex:k_logistic_tc -s 2 -c 3 python scripts/gpu_prof_tc.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import turing_jl_b200 as T
model = T.models.config3_logistic()
eng = T.Engine(model, T.NUTS(0.8, init_eps=0.01), 4096, 0, seed=1)
eng.set_option("logistic_tc", 1)
eng.init()
dt = eng.bench_gradient(6)
print("k_logistic_tc full width: %.3f ms per gradient of 4096 chains" % (dt * 1e3))
eng.close()
