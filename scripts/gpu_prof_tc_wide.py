"""profile target: the tcgen05 CTA-pair kernel (k_logistic_tc_wide) at BASELINE configs[4]'s per-GPU
width (D = 256, 256 chains; 1e6 of the GPU's 6.25e6 rows):
  ncu --set full -k regex:k_logistic_tc_wide -s 2 -c 1 python scripts/gpu_prof_tc_wide.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import turing_jl_b200 as T
model = T.models.synthetic_logistic(1_000_000, 256)
eng = T.Engine(model, T.NUTS(0.8, init_eps=0.01), 256, 0, seed=1)
eng.set_option("logistic_tc", 1)
eng.init()
dt = eng.bench_gradient(4)
print("k_logistic_tc_wide: %.3f ms per gradient of 256 chains" % (dt * 1e3))
eng.close()
