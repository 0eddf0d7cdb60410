"""profile target: the persistent tail kernel k_tc_persist (gradient + reduction + NUTS state machine
per leapfrog inside one cooperative launch) on BASELINE configs[2]'s data with 128 chains, i.e. the
regime of the last ~60 000 leapfrogs of the full job (one chain tile x 143 row chunks):
  ncu --set full -k regex:k_tc_persist -c 1 python scripts/gpu_prof_persist.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import turing_jl_b200 as T
model = T.models.config3_logistic()
e = T.Engine(model, T.NUTS(30, 0.8), 128, 30, seed=3)
e.set_option("logistic_tc", 1)
e.set_option("time_kernels", 1)
e.init()
e.run(6, 1, 1)
ks = e.kernel_times()
print("k_tc_persist: %d pumps in %.4f s = %.1f us per pump" % (ks[5], ks[4], ks[4] / ks[5] * 1e6))
e.close()
