import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import turing_jl_b200 as T
N, D, C = 100000, 100, 4096
model = T.models.synthetic_logistic(N, D)
eng = T.Engine(model, T.NUTS(0.8, init_eps=0.01), C, 0, seed=1)
eng.init()
dt = eng.bench_gradient(3)
print("logistic gradient N=%d D=%d C=%d: %.3f ms -> %.2f TFLOP/s" % (N, D, C, dt * 1e3, 4.0 * N * D * C / dt / 1e12))
eng.close()
