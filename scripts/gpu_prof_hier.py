import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import turing_jl_b200 as T
model = T.models.config4_hier()
eng = T.Engine(model, T.NUTS(0.8), 296, 60, seed=1)
eng.init(); eng.run(120, 60, 1)
print("hier block run: %.3f s, %d grads" % (eng.L.tnuts_device_seconds(eng.h), eng.grad_evals))
eng.close()
