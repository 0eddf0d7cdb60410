"""profile target: the per-chain state-machine kernel k_ls_pump at full width (config 3 shapes,
tcgen05 gradient), a few hundred pumps:  ncu -k regex:k_ls_pump -s 200 -c 1 python scripts/gpu_prof_pump.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import turing_jl_b200 as T
model = T.models.config3_logistic()
eng = T.Engine(model, T.NUTS(0.8), 4096, 500, seed=11)
eng.set_option("logistic_tc", 1)
eng.init()
eng.run(30, 0, 1)         # thirty warm-up transitions: some thousand pumps
print("launches", eng.launches)
eng.close()
