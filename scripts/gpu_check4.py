import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import turing_jl_b200 as T
model = T.EightSchools(); spl = T.NUTS(0.8); C = 65536; N = 1000
for k in range(2):
    ch = T.sample(model, spl, T.MCMCThreads(), N, C, seed=k, verbose=False); del ch
pr = cProfile.Profile(); pr.enable()
for k in range(3):
    t = time.perf_counter()
    ch = T.sample(model, spl, T.MCMCThreads(), N, C, seed=10 + k, verbose=False)
    print("wall", time.perf_counter() - t); del ch
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
