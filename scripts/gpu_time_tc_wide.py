"""times the tcgen05 logistic gradient for 128 < D <= 256 (k_logistic_tc_wide, a cluster of two CTAs per
tile) next to the fp64 DMMA kernel, at the per-GPU shapes of BASELINE configs[4] (256 chains, D = 256;
the rows of one of 8 GPUs are 6.25e6 -- a slice of that is timed and scaled by rows)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import turing_jl_b200 as T
for (N, D, C) in ((1_000_000, 256, 256), (200_000, 256, 4096), (1_000_000, 192, 256)):
    model = T.models.synthetic_logistic(N, D)
    for tc in (1, 0):
        eng = T.Engine(model, T.NUTS(0.8, init_eps=0.01), C, 0, seed=1)
        eng.set_option("logistic_tc", tc)
        eng.init()
        dt = eng.bench_gradient(5)
        print("logistic gradient N=%d D=%d C=%d %s: %.3f ms -> %.1f TFLOP/s (4ND per chain)"
              % (N, D, C, "tcgen05 bf16x3 (CTA pair)" if tc else "fp64 DMMA", dt * 1e3, 4.0 * N * D * C / dt / 1e12), flush=True)
        eng.close()
