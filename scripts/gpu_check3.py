"""e2e timing breakdown (dev script)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import turing_jl_b200 as T
from turing_jl_b200 import pinned

model = T.EightSchools(); spl = T.NUTS(0.8); C = 65536; N = 1000
def tm(f, *a, **k):
    t = time.perf_counter(); r = f(*a, **k); return r, time.perf_counter() - t
for rep in range(3):
    e, t_create = tm(T.Engine, model, spl, C, 500, 1)
    _, t_init = tm(e.init)
    buf, t_pin = tm(pinned.empty, (N, 22, C))
    _, t_run = tm(e.run, 1499, 500, 1)
    _, t_fetch = tm(e.fetch_draws, buf)
    pg = np.empty((N, 22, C))
    _, t_fetch_pg = tm(e.fetch_draws, pg)
    _, t_stream = (None, 0)
    e.init()
    _, t_stream = tm(e.run_streamed, 1499, 500, 1, buf, 100)
    e.init()
    _, t_stream50 = tm(e.run_streamed, 1499, 500, 1, buf, 50)
    _, t_close = tm(e.close)
    print("rep %d: create %.3f init %.3f pinned-alloc %.3f run %.3f fetch(pinned) %.3f (%.1f GB/s) fetch(pageable) %.3f streamed(100) %.3f streamed(50) %.3f close %.3f" % (
        rep, t_create, t_init, t_pin, t_run, t_fetch, buf.nbytes / t_fetch / 1e9, t_fetch_pg, t_stream, t_stream50, t_close))
    del buf
for k in range(4):
    t = time.perf_counter()
    ch = T.sample(model, spl, T.MCMCThreads(), N, C, seed=k, verbose=False)
    dt = time.perf_counter() - t
    print("sample() wall %.3f s, grad evals %.3e -> %.3e /s" % (dt, ch.info["grad_evals"], ch.info["grad_evals"] / dt))
    del ch
