"""Full-job measurement of the big BASELINE.json configs (too long for the default bench.py run):
  python scripts/run_config.py logistic   # configs[2]: N=100k, D=100, 4096 chains, NUTS(0.8), 1000 draws
  python scripts/run_config.py logistic_tc  # the same job with the tcgen05 (bf16x3) likelihood kernel
  python scripts/run_config.py hier       # configs[3]: 1000 groups x 100 obs, D=2005, 1024 chains
Prints one JSON line (same keys as bench.py where they apply)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import turing_jl_b200 as T

def cpu_baseline_logistic(model, eps, ess_per_grad, budget_s=15.0):
    """CPU restatement (oracle/), one chain per thread on all host cores, bounded sample: chains
    started at the posterior mode, unit metric, the adapted step size rescaled to it (eps here is
    eps_adapted * sqrt(mean M^-1)), a few NUTS transitions each.  The
    full job (500 warm-up + 1000 draws from U(-2,2)) would take hours on the host cores, so
    min-bulk-ESS/s is extrapolated with the ESS per gradient of the GPU run (same algorithm)."""
    import oracle as O
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from helpers import oracle_of
    cores = os.cpu_count() or 1
    om = oracle_of(model)
    X, y = np.asarray(model.X).reshape(model.N, model.D), np.asarray(model.y)
    b = np.zeros(model.D)
    for _ in range(8):   # Newton steps to the mode
        p = 1.0 / (1.0 + np.exp(-X @ b))
        g = X.T @ (y - p) - b
        H = (X * (p * (1 - p))[:, None]).T @ X + np.eye(model.D)
        b = b + np.linalg.solve(H, g)
    cfg = O.make_config(delta=0.8, n_adapts=0, init_eps=float(eps), seed=5, sampler_mode=1)
    th0 = np.tile(b, (cores, 1))
    n_tr, dt, r = 1, 0.0, None
    while True:
        t0 = time.perf_counter()
        r = O.sample_chains(om, cfg, cores, n_tr, theta0=th0, n_threads=cores)
        dt = time.perf_counter() - t0
        if dt > budget_s / 3 or n_tr >= 256:
            break
        n_tr *= 2
    rate = r["n_grad"] / dt
    return {"value": rate, "unit": "grad_evals/s", "cores": cores, "kind": "port",
            "sample": "%d chains x %d NUTS transitions from the posterior mode at the adapted step size "
                      "(%d gradient evaluations, %.1f s), one chain per thread, analytic gradient"
                      % (cores, n_tr, r["n_grad"], dt),
            "ess_bulk_min_per_sec_extrapolated": rate * ess_per_grad}


def main():
    which = sys.argv[1]
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    N = 1000
    tc = which == "logistic_tc"
    if which in ("logistic", "logistic_tc"):
        model, C, name = T.models.config3_logistic(), int(4096 * scale), "BASELINE.json configs[2]: logistic regression N=100000 D=100 fp64"
        flops_per_grad = 4.0 * model.N * model.D
    else:
        model, C, name = T.models.config4_hier(), int(1024 * scale), "BASELINE.json configs[3]: hierarchical regression 1000 groups x 100 obs, D=2005"
        flops_per_grad = None
    spl = T.NUTS(0.8)
    nad = 500
    t0 = time.perf_counter()
    eng = T.Engine(model, spl, C, nad, seed=11)
    if tc:
        eng.set_option("logistic_tc", 1)
    t_create = time.perf_counter() - t0
    t0 = time.perf_counter(); eng.init(); t_init = time.perf_counter() - t0
    g_init = eng.grad_evals
    grad_ms = eng.bench_gradient(5) * 1e3
    out = T.pinned.empty((N, eng.ncol, C))
    t0 = time.perf_counter()
    eng.run_streamed(nad + N - 1, nad, 1, out, 100)
    t_run = time.perf_counter() - t0
    grads = eng.grad_evals
    t0 = time.perf_counter(); rhat, ess = eng.diagnostics(); t_diag = time.perf_counter() - t0
    P = model.D
    st = out[:, P + 3:, :]
    line = {
        "config": {"workload": name + ", NUTS(0.8), %d chains, %d draws + %d warm-up" % (C, N, nad)},
        "metric": "grad_evals_per_sec", "value": grads / (t_init + t_run), "unit": "grad_evals/s",
        "e2e_seconds": t_create + t_init + t_run, "init_seconds": t_init, "run_seconds": t_run,
        "grad_evals": grads, "grad_evals_init": g_init,
        "ess_bulk_min": float(ess.min()), "ess_bulk_min_per_sec": float(ess.min() / (t_create + t_init + t_run)),
        "rhat_max": float(rhat.max()), "diag_seconds": t_diag,
        "tree_depth_mean": float(st[:, 6].mean()), "tree_depth_max": float(st[:, 6].max()),
        "n_steps_mean": float(st[:, 0].mean()), "acceptance_mean": float(st[:, 1].mean()),
        "divergent": int(st[:, 7].sum()), "step_size_mean": float(eng.stepsize().mean()),
        "gradient_kernel_ms_all_chains": grad_ms, "gpu_launches": int(eng.launches),
    }
    if flops_per_grad:
        if tc:
            # executed tensor work: 9 bf16 plane products per (row, padded feature, chain):
            # 6 for eta (K = 16*ceil(D/16)) and 3 for X^T R (N = 16*ceil(D/16)), 2 flop each
            dpad = 16 * ((model.D + 15) // 16)
            peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
                if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {}
            peak = peaks.get("bf16_tflops", 1614.0)
            ach = 18.0 * model.N * dpad * C / (grad_ms * 1e-3) / 1e12
            line["roofline"] = {"bound": "tensor", "unit": "TFLOP/s", "achieved": ach, "peak": peak,
                                "peak_source": "MEASURED_PEAKS.json bf16 burst (kernel timed alone)" if peaks else "fallback",
                                "kernel": "k_logistic_tc", "frac": ach / peak,
                                "fp64_equivalent_tflops": flops_per_grad * C / (grad_ms * 1e-3) / 1e12,
                                "note": "achieved = executed bf16 tensor flops (6 plane products for eta, 3 for X^T R); fp64_equivalent = 4ND per chain"}
        else:
            line["roofline"] = {"bound": "tensor", "unit": "TFLOP/s", "achieved": flops_per_grad * C / (grad_ms * 1e-3) / 1e12,
                                "peak": 37.19, "peak_source": "measured here: mma.sync f64 (scripts/fp64_peak.cu); tcgen05 has no f64 kind",
                                "kernel": "k_logistic_mma"}
            line["roofline"]["frac"] = line["roofline"]["achieved"] / 37.19
        line["sustained_tflops_whole_job"] = grads * flops_per_grad / (t_init + t_run) / 1e12
    else:
        line["roofline"] = {"bound": "hbm", "unit": "GB/s", "achieved": 6 * model.D * 8 * grads / t_run / 1e9, "peak": 6545.9}
        line["roofline"]["frac"] = line["roofline"]["achieved"] / 6545.9
    if flops_per_grad and not os.environ.get("TNUTS_NO_CPU_BASELINE"):
        eps_unit = line["step_size_mean"] * float(np.sqrt(eng.inv_metric().mean()))
        line["cpu_baseline"] = cpu_baseline_logistic(model, eps_unit, float(ess.min()) / grads)
        line["speedup_vs_cpu_grad_evals"] = line["value"] / line["cpu_baseline"]["value"]
    print(json.dumps(line))
    eng.close()

if __name__ == "__main__":
    main()
