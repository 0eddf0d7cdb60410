#!/usr/bin/env python
"""bench.py -- headline benchmark of the many-chain NUTS engine (contract: DESIGN.md 'Measurement').

Default workload = BASELINE.json configs[2], the configuration the north star quotes its target on:
Bayesian logistic regression N=100 000, D=100, NUTS(0.8), 4096 chains per GPU, 500 warm-up + 1000
draws, likelihood on the tcgen05 tensor cores (engine option `logistic_tc`, bf16x3 planes).

  full job .. ONE whole sampling job through the public API with host buffers
              (`sample(model, NUTS(0.8), MCMCThreads(), 1000, 4096)`: model upload, init step, 1499
              transitions, the (iterations x names x chains) array copied to pinned host memory).
              Its wall time gives `ess_bulk_min_per_sec` (end to end) and its CUDA-event time the
              device figure; R-hat gate in the line.  Untimed by the step loop, reported as `full_job`.
  step ...... a block of --block post-warm-up NUTS transitions of ALL chains, restored from the state
              blob the full job saved (`save_state` / `initial_state`, src/mcmc/hmc.jl:135-152),
              every draw's parameters, log-probs and statistics recorded.
  value ..... gradient evaluations per second over the timed steps, all GPUs; state and model
              resident in HBM, kept draws left in HBM; CUDA events on the engine's stream, max over ranks
  e2e ....... the same step through the public API with HOST buffers:
              `sample(model, spl, MCMCThreads(), block, 4096; initial_state=state)` -- model and
              state uploaded, block run, draws copied back; wall clock, max over ranks
  roofline .. the tcgen05 gradient kernel's launches inside the timed steps (CUDA events on the
              engine's stream around a systematic sample of them, engine option `time_kernels`)
  secondary . the eight-schools job of BASELINE.json configs[1] (65 536 chains), a few steps

`--workload eight_schools` makes configs[1] the primary line (one step = one whole job).
`--impl reference` times the CPU restatement (oracle/, one chain per thread, all host cores) on
the same unit; the reference itself is Julia and cannot run in this image.
"""
import argparse
import hashlib
import json
import os
import socket
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALG_BYTES_PER_GRAD = {"eight_schools": 6 * 10 * 8, "linreg": 6 * 3 * 8, "gdemo": 6 * 2 * 8}
FP64_MMA_PEAK_TFLOPS = 37.19   # measured on this pool's B200s: scripts/fp64_peak.cu (mma.sync f64)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    d = {}
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
    return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_burst": d.get("bf16_tflops", 1590.0),
            "bf16_sustained": d.get("bf16_tflops_sustained", 1400.0),
            "source": "MEASURED_PEAKS.json" if d else "fallback of B200_PROFILING.md"}


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/)"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(kernel)
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # median of the upper half = clocks under load (idle samples sit at the floor)
        sm_load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(sm_load)) if sm_load else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def bind_to_gpu_numa_node(local):
    """Pin this process (and, by first touch, its pinned host buffers) to the NUMA node its GPU hangs
    off, otherwise half of the ranks copy across the socket interconnect.  Best effort."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def build_workload(name, n_chains):
    import turing_jl_b200 as T
    if name == "eight_schools":
        return T.EightSchools(), T.NUTS(0.8), 1000, n_chains or 65536
    if name == "linreg":
        return T.models.config1_linear_regression(), T.NUTS(), 1000, n_chains or 4
    if name == "logistic":
        return T.models.config3_logistic(), T.NUTS(0.8), 1000, n_chains or 4096
    raise SystemExit("unknown workload " + name)


def oracle_model(model):
    import oracle as O
    return O.OracleModel(model.family, D=model.D, N=model.N, G=model.G, X=model.X, y=model.y,
                         sigma=model.sigma, group=model.group, hyper=model.hyper)


def workload_config(name, C, N, block=None, tc=True):
    nad = min(1000, N // 2)
    if name == "eight_schools":
        return {"workload": "BASELINE.json configs[1]: eight schools non-centred, NUTS(0.8), "
                            "%d chains per GPU, %d draws + %d warm-up" % (C, N, nad),
                "chains_per_gpu": C, "draws": N, "D": 10, "step": "one whole job",
                "l2": "not flushed: each step writes %.1f GB of draws, >> 126 MB L2"
                      % (C * N * 22 * 8 / 1e9)}
    if name == "logistic":
        return {"workload": "BASELINE.json configs[2]: Bayesian logistic regression N=100000 D=100, "
                            "NUTS(0.8), %d chains per GPU, %d draws + %d warm-up" % (C, N, nad),
                "chains_per_gpu": C, "draws": N, "D": 100, "N": 100000,
                "likelihood": "tcgen05 bf16x3 planes (engine option logistic_tc)" if tc else "fp64 DMMA",
                "step": "block of %s post-warm-up NUTS transitions of all chains, restored from the "
                        "state blob of the full job (full job reported as full_job)" % block,
                "l2": "not flushed between steps: one step is hundreds of dependent gradient launches, "
                      "each streaming the 60 MB of bf16 X planes per chain tile (L2-resident by design) "
                      "and ~100 MB of chain-state arrays; a flush would touch the first launch only"}
    return {"workload": name, "chains_per_gpu": C, "draws": N}


# ----------------------------------------------------------------------------------------------
# CPU legs (oracle/): the reference arm and the cpu_baseline object
# ----------------------------------------------------------------------------------------------
def _cpu_cache_path(name, model, cores, chains):
    import oracle as O
    lib = O.build()
    h = hashlib.sha256()
    h.update(("%s|%d|%d|%d|%d|%s|%d" % (name, model.N, model.D, cores, chains, socket.gethostname(),
                                       int(os.path.getmtime(lib)))).encode())
    return os.path.join(tempfile.gettempdir(), "tnuts_cpu_baseline_%s" % h.hexdigest()[:16])


def _cpu_gradient_rate(om, model, cores, n_transitions=12):
    """gradient evaluations per second of the CPU job itself: `cores` chains, one per thread, take a
    few NUTS transitions from a posterior-width cloud around the mode with the step size and metric an
    adapted chain has (depth-3 trees) -- the sampler's own code path, threads and memory traffic.
    (A first version timed bare gradient calls at theta = 0 from a Python thread pool: on the 16-core GPU
    boxes that read 1454 gradients/s where the sampling job then ran at 605, and the schedule overshot
    its time budget three times over.)"""
    import oracle as O
    X, y, D = model.X, model.y, model.D
    b = np.zeros(D)
    for _ in range(6):                                   # Newton steps to the posterior mode
        p = 1.0 / (1.0 + np.exp(-(X @ b)))
        H = (X * (p * (1.0 - p))[:, None]).T @ X + np.eye(D)
        b = b + np.linalg.solve(H, X.T @ (y - p) - b)
    var = np.diag(np.linalg.inv(H))
    th = b + np.sqrt(var) * np.random.default_rng(7).standard_normal((cores, D))
    cfg = O.make_config(delta=0.8, n_adapts=0, seed=77, sampler_mode=1)
    eps, minv = np.full(cores, 0.38), np.tile(var, (cores, 1))
    O.resume_chains(om, cfg, th, eps, minv, 1000, 1, n_threads=cores)          # warm the threads up
    t = time.perf_counter()
    r = O.resume_chains(om, cfg, th, eps, minv, 1000, n_transitions, n_threads=cores)
    return r["n_grad"] / (time.perf_counter() - t)


def cpu_full_job(name, model, spl, N, use_cache=True, quick=False, budget_s=420.0):
    """A whole sampling schedule (warm-up + draws from U(-2,2)) on `cores` chains, one chain per thread
    on all host cores: MEASURED min bulk-ESS/s and grad-evals/s, nothing extrapolated.  Eight schools:
    seconds.  configs[2] is bound by the host's memory bandwidth (every gradient of every chain streams
    the 80 MB of X: 569 gradients/s on the 16 cores of this pool's GPU boxes), so the configured
    500 + 1000 schedule takes 24 minutes there (measured once: DESIGN.md section 5).  The reference arm
    has to end within minutes: the gradient rate is calibrated first and, if the configured schedule does
    not fit `budget_s`, warm-up and draws are shortened by the same factor (never below 150 + 300: Stan's
    window schedule needs 150 warm-up transitions) -- the line says which schedule ran.  The result (and
    the adapted state the step loop resumes from) is cached in the temp directory for the other arm /
    the other N of the same driver session on this box."""
    import oracle as O
    from oracle import diagnostics as dg
    cores = host_cores()
    heavy = model.family == 3 and model.N * model.D > 1_000_000
    nad = min(1000, N // 2)
    if quick:          # tests only: a short schedule
        N, nad = 40, 20
    chains = cores if heavy else 128 * cores
    path = _cpu_cache_path(name + ("_quick" if quick else "") + "_b%d" % int(budget_s), model, cores, chains)
    if use_cache and os.path.exists(path + ".json") and os.path.exists(path + ".npz"):
        with open(path + ".json") as f:
            out = json.load(f)
        st = np.load(path + ".npz")
        out["cached"] = "measured %.0f s ago on this box by an earlier bench.py run of this session" \
                        % (time.time() - out["measured_at"])
        return out, dict(theta=st["theta"], eps=st["eps"], minv=st["minv"])
    om = oracle_model(model)
    schedule = "the configured schedule"
    if heavy and not quick and budget_s > 0:
        rate = _cpu_gradient_rate(om, model, cores)
        # gradient evaluations per chain of this workload, from the two schedules measured on this pool's
        # boxes (500 + 1000: 50 717, 150 + 300: 18 124): the deep trees of the first warm-up transitions
        # cost ~4 200, every transition after that ~31
        grads = lambda a, n: 4200.0 + 31.0 * (a + n)
        # one chain per thread with as many chains as threads: the job lasts as long as its slowest chain
        # (measured: 8 chains kept 3.5 of 8 cores busy on average; wall time = 1.66 x total work / rate)
        straggler = 1.7
        est = straggler * chains * grads(nad, N) / rate
        if est > budget_s:
            # shorten warm-up and draws by the same factor so that the model fits the budget
            f = max(0.0, (budget_s * rate / (straggler * chains) - 4200.0) / (31.0 * (nad + N)))
            nad_s, N_s = max(150, int(round(nad * f))), max(300, int(round(N * f)))
            schedule = ("%d warm-up + %d draws instead of the configured %d + %d: calibrated at %.0f "
                        "gradients/s on %d cores the configured schedule would take %.0f s, the budget is "
                        "%.0f s" % (nad_s, N_s, nad, N, rate, cores, est, budget_s))
            nad, N = nad_s, N_s
    cfg = O.make_config(delta=spl.delta, max_depth=spl.max_depth, delta_max=spl.Delta_max,
                        n_adapts=nad, seed=1234, sampler_mode=1)
    n_tr = nad + N - 1
    t = time.perf_counter()
    r = O.sample_chains(om, cfg, chains, n_tr, first_keep=nad, n_threads=cores)
    dt = time.perf_counter() - t
    s = dg.summarize(r["params"])
    out = {
        "value": r["n_grad"] / dt, "unit": "grad_evals/s", "cores": cores, "kind": "port",
        "sample": "%d chains x (%d warm-up + %d draws) from U(-2,2), one chain per thread, analytic "
                  "gradient; %s" % (chains, nad, N, schedule),
        "seconds": dt, "ess_bulk_min": float(s["ess_bulk"].min()),
        "ess_bulk_min_per_sec": float(s["ess_bulk"].min() / dt),
        "rhat_max": float(s["rhat"].max()), "grad_evals": int(r["n_grad"]),
        "n_steps_mean": float(r["stats"][..., 0].mean()), "measured_at": time.time(),
        "iter0": int(n_tr), "chains": chains, "n_adapts": nad,
    }
    st = dict(theta=r["theta"][:, -1].copy(), eps=r["eps"], minv=r["minv"])
    try:
        np.savez(path + ".npz", **st)
        with open(path + ".json", "w") as f:
            json.dump(out, f)
    except OSError:
        pass
    out["cached"] = False
    return out, st


def cpu_block(model, spl, full, st, n_transitions):
    """one step of the CPU arm: n_transitions post-warm-up transitions of the adapted chains"""
    import oracle as O
    om = oracle_model(model)
    cfg = O.make_config(delta=spl.delta, max_depth=spl.max_depth, delta_max=spl.Delta_max,
                        n_adapts=full["n_adapts"], seed=1234, sampler_mode=1)
    t = time.perf_counter()
    r = O.resume_chains(om, cfg, st["theta"], st["eps"], st["minv"], full["iter0"], n_transitions,
                        n_threads=full["cores"])
    dt = time.perf_counter() - t
    return r["n_grad"], dt


def cpu_baseline_object(full, block_rate=None, block_desc=None):
    cb = {k: full[k] for k in ("value", "unit", "cores", "kind", "sample", "seconds",
                               "ess_bulk_min_per_sec", "rhat_max", "grad_evals", "cached")}
    cb["value_is"] = "whole-job grad-evals/s (warm-up included)"
    if block_rate is not None:
        cb["value"] = block_rate
        cb["value_is"] = "grad-evals/s over blocks of post-warm-up transitions (the step unit)"
        cb["sample"] = block_desc + "; ESS/s from: " + full["sample"]
        cb["whole_job_grad_evals_per_sec"] = full["value"]
    return cb


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import turing_jl_b200 as T  # noqa: F401  host descriptors only (no GPU is touched on this arm)
    model, spl, N, C = build_workload(args.workload, args.chains)
    quick = args.ref_quick
    full, st = cpu_full_job(args.workload, model, spl, N, use_cache=not args.no_cache, quick=quick,
                            budget_s=args.ref_budget)
    heavy = args.workload == "logistic"
    vals = []
    if heavy:
        kc = args.ref_block
        for i in range(args.warmup + args.steps):
            g, dt = cpu_block(model, spl, full, st, kc)
            if i >= args.warmup:
                vals.append((g, dt))
        desc = "%d adapted chains x %d post-warm-up NUTS transitions per step, one chain per thread" \
               % (full["chains"], kc)
    else:
        # eight schools: a step is the whole job (seconds on the host cores)
        for i in range(args.warmup + args.steps):
            f2, _ = cpu_full_job(args.workload, model, spl, N, use_cache=False, quick=quick, budget_s=0)
            if i >= args.warmup:
                vals.append((f2["grad_evals"], f2["seconds"]))
        desc = full["sample"]
    v = float(sum(g for g, _ in vals) / sum(dt for _, dt in vals))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    cb = cpu_baseline_object(full, v, desc) if heavy else cpu_baseline_object(full)
    cb["value"] = v
    line = {
        "impl": "reference", "metric": "grad_evals_per_sec", "value": v, "unit": "grad_evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.workload, C, N, args.block, tc=not args.fp64),
        "ess_bulk_min_per_sec": full["ess_bulk_min_per_sec"],
        "rhat_max": full["rhat_max"],
        "full_job": {"seconds": full["seconds"], "grad_evals": full["grad_evals"],
                     "grad_evals_per_sec": full["value"], "chains": full["chains"],
                     "ess_bulk_min_per_sec": full["ess_bulk_min_per_sec"], "cached": full["cached"]},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": v, "unit": "grad_evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement (oracle/) of the reference algorithm on all host cores, one chain per "
                "thread (the MCMCThreads fan-out), analytic gradient -- an upper bound on Turing's "
                "ForwardDiff path; Turing.jl itself needs Julia, absent from this image",
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class Ctx:
    pass


def bench_eight_schools(cx, steps, n_warm, n_chains=0):
    """configs[1]: one step = one whole job (init + 1499 transitions, every kept draw written)"""
    T = cx.T
    model, spl, N, C = build_workload("eight_schools", n_chains)
    nad = min(1000, N // 2)
    n_tr = nad + N - 1
    seed = 20261019
    eng = T.Engine(model, spl, C, nad, seed, device=cx.local, chain_offset=cx.rank * C)
    if cx.dist is not None:
        T.distributed.attach_comm(eng, cx.dist, device="cuda:%d" % cx.local)

    def step():
        eng.init()
        eng.run(n_tr, nad, 1)
        return eng.L.tnuts_init_device_seconds(eng.h), eng.L.tnuts_device_seconds(eng.h)

    for _ in range(n_warm):
        step()
    cx.barrier()
    l0 = eng.launches
    dev_init = dev_run = 0.0
    grads = 0
    for _ in range(steps):
        a, b = step()
        dev_init += a
        dev_run += b
        grads += eng.grad_evals
    cx.barrier()
    launches = eng.launches - l0
    dev_total = cx.allmax(dev_init + dev_run)
    grads_all = cx.allsum(grads)
    rhat, ess = eng.diagnostics()
    div = eng.divergences
    eng.close()

    def e2e_step(k):
        ch = T.sample(model, spl, T.MCMCThreads(), N, C, seed=seed + k, devices=[cx.local],
                      chain_offset=cx.rank * C, verbose=False)
        res = (ch.info["grad_evals"], ch.info["gpu_launches"], ch.value.nbytes, dict(ch.info["host_timing"]),
               ch.info["device_seconds"])
        del ch
        return res

    for k in range(n_warm):
        e2e_step(k)
    cx.barrier()
    t0 = time.perf_counter()
    g_e2e = l_e2e = 0
    phases = []
    for k in range(steps):
        ge, le, d2h, tm, dsec = e2e_step(100 + k)
        g_e2e += ge
        l_e2e += le
        tm["device"] = dsec
        phases.append(tm)
    cx.barrier()
    wall_e2e = cx.allmax(time.perf_counter() - t0)
    hbm = cx.peaks["hbm_gbs"]
    alg = ALG_BYTES_PER_GRAD["eight_schools"]
    achieved = alg * grads / dev_run / 1e9
    return {
        "value": grads_all / dev_total, "unit": "grad_evals/s", "ms_per_step": dev_total / steps * 1e3,
        "steps": steps, "warmup": n_warm,
        "config": workload_config("eight_schools", C, N),
        "ess_bulk_min_per_sec": float(ess.min()) / (dev_total / steps), "rhat_max": float(rhat.max()),
        "divergent_per_gpu": int(div), "grad_evals_per_step": grads_all / steps,
        "e2e": {"value": cx.allsum(g_e2e) / wall_e2e, "unit": "grad_evals/s",
                "h2d_bytes_per_step": int(model.data_nbytes() + 256), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": wall_e2e / steps * 1e3,
                "ess_bulk_min_per_sec": float(ess.min()) / (wall_e2e / steps),
                "host_timing_ms": {k: float(np.mean([p[k] for p in phases])) * 1e3 for k in phases[0]},
                "api": "turing_jl_b200.sample(model, NUTS(0.8), MCMCThreads(), N, n_chains)"},
        "gpu_launches": int(launches + l_e2e),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                     "frac": achieved / hbm, "traffic": ncu_traffic("k_chain_run_dyn"),
                     "peak_source": cx.peaks["source"],
                     "kernel": "k_chain_run_dyn<EightSchools<8>> (k_chain_run when the grid fits one wave)",
                     "algorithmic_bytes_per_grad_eval": alg,
                     "note": "SURVEY 8(d) streaming figure 6*D*8 B per leapfrog; the kernel keeps the "
                             "state in registers, so it is fp64-issue bound, see DESIGN.md"},
    }


def bench_logistic(cx, args):
    T = cx.T
    model, spl, N, C = build_workload("logistic", args.chains)
    nad = min(1000, N // 2)
    K = args.block
    seed = 20261019
    tc = not args.fp64
    opts = {"logistic_tc": 1} if tc else {}
    n_warm = max(args.warmup, 3)

    # ---- the whole job, once, through the public API (host buffers) ----
    cx.barrier()
    t0 = time.perf_counter()
    ch = T.sample(model, spl, T.MCMCThreads(), N, C, seed=seed, devices=[cx.local],
                  chain_offset=cx.rank * C, verbose=False, engine_options=dict(opts, time_kernels=1),
                  save_state=True, keep_engine=True)
    cx.barrier()
    full_wall = cx.allmax(time.perf_counter() - t0)
    eng = ch.info["engines"][0]
    fk = eng.kernel_times()     # the gradient kernel's launches / the persistent tail inside the full job
    full_dev_rank = ch.info["device_seconds"]
    full_dev = cx.allmax(ch.info["device_seconds"])
    full_grads = cx.allsum(ch.info["grad_evals"])
    full_launches = ch.info["gpu_launches"]
    state = ch.info["samplerstate"]
    blob = state["blobs"][0]
    if cx.dist is not None:
        T.distributed.attach_comm(eng, cx.dist, device="cuda:%d" % cx.local)
    rhat, ess = eng.diagnostics()       # multi-GPU: pooled over all ranks' chains (NCCL exchange)
    stats = {"n_steps_mean": float(np.mean(ch["n_steps"])), "tree_depth_mean": float(np.mean(ch["tree_depth"])),
             "acceptance_mean": float(np.mean(ch["acceptance_rate"])),
             "step_size_mean": float(np.mean(ch.info["step_size"])), "divergent_per_gpu": int(ch.info["n_divergent"]),
             "host_timing_ms": {k: v * 1e3 for k, v in ch.info["host_timing"].items()}}
    full_d2h = ch.value.nbytes
    del ch

    # ---- arm 1: device-resident steps ----
    eng.set_option("time_kernels", 1)
    grad_ms_full_width = eng.bench_gradient(5) * 1e3

    def step():
        eng.set_state(blob)              # untimed: the inputs are resident when the timed region starts
        g0 = eng.grad_evals
        eng.run(K, 1, 1)
        ks, kl, ksm, kt, ps, pp = eng.kernel_times()
        return eng.L.tnuts_device_seconds(eng.h), eng.grad_evals - g0, ks, kl, kt, g0, ps, pp

    for _ in range(n_warm):
        blob_grads = step()[5]   # evaluations the restored chains had done (the counter is in the blob)
    clocks = ClockSampler(cx.local)
    if cx.rank == 0:
        clocks.start()
    cx.barrier()
    l0 = eng.launches
    dev = ksec = psec = 0.0
    grads = klaunch = ktiles = ppumps = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        d, g, ks, kl, kt, _, ps, pp = step()
        dev += d
        grads += g
        ksec += ks
        klaunch += kl
        ktiles += kt
        psec += ps
        ppumps += pp
    cx.barrier()
    wall_value = time.perf_counter() - t0
    launches = eng.launches - l0
    dev_max = cx.allmax(dev)
    grads_all = cx.allsum(grads)
    value = grads_all / dev_max
    eng.close()

    # ---- arm 2: the same step end to end through the public API (host buffers) ----
    def e2e_step():
        c2 = T.sample(model, spl, T.MCMCThreads(), K, C, initial_state=state, devices=[cx.local],
                      chain_offset=cx.rank * C, verbose=False, engine_options=opts)
        res = (c2.info["grad_evals"] - blob_grads, c2.info["gpu_launches"], c2.value.nbytes,
               float(c2.value[-1, 0, 0]), dict(c2.info["host_timing"]), c2.info["device_seconds"])
        del c2
        return res

    for _ in range(n_warm):
        e2e_step()
    cx.barrier()
    t0 = time.perf_counter()
    g_e2e = l_e2e = 0
    phases = []
    for _ in range(args.steps):
        ge, le, d2h, _, tm, dsec = e2e_step()
        g_e2e += ge
        l_e2e += le
        tm["device"] = dsec
        phases.append(tm)
    cx.barrier()
    wall_e2e = cx.allmax(time.perf_counter() - t0)
    e2e_value = cx.allsum(g_e2e) / wall_e2e
    clk = clocks.stop() if cx.rank == 0 else None

    Dm, Nr = model.D, model.N
    dpad = 16 * ((Dm + 15) // 16)
    gsec = ksec + psec      # time of the kernels that evaluate gradients
    useful = 4.0 * Nr * Dm * grads / gsec / 1e12 if gsec > 0 else None
    if tc:
        executed = 18.0 * Nr * dpad * 128.0 * ktiles / ksec / 1e12 if ksec > 0 else None
        pk = cx.peaks["bf16_sustained"]
        # the dominant kernel: k_logistic_tc (58 % of the full job's device time, every gradient of the
        # full-width and mid-width pumps).  ALGORITHMIC flops per launch (4*N*D per gradient
        # evaluation x the 4096 chains one full-width launch serves) / its launch time measured here
        # with CUDA events on the engine's stream
        fw_useful = 4.0 * Nr * Dm * C / (grad_ms_full_width * 1e-3) / 1e12
        fw_exec = 18.0 * Nr * dpad * 128.0 * ((C + 127) // 128) / (grad_ms_full_width * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": fw_useful, "peak": pk, "unit": "TFLOP/s",
                "frac": fw_useful / pk,
                "traffic": ncu_traffic("k_logistic_tc"),
                "kernel": "k_logistic_tc<128>, one full-width launch (%d chains x %d rows)" % (C, Nr),
                "peak_source": cx.peaks["source"] + " bf16 sustained (the kernel is timed inside a long step)",
                "algorithmic_flops_per_launch": 4.0 * Nr * Dm * C,
                "launch_ms": grad_ms_full_width,
                "executed_bf16_tflops": fw_exec, "executed_frac": fw_exec / pk,
                "executed_frac_of_burst_peak": fw_exec / cx.peaks["bf16_burst"],
                "x_fp64_mma_peak": fw_useful / FP64_MMA_PEAK_TFLOPS,
                "share_of_full_job": fk[0] / full_dev_rank if full_dev_rank > 0 else None,
                "share_of_step": ksec / dev,
                "in_step": {"launches_per_step": klaunch / args.steps, "seconds_per_step": ksec / args.steps,
                            "executed_bf16_tflops": executed,
                            "executed_frac": executed / pk if executed else None},
                "tail_kernel": {"kernel": "k_tc_persist<128> (gradient + reduction + NUTS state machine per "
                                          "leapfrog of the <= 512 unfinished chains, one cooperative launch)",
                                "bound": "latency (one leapfrog of a few chains at a time: 148 CTAs x 11-43 row "
                                         "tiles, three grid barriers per leapfrog)",
                                "seconds_per_step": psec / args.steps, "pumps_per_step": ppumps / args.steps,
                                "us_per_pump": psec / ppumps * 1e6 if ppumps else None,
                                "share_of_step": psec / dev,
                                "share_of_full_job": fk[4] / full_dev_rank if full_dev_rank > 0 else None,
                                "pumps_in_full_job": int(fk[5])},
                "step_kernels": {"useful_tflops": useful, "frac": useful / pk if useful else None,
                                 "seconds_per_step": gsec / args.steps, "share_of_step": gsec / dev,
                                 "note": "all gradient evaluations of the timed steps x 4*N*D / the time of both "
                                         "kernels (the tail kernel's time includes its reductions and state machine)"},
                "note": "achieved counts the SURVEY 8(d) algorithmic work (4*N*D flops per gradient "
                        "evaluation); the kernel evaluates the fp64 contraction as 9 bf16 plane products (6 for "
                        "eta = X.B, 3 for X^T R) on tiles padded to 128 chains x 112 features: "
                        "executed_bf16_tflops counts those; x_fp64_mma_peak compares the useful rate with the "
                        "37.19 TFLOP/s fp64 tensor peak measured on this chip (tcgen05 has no f64 kind)"}
    else:
        roof = {"bound": "tensor", "achieved": useful, "peak": FP64_MMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": useful / FP64_MMA_PEAK_TFLOPS if useful else None, "traffic": ncu_traffic("k_logistic_mma"),
                "kernel": "k_logistic_mma",
                "peak_source": "measured here: mma.sync f64, scripts/fp64_peak.cu (tcgen05 has no f64 kind)",
                "algorithmic_flops_per_grad_eval": 4.0 * Nr * Dm,
                "kernel_share_of_step": gsec / dev}
    ess_min = float(ess.min())
    line = {
        "metric": "grad_evals_per_sec", "value": value, "unit": "grad_evals/s",
        "n_gpus": cx.world, "steps": args.steps, "warmup": n_warm,
        "ms_per_step": dev_max / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 planes, fp32 accumulate, fp64 log-likelihood and state" if tc else "f64",
        "data": "synthetic",
        "config": workload_config("logistic", C, N, K, tc),
        "ess_bulk_min_per_sec": ess_min / full_dev, "rhat_max": float(rhat.max()),
        "grad_evals_per_step": grads_all / args.steps,
        "wall_ms_per_step": wall_value / args.steps * 1e3,
        "full_job": {"device_seconds": full_dev, "e2e_seconds": full_wall, "grad_evals": full_grads,
                     "grad_evals_per_sec_device": full_grads / full_dev,
                     "grad_evals_per_sec_e2e": full_grads / full_wall,
                     "ess_bulk_min": ess_min, "ess_bulk_min_per_sec_device": ess_min / full_dev,
                     "ess_bulk_min_per_sec_e2e": ess_min / full_wall, "rhat_max": float(rhat.max()),
                     "d2h_bytes": int(full_d2h) * cx.world, "gpu_launches_per_gpu": int(full_launches),
                     "fp64_equivalent_tflops_sustained": 4.0 * Nr * Dm * full_grads / full_dev / 1e12 / cx.world,
                     "k_logistic_tc_seconds": fk[0], "k_logistic_tc_launches": int(fk[1]),
                     "k_tc_persist_seconds": fk[4], "k_tc_persist_pumps": int(fk[5]),
                     "api": "turing_jl_b200.sample(model, NUTS(0.8), MCMCThreads(), 1000, %d; save_state=true)" % C,
                     **stats},
        "e2e": {"value": e2e_value, "unit": "grad_evals/s",
                "h2d_bytes_per_step": int(model.data_nbytes() + blob.nbytes),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": wall_e2e / args.steps * 1e3,
                "ess_bulk_min_per_sec": ess_min / full_wall,
                "host_timing_ms": {k: float(np.mean([p[k] for p in phases])) * 1e3 for k in phases[0]},
                "api": "turing_jl_b200.sample(model, NUTS(0.8), MCMCThreads(), %d, %d; initial_state=state)" % (K, C)},
        "gpu_launches": int(launches + l_e2e),
        "roofline": roof,
        "clocks": clk,
    }
    return line, model, spl, N


def bench_wide_gradient(cx, rows=1_000_000, Dm=256, C=256, reps=6):
    """BASELINE configs[4]'s per-GPU width (D = 256, 256 chains; 1e6 of a GPU's 6.25e6 rows): one
    likelihood gradient of all chains on the tcgen05 CTA-pair kernel (k_logistic_tc_wide: a cluster of
    two CTAs per tile, partial X.B exchanged through distributed shared memory) and on the fp64 DMMA
    kernel, CUDA events on the engine's stream (tnuts_bench_gradient)."""
    T = cx.T
    model = T.models.synthetic_logistic(rows, Dm, seed=4)
    out = {"workload": "logistic likelihood gradient, %d rows x D = %d, %d chains (BASELINE configs[4] per-GPU "
                       "width, a slice of its rows)" % (rows, Dm, C), "gpu_launches": 0}
    for name, tc in (("tcgen05_cta_pair", 1), ("fp64_dmma", 0)):
        eng = T.Engine(model, T.NUTS(0.8, init_eps=0.01), C, 0, seed=1, device=cx.local)
        eng.set_option("logistic_tc", tc)
        eng.init()
        l0 = eng.launches
        dt = eng.bench_gradient(reps if tc else 2)
        out["gpu_launches"] += int(eng.launches - l0)
        out[name] = {"ms_per_gradient": dt * 1e3, "useful_tflops": 4.0 * rows * Dm * C / dt / 1e12}
        eng.close()
    pk = cx.peaks["bf16_sustained"]
    out["tcgen05_cta_pair"]["frac_of_bf16_peak"] = out["tcgen05_cta_pair"]["useful_tflops"] / pk
    out["tcgen05_cta_pair"]["traffic"] = ncu_traffic("k_logistic_tc_wide")
    out["tcgen05_cta_pair"]["algorithmic_bytes"] = 3 * rows * Dm * 2
    out["speedup_over_fp64_dmma"] = out["fp64_dmma"]["ms_per_gradient"] / out["tcgen05_cta_pair"]["ms_per_gradient"]
    return out


def run_b200(args):
    import torch
    import turing_jl_b200 as T
    cx = Ctx()
    cx.T = T
    cx.world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.rank = int(os.environ.get("RANK", "0"))
    cx.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the engine has no CPU fallback")
    torch.cuda.set_device(cx.local)
    numa = bind_to_gpu_numa_node(cx.local) if cx.world > 1 else None
    cx.dist = None
    if cx.world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", cx.local))
        cx.dist = dist
    cx.peaks = measured_peaks()

    def barrier():
        if cx.dist is not None:
            cx.dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if cx.dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        cx.dist.all_reduce(t, op=cx.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if cx.dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        cx.dist.all_reduce(t, op=cx.dist.ReduceOp.SUM)
        return float(t.item())

    cx.barrier, cx.allmax, cx.allsum = barrier, allmax, allsum

    # multi-GPU: equality of the row-sharded / mesh / pooled-diagnostics modes with a single-GPU
    # engine, before anything is timed
    parity = None
    if cx.world > 1 and not args.no_parity_checks:
        from turing_jl_b200 import multi_gpu_checks
        parity = multi_gpu_checks.run_all(cx.dist, cx.rank, cx.world, cx.local)

    if args.workload == "logistic":
        line, model, spl, N = bench_logistic(cx, args)
        if not args.no_secondary:
            sec = bench_eight_schools(cx, 3, 3)
            line["gpu_launches"] += sec["gpu_launches"]
            line["secondary"] = sec
            if cx.rank == 0:
                line["wide_kernel"] = bench_wide_gradient(cx)
                line["gpu_launches"] += line["wide_kernel"].pop("gpu_launches")
    else:
        n_warm = max(args.warmup, 3)
        clocks = ClockSampler(cx.local)
        if cx.rank == 0:
            clocks.start()
        sec = bench_eight_schools(cx, args.steps, n_warm, args.chains)
        clk = clocks.stop() if cx.rank == 0 else None
        model, spl, N, _ = build_workload("eight_schools", args.chains)
        line = {"metric": "grad_evals_per_sec", "value": sec.pop("value"), "unit": sec.pop("unit"),
                "n_gpus": cx.world, "steps": args.steps, "warmup": n_warm,
                "ms_per_step": sec.pop("ms_per_step"), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
        sec.pop("steps"), sec.pop("warmup")
        line.update(sec)
        line["clocks"] = clk
    if cx.rank == 0:
        if cx.world > 1:
            line["e2e"]["host_numa_binding"] = ("rank pinned to its GPU's NUMA node" if numa is not None
                                                else "none (single NUMA node or not exposed)")
        if parity is not None:
            line["parity_checks"] = parity
        if not args.no_cpu_baseline and cx.world == 1:
            full, st = cpu_full_job(args.workload, model, spl, N, use_cache=not args.no_cache,
                                    budget_s=args.ref_budget)
            if args.workload == "logistic":
                vals = [cpu_block(model, spl, full, st, args.ref_block) for _ in range(3)]
                rate = sum(g for g, _ in vals[1:]) / sum(dt for _, dt in vals[1:])
                line["cpu_baseline"] = cpu_baseline_object(
                    full, rate, "%d adapted chains x %d post-warm-up NUTS transitions (x2 after one warm-up "
                                "block), one chain per thread" % (full["chains"], args.ref_block))
            else:
                line["cpu_baseline"] = cpu_baseline_object(full)
        print(json.dumps(line))
    if cx.dist is not None:
        cx.dist.barrier()
        cx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="logistic", choices=["logistic", "eight_schools"])
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the workload's)")
    ap.add_argument("--block", type=int, default=20, help="logistic: post-warm-up transitions per step")
    ap.add_argument("--ref-block", type=int, default=2, help="CPU legs: post-warm-up transitions per step")
    ap.add_argument("--ref-budget", type=float, default=420.0,
                    help="CPU legs: seconds the measured sampling schedule may take (0: the configured schedule "
                         "whatever it costs -- 24 minutes for configs[2] on 16 cores)")
    ap.add_argument("--ref-quick", action="store_true", help="tests: a short CPU schedule")
    ap.add_argument("--fp64", action="store_true", help="logistic: fp64 DMMA likelihood instead of tcgen05")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-parity-checks", action="store_true")
    ap.add_argument("--no-cache", action="store_true", help="CPU legs: measure again even if a result of "
                                                            "this session is cached in the temp directory")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


def _only_json_on_stdout():
    """Libraries below us write to the C-level stdout (NCCL prints its version banner there when
    NCCL_DEBUG=VERSION is in the environment).  The contract is ONE JSON line on stdout: point fd 1
    at stderr for the duration of the run and hand the real stdout to `print` only."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real


if __name__ == "__main__":
    _only_json_on_stdout()
    main()
    sys.stdout.flush()
