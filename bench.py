#!/usr/bin/env python
"""bench.py -- headline benchmark of the many-chain NUTS engine (contract: see DESIGN.md 'Measurement').

One "step" = one whole sampling job of the workload, i.e. what
`sample(model, NUTS(0.8), MCMCThreads(), 1000, n_chains)` does: init step (initial parameters,
step-size search), 500 warm-up transitions with windowed adaptation, 1000 kept transitions with
constrained parameters, log-probs and sampler statistics recorded.

  value ..... gradient evaluations per second, all GPUs, model resident in HBM, kept draws left
              in HBM; timed with CUDA events on the engine's stream (max over ranks)
  e2e ....... the same metric through the public API with HOST buffers: model upload, init, run,
              and the device->host copy of the full (iterations x names x chains) chain array
  min bulk-ESS/s is reported next to both (ess_bulk_min_per_sec).

`--workload logistic [--tc]` runs BASELINE.json configs[2] (logistic regression N=100000, D=100,
4096 chains) instead of the default configs[1]; one step of it takes about a minute (--tc: the
tcgen05 bf16x3 likelihood kernel) or eight (fp64 DMMA), so give it `--steps 1 --warmup 1`.

`--impl reference` times the CPU restatement (oracle/, one chain per thread, all host cores):
the reference itself is Julia and cannot run in this image.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALG_BYTES_PER_GRAD = {"eight_schools": 6 * 10 * 8, "linreg": 6 * 3 * 8, "gdemo": 6 * 2 * 8}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


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
    """Multi-GPU runs move 11.5 GB of draws per GPU and step to host memory: pin this process (and,
    by first touch, its pinned host buffers) to the NUMA node its GPU hangs off, otherwise half of
    the ranks copy across the socket interconnect.  Best effort; returns the node or None."""
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


def cpu_baseline_logistic(model, budget_s=15.0):
    """config 3 on the host cores: the full job would take hours, so the bounded sample is a few
    NUTS transitions per core from the posterior mode (scripts/run_config.py has the details)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("run_config", os.path.join(ROOT, "scripts", "run_config.py"))
    rc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rc)
    eps_unit = 0.384 * 2.0 / np.sqrt(model.N)   # adapted step size x posterior sd (unit metric)
    t = time.perf_counter()
    out = rc.cpu_baseline_logistic(model, eps_unit, 0.0, budget_s)
    out["seconds"] = time.perf_counter() - t
    out.pop("ess_bulk_min_per_sec_extrapolated", None)
    out["ess_bulk_min_per_sec"] = None
    return out


def cpu_baseline(model, spl, N, budget_s=15.0):
    """oracle (CPU restatement), one chain per thread on all host cores, bounded sample"""
    if model.family == 3 and model.N * model.D > 1_000_000:
        return cpu_baseline_logistic(model, budget_s)
    import oracle as O
    from oracle import diagnostics as dg
    cores = os.cpu_count() or 1
    om = oracle_model(model)
    nad = min(1000, N // 2)
    cfg = O.make_config(delta=spl.delta, max_depth=spl.max_depth, delta_max=spl.Delta_max,
                        n_adapts=nad, seed=1234, sampler_mode=1)
    n_tr = nad + N - 1
    chains = 2 * cores
    t = time.perf_counter()
    r = O.sample_chains(om, cfg, chains, n_tr, first_keep=nad, n_threads=cores)
    dt = time.perf_counter() - t
    # scale the sample to the budget (bounded above so that the ESS estimate stays cheap)
    scale = int(max(1, min(64, budget_s / max(dt, 1e-3))))
    if scale > 1:
        chains = chains * scale
        t = time.perf_counter()
        r = O.sample_chains(om, cfg, chains, n_tr, first_keep=nad, n_threads=cores)
        dt = time.perf_counter() - t
    s = dg.summarize(r["params"])
    return {
        "value": r["n_grad"] / dt, "unit": "grad_evals/s", "cores": cores, "kind": "port",
        "sample": "%d chains x (%d warm-up + %d draws), one chain per thread, analytic gradient"
                  % (chains, nad, N),
        "seconds": dt, "ess_bulk_min_per_sec": float(s["ess_bulk"].min() / dt),
        "rhat_max": float(s["rhat"].max()), "grad_evals": r["n_grad"],
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import turing_jl_b200 as T  # host descriptors only (no GPU is touched on this arm)
    model, spl, N, C = build_workload(args.workload, args.chains)
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_baseline(model, spl, N, budget_s=args.ref_budget)
        if i >= args.warmup:
            vals.append(last)
    v = float(np.mean([x["value"] for x in vals]))
    ms = float(np.mean([x["seconds"] for x in vals])) * 1e3
    for x in vals:
        if x.get("ess_bulk_min_per_sec") is None:
            x["ess_bulk_min_per_sec"] = float("nan")
    line = {
        "impl": "reference", "metric": "grad_evals_per_sec", "value": v, "unit": "grad_evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.workload, C, N),
        "ess_bulk_min_per_sec": float(np.mean([x["ess_bulk_min_per_sec"] for x in vals])),
        "cpu_baseline": {k: last[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": v, "unit": "grad_evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement (oracle/) of the reference algorithm, all host cores, bounded "
                "sample of the workload; Turing.jl itself needs Julia, absent from this image",
    }
    line["cpu_baseline"]["value"] = v
    print(json.dumps(line))


def workload_config(name, C, N):
    if name == "eight_schools":
        return {"workload": "BASELINE.json configs[1]: eight schools non-centred, NUTS(0.8), "
                            "%d chains per GPU, %d draws + %d warm-up" % (C, N, min(1000, N // 2)),
                "chains_per_gpu": C, "draws": N, "D": 10,
                "l2": "not flushed: each step writes %.1f GB of draws, >> 126 MB L2"
                      % (C * N * 22 * 8 / 1e9)}
    if name == "logistic":
        return {"workload": "BASELINE.json configs[2]: Bayesian logistic regression N=100000 D=100, "
                            "NUTS(0.8), %d chains per GPU, %d draws + %d warm-up" % (C, N, min(1000, N // 2)),
                "chains_per_gpu": C, "draws": N, "D": 100,
                "l2": "not flushed: every gradient streams the 60 MB of X planes per chain tile and "
                      "each step writes %.1f GB of draws" % (C * N * 112 * 8 / 1e9)}
    return {"workload": name, "chains_per_gpu": C, "draws": N}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="eight_schools")
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: workload's)")
    ap.add_argument("--ref-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tc", action="store_true",
                    help="logistic workload: tcgen05 bf16x3 likelihood kernel (engine option logistic_tc)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import turing_jl_b200 as T
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    model, spl, N, C = build_workload(args.workload, args.chains)
    nad = min(1000, N // 2)
    n_tr = nad + N - 1
    seed = 20261019

    # ---------------- arm 1: device-resident (value) ----------------
    opts = {"logistic_tc": 1} if args.tc else {}
    eng = T.Engine(model, spl, C, nad, seed, device=local, chain_offset=rank * C)
    for k_, v_ in opts.items():
        eng.set_option(k_, v_)
    if dist is not None:
        # chains shard over the ranks; the communicator is only used by the pooled diagnostics
        T.distributed.attach_comm(eng, dist, device="cuda:%d" % local)

    def step():
        eng.init()
        eng.run(n_tr, nad, 1)
        return eng.L.tnuts_init_device_seconds(eng.h), eng.L.tnuts_device_seconds(eng.h)

    n_warm = max(args.warmup, 3) if args.workload != "logistic" else max(args.warmup, 1)
    for _ in range(n_warm):
        step()
    grad_ms = eng.bench_gradient(5) * 1e3 if args.workload == "logistic" else None
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    l0 = eng.launches
    g0 = 0
    dev_init = dev_run = 0.0
    grads = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        a, b = step()
        dev_init += a
        dev_run += b
        grads += eng.grad_evals
    barrier()
    wall_value = time.perf_counter() - t0
    launches = eng.launches - l0
    dev_total = allmax(dev_init + dev_run)
    dev_run_max = allmax(dev_run)
    grads_all = allsum(grads)
    value = grads_all / dev_total
    # diagnostics of the last step's draws (outside the timed region: measurement, not sampling)
    rhat, ess = eng.diagnostics()   # multi-GPU: pooled over all ranks' chains (NCCL exchange)
    ess_min_all = float(ess.min())
    rhat_max = float(rhat.max())
    ess_per_sec = ess_min_all / (dev_total / args.steps)
    div = eng.divergences
    eng.close()

    # ---------------- arm 2: end to end through the public API (host buffers) ----------------
    def e2e_step(k):
        # the user-facing call: host model in, host chain array (iterations x names x chains) out
        ch = T.sample(model, spl, T.MCMCThreads(), N, C, seed=seed + k, devices=[local],
                      chain_offset=rank * C, verbose=False, engine_options=opts)
        res = (ch.info["grad_evals"], ch.info["gpu_launches"], ch.value.nbytes,
               float(ch.value[-1, 0, 0]))
        if os.environ.get("TNUTS_BENCH_VERBOSE"):
            sys.stderr.write("e2e step %d: wall %.3f host_timing %s\n" % (k, ch.info["wall_seconds"], ch.info["host_timing"]))
        del ch  # the chain's pinned buffer returns to the pool for the next step
        return res

    for k in range(n_warm):
        e2e_step(k)
    barrier()
    t0 = time.perf_counter()
    g_e2e = 0
    l_e2e = 0
    for k in range(args.steps):
        ge, le, d2h, _ = e2e_step(100 + k)
        g_e2e += ge
        l_e2e += le
    barrier()
    wall_e2e = allmax(time.perf_counter() - t0)
    e2e_value = allsum(g_e2e) / wall_e2e
    clk = clocks.stop() if rank == 0 else None

    if rank == 0:
        hbm, _, which = measured_peaks()
        alg = ALG_BYTES_PER_GRAD.get(args.workload, 0)
        achieved = alg * (grads / args.steps) / (dev_run / args.steps) / 1e9
        line = {
            "metric": "grad_evals_per_sec", "value": value, "unit": "grad_evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": n_warm,
            "ms_per_step": dev_total / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, C, N),
            "ess_bulk_min_per_sec": ess_per_sec, "rhat_max": rhat_max,
            "divergent_per_gpu": int(div),
            "grad_evals_per_step": grads_all / args.steps,
            "wall_ms_per_step": wall_value / args.steps * 1e3,
            "e2e": {"value": e2e_value, "unit": "grad_evals/s",
                    "h2d_bytes_per_step": int(model.data_nbytes() + 256),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": wall_e2e / args.steps * 1e3,
                    "ess_bulk_min_per_sec": ess_min_all / (wall_e2e / args.steps),
                    "api": "turing_jl_b200.sample(model, NUTS(0.8), MCMCThreads(), N, n_chains)"},
            "gpu_launches": int(launches + l_e2e),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                         "frac": achieved / hbm, "traffic": None, "peak_source": which,
                         "kernel": "k_chain_run_dyn<EightSchools<8>> (k_chain_run when the grid fits one wave)",
                         "algorithmic_bytes_per_grad_eval": alg,
                         "note": "SURVEY 8(d) streaming figure 6*D*8 B per leapfrog; the kernel keeps "
                                 "the state in registers, so it is fp64-issue bound, see DESIGN.md"},
            "clocks": clk,
        }
        if world > 1:
            line["e2e"]["host_numa_binding"] = "rank pinned to its GPU's NUMA node" if numa is not None else "none"
        if args.workload == "logistic":
            Dm, Nr = model.D, model.N
            if args.tc:
                _, _, _ = measured_peaks()
                pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops", 1614.0) \
                    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1590.0
                dpad = 16 * ((Dm + 15) // 16)
                ach = 18.0 * Nr * dpad * C / (grad_ms * 1e-3) / 1e12
                line["dtype"] = "bf16x3 planes, fp32 accumulate, fp64 log-likelihood"
                line["roofline"] = {
                    "bound": "tensor", "achieved": ach, "peak": pk, "unit": "TFLOP/s", "frac": ach / pk,
                    "traffic": None, "kernel": "k_logistic_tc<128> (+ reduce/finish)",
                    "peak_source": "MEASURED_PEAKS.json bf16 burst" if which == "measured" else "fallback",
                    "algorithmic_flops_per_grad_eval": 18.0 * Nr * dpad,
                    "fp64_equivalent_tflops": 4.0 * Nr * Dm * C / (grad_ms * 1e-3) / 1e12,
                    "note": "achieved = executed bf16 tensor flops (6 plane products for eta, 3 for X^T R) "
                            "per launch / CUDA-event time of gradient launches for all chains"}
            else:
                ach = 4.0 * Nr * Dm * C / (grad_ms * 1e-3) / 1e12
                line["roofline"] = {
                    "bound": "tensor", "achieved": ach, "peak": 37.19, "unit": "TFLOP/s", "frac": ach / 37.19,
                    "traffic": None, "kernel": "k_logistic_mma (+ reduce/finish)",
                    "peak_source": "measured here: mma.sync f64, scripts/fp64_peak.cu (tcgen05 has no f64 kind)",
                    "algorithmic_flops_per_grad_eval": 4.0 * Nr * Dm}
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_baseline(model, spl, N)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


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
