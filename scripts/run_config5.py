"""BASELINE.json configs[4]: logistic regression with the data ROWS sharded over the GPUs of one
box and a per-leapfrog NCCL all-reduce of the [D+1][chains] likelihood/gradient block.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node=G --master-addr 127.0.0.1 \
      --master-port 29613 scripts/run_config5.py [rows_per_gpu] [D] [chains] [warm-up] [draws] [--tc]
Defaults are the config's per-GPU shard (6.25 M rows x 256, 256 chains) and a bounded schedule
(12 warm-up transitions, no draws); with draws > 0 the kept draws give R-hat and bulk ESS.
--tc: likelihood on the tcgen05 tensor cores (D > 128: the CTA-pair kernel k_logistic_tc_wide)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import turing_jl_b200 as T
from turing_jl_b200 import distributed as D


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    a = [x for x in sys.argv[1:] if not x.startswith("--")]
    tc = "--tc" in sys.argv
    rows = int(a[0]) if len(a) > 0 else 6_250_000
    Dm = int(a[1]) if len(a) > 1 else 256
    C = int(a[2]) if len(a) > 2 else 256
    n_warm = int(a[3]) if len(a) > 3 else 12
    n_draws = int(a[4]) if len(a) > 4 else 0
    n_tr = n_warm + n_draws
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    # synthetic shard generated on the device (seed 3): X ~ N(0,1), y ~ Bernoulli(sigma(X b*))
    g = torch.Generator(device=dev); g.manual_seed(3)
    beta = torch.randn(Dm, generator=g, device=dev, dtype=torch.float64) * (2.0 / Dm ** 0.5)
    g.manual_seed(1000 + rank)
    X = torch.empty((rows, Dm), device=dev, dtype=torch.float64)
    step = 1 << 20
    for i in range(0, rows, step):
        X[i:i + step].normal_(generator=g)
    y = (torch.rand(rows, generator=g, device=dev, dtype=torch.float64) < torch.sigmoid(X @ beta)).double()
    model = T.models.DeviceLogisticRegression(X, y)
    spl = T.NUTS(n_warm, 0.8)
    eng = T.Engine(model, spl, C, n_warm, 7, device=local, chain_offset=0, row_shards=world, row_rank=rank, strategy=2)
    if tc:
        eng.set_option("logistic_tc", 1)
    eng.set_option("time_kernels", 1)
    if world > 1:
        D.attach_comm(eng, dist, device="cuda:%d" % local)
    torch.cuda.synchronize(); dist.barrier()
    # start near the data-generating coefficients: from U(-2,2) the first warm-up trees of a
    # 50M-row posterior (sd ~ 1e-4) hit max depth (1023 whole-data gradients per transition)
    th0 = (beta.cpu().numpy()[None, :] + 1e-4 * np.random.default_rng(5).standard_normal((C, Dm)))
    t0 = time.perf_counter(); eng.init(th0); torch.cuda.synchronize(); dist.barrier(); t_init = time.perf_counter() - t0
    g_init = eng.grad_evals
    grad_ms = eng.bench_gradient(3) * 1e3
    t0 = time.perf_counter(); eng.run(n_tr, n_warm + 1 if n_draws else 1, 1); torch.cuda.synchronize(); dist.barrier(); t_run = time.perf_counter() - t0
    grads = eng.grad_evals - g_init
    kt = eng.kernel_times()
    rhat = ess = None
    if n_draws >= 8:
        rhat, ess = eng.diagnostics()
    th = torch.from_numpy(eng.theta()).cuda(); ref = th.clone(); dist.broadcast(ref, src=0)
    same = bool(torch.equal(th, ref))
    if rank == 0:
        flops = 4.0 * rows * world * Dm
        print(json.dumps({
            "config": {"workload": "BASELINE.json configs[4]-shaped: logistic regression, rows sharded over %d GPU(s), %d rows x %d per GPU, %d chains, NUTS(0.8), %d warm-up + %d kept transitions timed, likelihood: %s"
                       % (world, rows, Dm, C, n_warm, n_draws, "tcgen05 bf16x3" if tc else "fp64 DMMA")},
            "n_gpus": world, "rows_total": rows * world, "D": Dm, "chains": C,
            "metric": "grad_evals_per_sec", "value": grads / t_run, "unit": "grad_evals/s (whole-data gradients)",
            "run_seconds": t_run, "init_seconds": t_init, "grad_evals": int(grads),
            "tflops_all_gpus": grads / t_run * flops / 1e12,
            "gradient_ms_all_chains_incl_allreduce": grad_ms,
            "gradient_kernel_seconds_in_run": kt[0], "gradient_kernel_launches_in_run": kt[1],
            "gradient_kernel_ms_per_launch_in_run": (kt[0] / kt[1] * 1e3) if kt[1] else None,
            "gradient_kernel_share_of_run": kt[0] / t_run,
            "rhat_max": None if rhat is None else float(np.max(rhat)),
            "ess_bulk_min": None if ess is None else float(np.min(ess)),
            "ess_bulk_min_per_sec": None if ess is None else float(np.min(ess)) / t_run,
            "step_size_mean": float(np.mean(eng.stepsize())),
            "gradient_tflops_per_gpu": 4.0 * rows * Dm * C / (grad_ms * 1e-3) / 1e12,
            "allreduce_bytes_per_leapfrog": (Dm + 1) * eng.C * 8,
            "identical_across_ranks": same, "gpu_launches": int(eng.launches)}))
    eng.close()
    dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
