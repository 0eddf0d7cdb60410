"""torchrun script: the multi-GPU modes against a single-GPU engine (turing.jl_b200/multi_gpu_checks.py).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node=N --master-addr 127.0.0.1 \
      --master-port 29611 scripts/multi_gpu_check.py --rows [--tc] [--wide] [--big] | --diag | --mesh | --all"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import turing_jl_b200 as T
from turing_jl_b200 import multi_gpu_checks as M


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t0 = time.perf_counter()
    if "--all" in sys.argv:
        res = M.run_all(dist, rank, world, local)
        ok, tag = all(v["ok"] for v in res.values()), "ALL-OK"
    elif "--diag" in sys.argv:
        ok, det = M.diag_check(dist, rank, world, local)
        res, tag = {"pooled_diagnostics": det}, "DIAG-OK"
    elif "--mesh" in sys.argv:
        ok, det = M.mesh_check(dist, rank, world, local, 2)
        res, tag = {"mesh_chains_x_rows": det}, "MESH-OK"
    else:
        ok, det = M.rows_check(dist, rank, world, local, tc="--tc" in sys.argv, big="--big" in sys.argv,
                               wide="--wide" in sys.argv)
        res, tag = {"rows_sharded": det}, "ROWS-OK"
    if rank == 0:
        print(json.dumps(res))
        print("%.1f s" % (time.perf_counter() - t0))
        if ok:
            print(tag)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
