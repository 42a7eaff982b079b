"""torchrun worker: CG + row-partitioned multigrid (petsc-openacc_b200/dgamg.py) on the reference
problem, one rank per GPU.  Prints one JSON line on rank 0.
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/dgamg_worker.py 300
Round 2: 200^3 on 2 GPUs and 300^3 on 8 GPUs (profiles/r02_dgamg_*.log)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import petsc_openacc_b200 as pk
from petsc_openacc_b200 import dgamg


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pk.init(local)
    comm = dgamg.TorchComm()
    g = pk.gen_poisson7(N, world, rank, vectors=True)
    t0 = time.perf_counter()
    levels = dgamg.setup(comm, g["base"], g["ai"], g["aj"], g["aa"])
    t_setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    sv = dgamg.Solver(comm, levels)
    t_upload = time.perf_counter() - t0
    b = torch.from_numpy(g["rhs"]).to(dev)
    x = torch.zeros_like(b)
    times = []
    for _ in range(2):                      # the second solve is the warm one
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        its, reason, rnorm = sv.solve(b, x)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    err = (x - torch.from_numpy(g["exact"]).to(dev)).abs().max().reshape(1)
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"N": N, "ranks": world, "levels": [int(L.base[-1]) for L in levels], "its": its, "reason": reason,
                          "rnorm": rnorm, "linf_err": float(err.item()), "setup_s": t_setup, "upload_s": t_upload,
                          "solve_s_first": times[0], "solve_s_warm": times[1]}), flush=True)
    dist.barrier()
    sv.destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
