"""torchrun worker: KSPCG + PCJACOBI on the row-partitioned reference problem, one rank per GPU.
Prints one JSON line on rank 0.  Usage: torchrun ... tests/mpiaij_cg_worker.py N [rtol]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import petsc_openacc_b200 as pk


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rtol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-10
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pk.init(local)
    g = pk.gen_poisson7(N, world, rank, vectors=True)
    M = pk.MpiAij(world, rank, g["base"], g["ai"], g["aj"], g["aa"])
    garrays = [None] * world
    dist.all_gather_object(garrays, M.garray())
    for q in range(world):
        M.set_peer_garray(q, garrays[q])
    M.upload()
    handles = [None] * world
    dist.all_gather_object(handles, M.ipc_handle())
    for q in range(world):
        if q != rank and len(M.send_list(q)[0]):
            M.open_peer_window(q, handles[q])
    for q in range(world):
        M.set_rank_window(q, handle=handles[q] if q != rank else None)
    # all-reduce self-test: sum of (rank + 1) and of a rank-dependent fraction, identical bits everywhere
    v = torch.tensor([rank + 1.0, 0.1 * (rank + 1), -3.0], dtype=torch.float64, device=dev)
    M.allreduce_sum(v)
    torch.cuda.synchronize()
    expect = np.array([sum(r + 1.0 for r in range(world)), 0.0, -3.0 * world])
    s = 0.0
    for r in range(world):
        s += 0.1 * (r + 1)
    expect[1] = s
    ok_red = bool(np.array_equal(v.cpu().numpy(), expect))
    b = torch.from_numpy(g["rhs"]).to(dev)
    x = torch.zeros(M.nloc, dtype=torch.float64, device=dev)
    dist.barrier()
    res = M.cg_jacobi(b, x, rtol=rtol, atol=1e-50, max_it=20000, mode=pk.MODE_EXACT)
    err = torch.tensor([float((x - torch.from_numpy(g["exact"]).to(dev)).abs().max())], device=dev, dtype=torch.float64)
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    its = torch.tensor([res.its], device=dev)
    its_all = [torch.zeros_like(its) for _ in range(world)]
    dist.all_gather(its_all, its)
    if rank == 0:
        print(json.dumps({"N": N, "ranks": world, "its": res.its, "its_all": [int(t.item()) for t in its_all],
                          "reason": res.reason, "rnorm": res.rnorm, "rnorm0": res.rnorm0, "linf_err": float(err.item()),
                          "solve_ms": res.solve_ms, "launches": int(res.launches), "allreduce_ok": ok_red}), flush=True)
    dist.barrier()
    M.destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
