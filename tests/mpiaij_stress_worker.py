"""torchrun worker: halo stress test of the fused MatMult_MPIAIJ (one rank per GPU).

STEPS back-to-back b200_mpiaij_mult calls, a different x every step, nothing synchronised in between,
every result checked on the device.  x_k = 2^(e_k) * x_0 with e_k cycling through -8..8, so the exact
answer is y_k = 2^(e_k) * y_0 bit for bit (scaling by a power of two is exact) and y_0 is checked
against the oracle once: a ghost value taken from the wrong MatMult (stale buffer, flag seen before
the data, a peer running ahead) carries another scale factor and shows up as a mismatch.
Usage: torchrun ... tests/mpiaij_stress_worker.py N STEPS"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import oracle
import petsc_openacc_b200 as pk


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pk.init(local)
    os.environ.setdefault("B200_HOST_BLOCK_ROWS", "8192")     # several row blocks even on a small grid
    g = pk.gen_poisson7(N, world, rank)
    base = g["base"]
    M = pk.MpiAij(world, rank, base, g["ai"], g["aj"], g["aa"])
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
    xg = pk.gen_vector(int(base[-1]), 0xB200)
    x0 = torch.from_numpy(xg[base[rank]:base[rank + 1]].copy()).to(dev)
    x = torch.empty_like(x0)
    y = torch.empty_like(x0)
    y0 = torch.empty_like(x0)
    want = torch.empty_like(x0)
    bad = torch.zeros(1, dtype=torch.int64, device=dev)
    M.mult(x0, y0, pk.MODE_EXACT)
    torch.cuda.synchronize()
    M.check()
    Ai, Aj, Aa = M.block(0)
    Bi, Bj, Ba = M.block(1)
    ref = oracle.matmult(Ai, Aj, Aa, xg[base[rank]:base[rank + 1]])
    if M.nghost:
        ref = oracle.matmultadd(Bi, Bj, Ba, xg[garrays[rank]], ref)
    first_ok = bool(np.array_equal(ref, y0.cpu().numpy()))
    # the host-vector entry (row-blocked pipeline: uploads, block kernels, downloads and the ghost-row
    # patch overlap): same bits, several times in a row
    hx, hy = pk.PinnedArray(M.nloc), pk.PinnedArray(M.nloc)
    hx.array[:] = xg[base[rank]:base[rank + 1]]
    host_ok = True
    for _ in range(5):
        hy.array[:] = np.nan
        M.mult_host(hx.array, hy.array, pk.MODE_EXACT)
        host_ok = host_ok and bool(np.array_equal(ref, hy.array))
    first_ok = first_ok and host_ok
    dist.barrier()
    # ranks are deliberately skewed: odd ranks do extra local work every few steps
    junk = torch.zeros(1 << 20, dtype=torch.float64, device=dev)
    for k in range(steps):
        s = 2.0 ** ((k % 17) - 8)
        torch.mul(x0, s, out=x)
        if (rank + k) % 7 == 0:
            junk.add_(1.0)
        M.mult(x, y, pk.MODE_EXACT)
        torch.mul(y0, s, out=want)
        bad += (y != want).sum()
    torch.cuda.synchronize()
    M.check()
    tot = bad.clone()
    dist.all_reduce(tot)
    ok0 = torch.tensor([1 if first_ok else 0], device=dev)
    dist.all_reduce(ok0, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"N": N, "ranks": world, "steps": steps, "mismatches": int(tot.item()),
                          "first_result_equals_oracle": bool(int(ok0.item()))}), flush=True)
    dist.barrier()
    M.destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
