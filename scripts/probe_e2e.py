"""Host-vector MatMult (x up, y down every call) on one GPU: the row-block size of the pipeline, and the
ceiling of this box's PCIe path (pinned copies one way, both ways at once).
Usage: python scripts/probe_e2e.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import petsc_openacc_b200 as pk

pk.init(0)
g = pk.gen_poisson7(300)
m = len(g["ai"]) - 1
hx, hy = pk.PinnedArray(m), pk.PinnedArray(m)
hx.array[:] = pk.gen_vector(m, 1)
# ceiling: pinned copies of the same 216 MB
tx, ty = torch.from_numpy(hx.array), torch.from_numpy(hy.array)
dx = torch.empty(m, dtype=torch.float64, device="cuda")
dy = torch.zeros(m, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def both():
    with torch.cuda.stream(s1):
        dx.copy_(tx, non_blocking=True)
    with torch.cuda.stream(s2):
        ty.copy_(dy, non_blocking=True)


up = timed(lambda: dx.copy_(tx, non_blocking=True))
dn = timed(lambda: ty.copy_(dy, non_blocking=True))
bo = timed(both)
print(f"pinned 216 MB: up {up:.2f} ms ({m*8/up/1e6:.1f} GB/s), down {dn:.2f} ms ({m*8/dn/1e6:.1f} GB/s), both at once {bo:.2f} ms ({m*8/bo/1e6:.1f} GB/s each way)", flush=True)
for rows in (1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 23):
    os.environ["B200_HOST_BLOCK_ROWS"] = str(rows)
    A = pk.Csr(g["ai"], g["aj"], g["aa"])
    ms = timed(lambda: A.mult_host(hx.array, hy.array, pk.MODE_EXACT), 15)
    print(f"B200_HOST_BLOCK_ROWS={rows:8d}: {ms:.3f} ms per MatMult ({m*8/ms/1e6:.1f} GB/s each way)", flush=True)
    A.destroy()
