"""Small matrices (GAMG coarse levels): per-launch latency of the stream kernel vs thread-per-row."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import petsc_openacc_b200 as pk

pk.init(0)


def timeit(fn, n=400, w=50):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for N in (6, 10, 16, 24, 32, 48, 64, 96, 128):
    g = pk.gen_poisson7(N)
    A = pk.Csr(g["ai"], g["aj"], g["aa"])
    m = A.m
    x = torch.from_numpy(pk.gen_vector(m)).cuda()
    y = torch.empty(m, dtype=torch.float64, device="cuda")
    out = []
    for k in (pk.KERNEL_STREAM, pk.KERNEL_ROW):
        A.set_kernel(k)
        out.append(timeit(lambda: A.mult(x, y, pk.MODE_EXACT)))
    A.set_kernel(pk.KERNEL_AUTO)
    auto = timeit(lambda: A.mult(x, y, pk.MODE_EXACT))
    print(f"N={N:4d} rows={m:8d} tiles={A.info().stream_tiles:6d}  stream {out[0]:7.2f} us   row {out[1]:7.2f} us   auto {auto:7.2f} us", flush=True)
    A.destroy()
