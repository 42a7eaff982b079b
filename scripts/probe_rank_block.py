"""Times the diagonal-block kernel of one rank of an 8-rank 300^3 decomposition on one GPU."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import petsc_openacc_b200 as pk

pk.init(0)
size = int(sys.argv[1]) if len(sys.argv) > 1 else 8
g = pk.gen_poisson7(300, size, 0)
M = pk.MpiAij(size, 0, g["base"], g["ai"], g["aj"], g["aa"])
M.upload()
x = torch.from_numpy(pk.gen_vector(M.nloc)).cuda()
y = torch.zeros(M.nloc, dtype=torch.float64, device="cuda")
Ai, Aj, Aa = M.block(0)
nbytes = len(Aj) * 12 + M.nloc * 20


def timeit(fn, n=300, w=30):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for env in ({}, {"B200_STREAM_THREADS": "128"}, {"B200_STREAM_STAGES": "1"}, {"B200_STREAM_THREADS": "128", "B200_STREAM_STAGES": "1"}):
    os.environ.update(env)
    A = pk.Csr(Ai, Aj, Aa)
    info = A.info()
    t = timeit(lambda: A.mult(x, y, pk.MODE_EXACT))
    print(f"size={size} rows={M.nloc} env={env} tiles={info.stream_tiles}: {t:.2f} us  {nbytes/t/1e3:.0f} GB/s (ideal at 6995 GB/s: {nbytes/6995e3:.2f} us)")
    A.set_kernel(pk.KERNEL_ROW)
    t = timeit(lambda: A.mult(x, y, pk.MODE_EXACT))
    print(f"   k_row: {t:.2f} us {nbytes/t/1e3:.0f} GB/s")
    A.destroy()
    for k in env:
        os.environ.pop(k)
