"""How much of x stays in L2 under the power-law MatMult?  Same 10 M rows / row lengths, column count
(= length of x) swept: 2.5 M (20 MB) ... 10 M (80 MB).  Usage: python scripts/probe_powerlaw_l2.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import petsc_openacc_b200 as pk

pk.init(0)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
for n in (m // 4, m // 2, 3 * m // 4, m):
    ai, aj, aa = pk.gen_powerlaw(m, n)
    A = pk.Csr(ai, aj, aa, n=n)
    x = torch.from_numpy(pk.gen_vector(n, 1)).cuda()
    y = torch.zeros(m, dtype=torch.float64, device="cuda")
    for _ in range(5):
        A.mult(x, y, pk.MODE_EXACT)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        A.mult(x, y, pk.MODE_EXACT)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 20
    nz = len(aj)
    print(f"n={n:9d} (x = {n*8/1e6:5.1f} MB) nnz={nz} ms={ms:.4f}  {nz/ms/1e6:.1f} Gnnz/s  alg {(nz*12+m*20)/ms/1e6:.0f} GB/s", flush=True)
    A.destroy(); del x, y
