"""Single-GPU probe of the fused MatMult_MPIAIJ kernel: all ranks of a 300^3 decomposition live in
this process; per iteration every rank pushes first (stand-alone push kernels), then one rank's
A-only kernel, fused A+B kernel and A,B two-kernel sequence are timed in isolation."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import petsc_openacc_b200 as pk

pk.init(0)
size = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
ranks = []
for r in range(size):
    g = pk.gen_poisson7(N, size, r)
    ranks.append(pk.MpiAij(size, r, g["base"], g["ai"], g["aj"], g["aa"]))
base = g["base"]
garrays = [M.garray() for M in ranks]
for M in ranks:
    for q in range(size):
        M.set_peer_garray(q, garrays[q])
    M.upload()
for M in ranks:
    for q in range(size):
        if q != M.rank:
            M.set_peer_window(q, ranks[q].window_ptr())
xg = pk.gen_vector(N ** 3, 0xB200)
xs = [torch.from_numpy(xg[base[r]:base[r + 1]].copy()).cuda() for r in range(size)]
ys = [torch.zeros(ranks[r].nloc, dtype=torch.float64, device="cuda") for r in range(size)]


def timed(which, fn, n=100):
    ts = []
    for it in range(n + 10):
        for r, M in enumerate(ranks):
            M.mult_begin(xs[r])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        if it >= 10:
            ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"  {which:34s} median {np.median(ts):8.2f} us  min {np.min(ts):8.2f} us", flush=True)


for r in sorted({0, size - 1}):
    M = ranks[r]
    print(f"size={size} rank={r}: rows={M.nloc} ghosts={M.nghost} B rows={M.brows}")
    timed("A only (mult_local)", lambda: M.mult_local(xs[r], ys[r], pk.MODE_EXACT))
    timed("A then B, two kernels", lambda: (M.mult_local(xs[r], ys[r], pk.MODE_EXACT), M.mult_end(ys[r], pk.MODE_EXACT)))
    timed("fused A+B (mult_finish)", lambda: M.mult_finish(xs[r], ys[r], pk.MODE_EXACT))
    M.check()
