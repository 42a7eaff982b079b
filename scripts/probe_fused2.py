"""Single-GPU proxy of the per-rank kernel of the 8-GPU (or 2-, 4-GPU) MatMult_MPIAIJ: all ranks of
the 300^3 decomposition live in this process; the pushes are issued once, then ONE rank's kernel is
launched back to back (steady state, programmatic dependent launch in effect) and timed between two
events: A only, fused A + ghost rows, and the ideal time of its bytes at the rate the 1-GPU MatMult
reaches.  Usage: python scripts/probe_fused2.py [ranks=8] [N=300] [launches=400]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import petsc_openacc_b200 as pk

pk.init(0)
size = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 400
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
for r, M in enumerate(ranks):
    M.mult_begin(xs[r])
torch.cuda.synchronize()


def timed(which, fn):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"  {which:40s} {us:8.2f} us per launch", flush=True)
    return us


print(f"pdl={os.environ.get('B200_PDL', '1')} ranks={size} N={N}")
for r in sorted({0, size - 1}):
    M = ranks[r]
    nzA = M.annz
    phys = nzA * 9 + M.nloc * 20
    print(f"rank {r}: rows={M.nloc} nnz(A)={nzA} ghosts={M.nghost} B rows={M.brows}; {phys/1e6:.1f} MB streamed -> {phys/7.10e6:.1f} us at 7.10 TB/s")
    timed("A only (mult_local), back to back", lambda: M.mult_local(xs[r], ys[r], pk.MODE_EXACT))
    timed("fused A + ghost rows (mult_finish)", lambda: M.mult_finish(xs[r], ys[r], pk.MODE_EXACT))
    M.check()
