"""A few fused A+B launches of rank 0 of an 8-rank 300^3 decomposition on one GPU, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import petsc_openacc_b200 as pk

pk.init(0)
size, N = 8, 300
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
y = torch.zeros(ranks[0].nloc, dtype=torch.float64, device="cuda")
for it in range(4):
    for r, M in enumerate(ranks):
        M.mult_begin(xs[r])
    ranks[0].mult_finish(xs[0], y, pk.MODE_EXACT)
torch.cuda.synchronize()
ranks[0].check()
print("fused rank-0 launches done; rows", ranks[0].nloc, "ghost rows", ranks[0].brows)
