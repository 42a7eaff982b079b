"""A compact pass over every kernel for compute-sanitizer (one tool per run):
stream (1-byte codes and int32), row, vector, merge + fix-up, cprow, transposes, fused epilogues, the
fused MatMult_MPIAIJ kernel with two in-process ranks, vector kernels and the CG driver."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gen
import petsc_openacc_b200 as pk

pk.init(0)
rng = np.random.default_rng(1)


def run_all(ai, aj, aa, n, kernels):
    m = len(ai) - 1
    A = pk.Csr(ai, aj, aa, n=n)
    x = torch.from_numpy(gen.uniform_pm1(n, 1)).cuda()
    xt = torch.from_numpy(gen.uniform_pm1(m, 2)).cuda()
    y = torch.zeros(m, dtype=torch.float64, device="cuda")
    yt = torch.zeros(n, dtype=torch.float64, device="cuda")
    for k in kernels:
        try:
            A.set_kernel(k)
        except pk.B200Error:
            continue
        for mode in (pk.MODE_EXACT, pk.MODE_FAST):
            if mode == pk.MODE_EXACT and k in (pk.KERNEL_VECTOR, pk.KERNEL_MERGE):
                continue
            A.mult(x, y, mode)
            A.mult_add(x, y, y, mode)
    A.set_kernel(pk.KERNEL_AUTO)
    A.mult_transpose(xt, yt, pk.MODE_EXACT)
    if m == n:
        b = torch.ones(m, dtype=torch.float64, device="cuda")
        A.residual(x, b, y, pk.MODE_EXACT)
        A.jacobi_sweep(x, b, b, y, pk.MODE_EXACT)
    A.mult_host(x.cpu().numpy(), mode=pk.MODE_EXACT)
    torch.cuda.synchronize()
    A.destroy()


g = pk.gen_poisson7(12)
ALLK = (pk.KERNEL_STREAM, pk.KERNEL_ROW, pk.KERNEL_VECTOR, pk.KERNEL_MERGE, pk.KERNEL_CPROW)
run_all(g["ai"], g["aj"], g["aa"], 12 ** 3, ALLK)
os.environ["B200_INDEX8"] = "0"
run_all(g["ai"], g["aj"], g["aa"], 12 ** 3, (pk.KERNEL_STREAM,))
os.environ.pop("B200_INDEX8")
ai, aj, aa = gen.stencil27(8)
run_all(ai, aj, aa, 512, (pk.KERNEL_STREAM,))
ai, aj, aa = gen.powerlaw(3000, lmax=2500)
run_all(ai, aj, aa, 3000, ALLK)
ai, aj, aa = gen.random_csr(2000, 300, 3, rng, empty_frac=0.9)
run_all(ai, aj, aa, 300, ALLK)
# fused MatMult_MPIAIJ, two ranks in this process
ranks = []
for r in range(2):
    gg = pk.gen_poisson7(10, 2, r, vectors=True)
    ranks.append((pk.MpiAij(2, r, gg["base"], gg["ai"], gg["aj"], gg["aa"]), gg))
for M, _ in ranks:
    for q in range(2):
        M.set_peer_garray(q, ranks[q][0].garray())
    M.upload()
for M, _ in ranks:
    M.set_peer_window(1 - M.rank, ranks[1 - M.rank][0].window_ptr())
xs = [torch.from_numpy(pk.gen_vector(M.nloc, 3)).cuda() for M, _ in ranks]
ys = [torch.zeros(M.nloc, dtype=torch.float64, device="cuda") for M, _ in ranks]
for it in range(2):
    for (M, _), x in zip(ranks, xs):
        M.mult_begin(x)
    for (M, _), x, y in zip(ranks, xs, ys):
        M.mult_finish(x, y, pk.MODE_EXACT)
    for (M, _), x in zip(ranks, xs):
        M.mult_begin(x)
    for (M, _), x, y in zip(ranks, xs, ys):
        M.mult_local(x, y, pk.MODE_EXACT)
        M.mult_end(y, pk.MODE_EXACT)
torch.cuda.synchronize()
for M, _ in ranks:
    M.check()
    M.destroy()
# CG
A = pk.Csr(g["ai"], g["aj"], g["aa"])
gv = pk.gen_poisson7(12, vectors=True)
res = A.cg_jacobi(torch.from_numpy(gv["rhs"]).cuda(), torch.zeros(A.m, dtype=torch.float64, device="cuda"), rtol=1e-8, atol=1e-50, max_it=500, mode=pk.MODE_EXACT)
A.destroy()
print("sanitize_case done: cg its", res.its, "launches so far", pk.launch_count())
