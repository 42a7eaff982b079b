"""A handful of launches of one kernel on one workload, for ncu.
Usage: python scripts/profile_case.py {poisson300|stencil27|powerlaw|cg300} [launches]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import petsc_openacc_b200 as pk

case = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
pk.init(0)
mode = pk.MODE_EXACT
if case in ("poisson300", "cg300"):
    g = pk.gen_poisson7(300, vectors=True)
    ai, aj, aa = g["ai"], g["aj"], g["aa"]
elif case == "stencil27":
    ai, aj, aa = pk.gen_stencil27(200)
else:
    ai, aj, aa = pk.gen_powerlaw(10_000_000)
m = len(ai) - 1
A = pk.Csr(ai, aj, aa)
x = torch.from_numpy(pk.gen_vector(m)).cuda()
y = torch.empty(m, dtype=torch.float64, device="cuda")
if case == "cg300":
    b = torch.from_numpy(g["rhs"]).cuda()
    A.cg_jacobi(b, y, rtol=1e-30, atol=1e-300, max_it=n, mode=mode)
else:
    for _ in range(n):
        A.mult(x, y, mode)
torch.cuda.synchronize()
info = A.info()
print(case, "m", m, "nz", len(aj), "kernel", pk.KERNEL_NAMES[info.kernel_exact], "algorithmic bytes", len(aj) * 12 + m * 20)
