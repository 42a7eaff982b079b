"""A handful of launches of one kernel on one workload, for ncu.
Usage: python scripts/profile_case.py {poisson300|stencil27|powerlaw} [launches]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gen
import petsc_openacc_b200 as pk

case = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
pk.init(0)
if case == "poisson300":
    g = pk.gen_poisson7(300)
    ai, aj, aa = g["ai"], g["aj"], g["aa"]
    mode = pk.MODE_EXACT
elif case == "stencil27":
    ai, aj, aa = gen.stencil27(200)
    mode = pk.MODE_EXACT
else:
    ai, aj, aa = gen.powerlaw(4_000_000)
    mode = pk.MODE_FAST
m = len(ai) - 1
A = pk.Csr(ai, aj, aa)
x = torch.from_numpy(pk.gen_vector(m)).cuda()
y = torch.empty(m, dtype=torch.float64, device="cuda")
for _ in range(n):
    A.mult(x, y, mode)
torch.cuda.synchronize()
info = A.info()
print(case, "m", m, "nz", len(aj), "kernel", pk.KERNEL_NAMES[info.kernel_fast if mode == pk.MODE_FAST else info.kernel_exact],
      "algorithmic bytes", len(aj) * 12 + m * 20)
