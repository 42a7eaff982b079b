import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import petsc_openacc_b200 as pk
pk.init(0)
ai, aj, aa = pk.gen_powerlaw(10_000_000)
A = pk.Csr(ai, aj, aa)
m = A.m
x = torch.from_numpy(pk.gen_vector(m)).cuda()
y = torch.empty(m, dtype=torch.float64, device="cuda")
for _ in range(3):
    A.mult(x, y, pk.MODE_EXACT)
    A.mult(x, y, pk.MODE_FAST)
torch.cuda.synchronize()
print("done")
