"""Times one Richardson(1)+Jacobi sweep on the 300^3 matrix: fused epilogue vs the four separate passes."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import petsc_openacc_b200 as pk

pk.init(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
g = pk.gen_poisson7(N)
A = pk.Csr(g["ai"], g["aj"], g["aa"])
m = A.m
x = torch.from_numpy(pk.gen_vector(m, 1)).cuda()
b = torch.from_numpy(pk.gen_vector(m, 2)).cuda()
dinv = torch.from_numpy(1.0 / (1.5 + pk.gen_vector(m, 3))).cuda()
out, r = torch.empty_like(x), torch.empty_like(x)


def timeit(fn, n=50, w=5):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n


def unfused():
    A.mult(x, r, pk.MODE_EXACT)          # r = A x
    pk.vec_aypx(r, -1.0, b)              # r = b - r
    pk.vec_pointwise_mult(r, dinv, r)    # z = dinv .* r
    pk.vec_copy(out, x)
    pk.vec_axpy(out, 1.0, r)             # xnew = x + z


t_mult = timeit(lambda: A.mult(x, r, pk.MODE_EXACT))
t_res = timeit(lambda: A.residual(x, b, r, pk.MODE_EXACT))
t_fused = timeit(lambda: A.jacobi_sweep(x, b, dinv, out, pk.MODE_EXACT))
t_unf = timeit(unfused)
ref = out.clone(); unfused()
print(f"{N}^3: MatMult {t_mult:.4f} ms | fused residual {t_res:.4f} ms | fused Jacobi sweep {t_fused:.4f} ms | "
      f"4 separate passes {t_unf:.4f} ms | same bits: {bool(torch.equal(ref, out))}")
