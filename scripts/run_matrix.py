"""Runs MatMult / MatMultTranspose on an external matrix (.mtx MatrixMarket or PETSc binary) through
the PETSc-named host layer and the C ABI; prints the kernel plan and the achieved rate.
Usage: python scripts/run_matrix.py path/to/matrix.{mtx,petsc,bin}"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import hostlib
import petsc_openacc_b200 as pk


def main():
    path = sys.argv[1]
    L = hostlib.lib()
    A = C.c_void_p(0)
    if path.endswith(".mtx"):
        hostlib.chk(L.MatLoadMatrixMarketB200(path.encode(), C.byref(A)))
    else:
        v = C.c_void_p(0)
        hostlib.chk(L.PetscViewerBinaryOpen(2, path.encode(), 0, C.byref(v)))
        hostlib.chk(L.MatLoad(C.byref(A), v))
        hostlib.chk(L.PetscViewerDestroy(C.byref(v)))
    m, n, nz = C.c_int(0), C.c_int(0), C.c_int(0)
    pi, pj, pa = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
    hostlib.chk(L.MatSeqAIJGetCSRB200(A, C.byref(m), C.byref(n), C.byref(nz), C.byref(pi), C.byref(pj), C.byref(pa)))
    ai = np.ctypeslib.as_array(pi, shape=(m.value + 1,))
    aj = np.ctypeslib.as_array(pj, shape=(max(nz.value, 1),))[:nz.value]
    aa = np.ctypeslib.as_array(pa, shape=(max(nz.value, 1),))[:nz.value]
    pk.init(0)
    H = pk.Csr(ai, aj, aa, n=n.value)
    info = H.info()
    print(f"{os.path.basename(path)}: {m.value} x {n.value}, nnz {nz.value}, longest row {info.rmax}, "
          f"plan fast={pk.KERNEL_NAMES[info.kernel_fast]} exact={pk.KERNEL_NAMES[info.kernel_exact]} "
          f"diagonal codes={info.index8_diagonals}")
    x = torch.from_numpy(pk.gen_vector(n.value)).cuda()
    y = torch.empty(m.value, dtype=torch.float64, device="cuda")
    for mode, name in ((pk.MODE_EXACT, "exact"), (pk.MODE_FAST, "fast")):
        for _ in range(5):
            H.mult(x, y, mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            H.mult(x, y, mode)
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / 50
        print(f"  MatMult {name:5s}: {ms * 1e3:9.2f} us  {(nz.value * 12 + m.value * 20) / ms / 1e6:8.1f} GB/s algorithmic  {2 * nz.value / ms / 1e6:8.1f} GFLOP/s")
    H.destroy()
    hostlib.chk(L.MatDestroy(C.byref(A)))


if __name__ == "__main__":
    main()
