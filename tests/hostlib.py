"""ctypes view of libb200petsc.so (the PETSc-named host layer) for the tests."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "petsc-openacc_b200", "libb200petsc.so")
BIN = os.path.join(ROOT, "petsc-openacc_b200", "bin")

_lib = None


def lib():
    global _lib
    if _lib is None:
        C.CDLL(os.path.join(ROOT, "petsc-openacc_b200", "libb200aij.so"), mode=C.RTLD_GLOBAL)
        _lib = C.CDLL(PATH)
    return _lib


def chk(rc):
    assert rc == 0, f"PetscErrorCode {rc}"


class System:
    """createSystem(N) through the PETSc-shaped API."""

    def __init__(self, N):
        L = lib()
        self.N = N
        self.da, self.A, self.lhs, self.rhs, self.exact = (C.c_void_p(0) for _ in range(5))
        chk(L.b200_create_poisson_system(C.c_int(N), C.byref(self.da), C.byref(self.A), C.byref(self.lhs),
                                         C.byref(self.rhs), C.byref(self.exact)))

    def csr(self):
        m, n, nz = C.c_int(0), C.c_int(0), C.c_int(0)
        pi, pj, pa = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
        chk(lib().MatSeqAIJGetCSRB200(self.A, C.byref(m), C.byref(n), C.byref(nz), C.byref(pi), C.byref(pj), C.byref(pa)))
        ai = np.ctypeslib.as_array(pi, shape=(m.value + 1,)).copy()
        aj = np.ctypeslib.as_array(pj, shape=(max(nz.value, 1),))[:nz.value].copy()
        aa = np.ctypeslib.as_array(pa, shape=(max(nz.value, 1),))[:nz.value].copy()
        return ai, aj, aa

    def info(self):
        a, b, c, d, e = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        chk(lib().MatSeqAIJGetInfoB200(self.A, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e)))
        return dict(nonzerorowcnt=a.value, rmax=b.value, compressedrow=bool(c.value), cprow_nrows=d.value, fshift=e.value)

    def destroy(self):
        chk(lib().b200_destroy_poisson_system(C.byref(self.da), C.byref(self.A), C.byref(self.lhs), C.byref(self.rhs),
                                              C.byref(self.exact)))


def vec_array(v, n):
    p = C.POINTER(C.c_double)()
    chk(lib().VecGetArrayRead(v, C.byref(p)))
    out = np.ctypeslib.as_array(p, shape=(n,)).copy()
    chk(lib().VecRestoreArrayRead(v, C.byref(p)))
    return out


def vec_from(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    v = C.c_void_p(0)
    chk(lib().VecCreateSeq(C.c_int(2), C.c_int(len(a)), C.byref(v)))
    p = C.POINTER(C.c_double)()
    chk(lib().VecGetArray(v, C.byref(p)))
    np.ctypeslib.as_array(p, shape=(max(len(a), 1),))[:len(a)] = a
    chk(lib().VecRestoreArray(v, C.byref(p)))
    return v


def vec_destroy(v):
    chk(lib().VecDestroy(C.byref(v)))


def mat_from_csr(ai, aj, aa, n, slack=0, rng=None):
    """MatCreateSeqAIJ + MatSetValues row by row (optionally over-preallocated / shuffled)."""
    L = lib()
    m = len(ai) - 1
    nnz = np.maximum(np.diff(ai) + slack, 0).astype(np.int32)
    A = C.c_void_p(0)
    chk(L.MatCreateSeqAIJ(C.c_int(2), C.c_int(m), C.c_int(n), C.c_int(0), nnz.ctypes.data_as(C.c_void_p), C.byref(A)))
    for i in range(m):
        cols = np.ascontiguousarray(aj[ai[i]:ai[i + 1]], dtype=np.int32)
        vals = np.ascontiguousarray(aa[ai[i]:ai[i + 1]], dtype=np.float64)
        if rng is not None and len(cols) > 1:
            p = rng.permutation(len(cols))
            cols, vals = np.ascontiguousarray(cols[p]), np.ascontiguousarray(vals[p])
        row = np.array([i], np.int32)
        chk(L.MatSetValues(A, C.c_int(1), row.ctypes.data_as(C.c_void_p), C.c_int(len(cols)),
                           cols.ctypes.data_as(C.c_void_p), vals.ctypes.data_as(C.c_void_p), C.c_int(1)))
    chk(L.MatAssemblyBegin(A, C.c_int(0)))
    chk(L.MatAssemblyEnd(A, C.c_int(0)))
    return A
