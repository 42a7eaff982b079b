"""Input formats: PETSc binary Mat/Vec files and MatrixMarket files load into the same CSR that
scipy builds; malformed files are errors, not crashes."""
import ctypes as C
import os
import struct

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

import gen
import hostlib


def write_petsc_mat(path, ai, aj, aa, n):
    m = len(ai) - 1
    with open(path, "wb") as f:
        f.write(struct.pack(">4i", 1211216, m, n, len(aj)))
        f.write(np.diff(ai).astype(">i4").tobytes())
        f.write(np.asarray(aj).astype(">i4").tobytes())
        f.write(np.asarray(aa).astype(">f8").tobytes())


def write_petsc_vec(path, v):
    with open(path, "wb") as f:
        f.write(struct.pack(">2i", 1211214, len(v)))
        f.write(np.asarray(v).astype(">f8").tobytes())


def csr_of(A):
    L = hostlib.lib()
    m, n, nz = C.c_int(0), C.c_int(0), C.c_int(0)
    pi, pj, pa = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
    hostlib.chk(L.MatSeqAIJGetCSRB200(A, C.byref(m), C.byref(n), C.byref(nz), C.byref(pi), C.byref(pj), C.byref(pa)))
    k = max(nz.value, 1)
    return (m.value, n.value, np.ctypeslib.as_array(pi, shape=(m.value + 1,)).copy(),
            np.ctypeslib.as_array(pj, shape=(k,))[:nz.value].copy(), np.ctypeslib.as_array(pa, shape=(k,))[:nz.value].copy())


def test_petsc_binary_roundtrip(tmp_path):
    rng = np.random.default_rng(2)
    ai, aj, aa = gen.random_csr(300, 170, 12, rng, empty_frac=0.3)
    p = str(tmp_path / "A.petsc")
    write_petsc_mat(p, ai, aj, aa, 170)
    L = hostlib.lib()
    v, A = C.c_void_p(0), C.c_void_p(0)
    hostlib.chk(L.PetscViewerBinaryOpen(2, p.encode(), 0, C.byref(v)))
    hostlib.chk(L.MatLoad(C.byref(A), v))
    hostlib.chk(L.PetscViewerDestroy(C.byref(v)))
    m, n, bi, bj, ba = csr_of(A)
    assert (m, n) == (300, 170) and np.array_equal(bi, ai) and np.array_equal(bj, aj) and np.array_equal(ba, aa)
    hostlib.chk(L.MatDestroy(C.byref(A)))
    x = rng.uniform(-1, 1, 1234)
    pv = str(tmp_path / "x.petsc")
    write_petsc_vec(pv, x)
    vec = C.c_void_p(0)
    hostlib.chk(L.PetscViewerBinaryOpen(2, pv.encode(), 0, C.byref(v)))
    hostlib.chk(L.VecLoad(C.byref(vec), v))
    hostlib.chk(L.PetscViewerDestroy(C.byref(v)))
    assert np.array_equal(hostlib.vec_array(vec, 1234), x)
    hostlib.vec_destroy(vec)


@pytest.mark.parametrize("kind", ["general", "symmetric", "pattern", "duplicates"])
def test_matrix_market(tmp_path, kind):
    rng = np.random.default_rng(5)
    p = str(tmp_path / f"{kind}.mtx")
    if kind == "general":
        S = sp.random(60, 45, density=0.1, random_state=3, format="coo")
        scipy.io.mmwrite(p, S)
        ref = S.tocsr()
    elif kind == "symmetric":
        B = sp.random(50, 50, density=0.08, random_state=4, format="csr")
        S = (B + B.T).tocoo()
        scipy.io.mmwrite(p, S, symmetry="symmetric")
        ref = S.tocsr()
    elif kind == "pattern":
        S = sp.random(30, 70, density=0.1, random_state=5, format="coo")
        scipy.io.mmwrite(p, S, field="pattern")
        ref = sp.csr_matrix((np.ones(S.nnz), (S.row, S.col)), shape=S.shape)
    else:
        with open(p, "w") as f:
            f.write("%%MatrixMarket matrix coordinate real general\n% comment\n3 3 5\n1 1 1.5\n1 1 2.0\n3 2 -1\n2 3 4\n3 2 0.25\n")
        ref = sp.csr_matrix((np.array([3.5, 4.0, -0.75]), (np.array([0, 1, 2]), np.array([0, 2, 1]))), shape=(3, 3))
    ref.sort_indices()
    L = hostlib.lib()
    A = C.c_void_p(0)
    hostlib.chk(L.MatLoadMatrixMarketB200(p.encode(), C.byref(A)))
    m, n, bi, bj, ba = csr_of(A)
    assert (m, n) == ref.shape
    assert np.array_equal(bi, ref.indptr) and np.array_equal(bj, ref.indices)
    np.testing.assert_allclose(ba, ref.data, rtol=0, atol=1e-15)
    hostlib.chk(L.MatDestroy(C.byref(A)))


def test_malformed_files_are_errors(tmp_path):
    L = hostlib.lib()
    A, v = C.c_void_p(0), C.c_void_p(0)
    assert L.PetscViewerBinaryOpen(2, str(tmp_path / "missing").encode(), 0, C.byref(v)) == 65
    bad = tmp_path / "bad.petsc"
    bad.write_bytes(struct.pack(">4i", 1211216, 5, 5, 100))   # truncated
    hostlib.chk(L.PetscViewerBinaryOpen(2, str(bad).encode(), 0, C.byref(v)))
    assert L.MatLoad(C.byref(A), v) == 79
    hostlib.chk(L.PetscViewerDestroy(C.byref(v)))
    notmat = tmp_path / "vec.petsc"
    write_petsc_vec(str(notmat), np.ones(3))
    hostlib.chk(L.PetscViewerBinaryOpen(2, str(notmat).encode(), 0, C.byref(v)))
    assert L.MatLoad(C.byref(A), v) == 79
    hostlib.chk(L.PetscViewerDestroy(C.byref(v)))
    mm = tmp_path / "bad.mtx"
    mm.write_text("%%MatrixMarket matrix coordinate real general\n2 2 3\n1 1 1.0\n5 1 2.0\n")
    assert L.MatLoadMatrixMarketB200(str(mm).encode(), C.byref(A)) == 79
    mm.write_text("not a matrix market file\n")
    assert L.MatLoadMatrixMarketB200(str(mm).encode(), C.byref(A)) == 79
