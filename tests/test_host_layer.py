"""CPU tests of the PETSc-shaped host layer: the MatSetValues -> MatAssemblyEnd_SeqAIJ route of
createSystem gives the oracle's CSR bit for bit; error conventions; symbols exported."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import gen
import hostlib
import oracle


@pytest.mark.parametrize("N", [3, 8, 20])
def test_create_system_matches_oracle_bit_exact(N):
    s = hostlib.System(N)
    ai, aj, aa = s.csr()
    p = oracle.poisson7(N)
    assert np.array_equal(ai, p["ai"]) and np.array_equal(aj, p["aj"]) and np.array_equal(aa, p["aa"])
    assert np.array_equal(hostlib.vec_array(s.rhs, N ** 3), p["rhs"])
    assert np.array_equal(hostlib.vec_array(s.exact, N ** 3), p["exact"])
    assert np.array_equal(hostlib.vec_array(s.lhs, N ** 3), np.zeros(N ** 3))
    info = s.info()
    assert info["nonzerorowcnt"] == N ** 3 and info["rmax"] == 7 and not info["compressedrow"]
    s.destroy()


def test_assembly_with_slack_shuffled_inserts_and_compressed_rows():
    """MatAssemblyEnd_SeqAIJ compaction (unused preallocated slots) and MatCheckCompressedRow."""
    rng = np.random.default_rng(4)
    ai, aj, aa = gen.random_csr(400, 300, 9, rng, empty_frac=0.7)
    A = hostlib.mat_from_csr(ai, aj, aa, 300, slack=3, rng=rng)
    m, n, nz = C.c_int(0), C.c_int(0), C.c_int(0)
    pi, pj, pa = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
    L = hostlib.lib()
    hostlib.chk(L.MatSeqAIJGetCSRB200(A, C.byref(m), C.byref(n), C.byref(nz), C.byref(pi), C.byref(pj), C.byref(pa)))
    assert (m.value, n.value, nz.value) == (400, 300, len(aj))
    assert np.array_equal(np.ctypeslib.as_array(pi, shape=(401,)), ai)
    assert np.array_equal(np.ctypeslib.as_array(pj, shape=(nz.value,)), aj)
    assert np.array_equal(np.ctypeslib.as_array(pa, shape=(nz.value,)), aa)
    a, b, c, d, e = (C.c_int(0) for _ in range(5))
    hostlib.chk(L.MatSeqAIJGetInfoB200(A, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e)))
    nzr = int((np.diff(ai) > 0).sum())
    use, cpi, ridx = oracle.check_compressed_row(ai, nzr)
    assert a.value == nzr and b.value == int(np.diff(ai).max()) and bool(c.value) == use
    assert d.value == (len(ridx) if use else 0)
    assert e.value == 3 * 400  # fshift = the unused slots
    hostlib.chk(L.MatDestroy(C.byref(A)))


def test_under_preallocation_reallocates_like_petsc():
    rng = np.random.default_rng(6)
    ai, aj, aa = gen.random_csr(60, 80, 30, rng)
    A = hostlib.mat_from_csr(ai, aj, aa, 80, slack=-1000)  # clamps to 0 reserved slots
    pi, pj, pa = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
    nz = C.c_int(0)
    L = hostlib.lib()
    hostlib.chk(L.MatSeqAIJGetCSRB200(A, None, None, C.byref(nz), C.byref(pi), C.byref(pj), C.byref(pa)))
    assert nz.value == len(aj)
    assert np.array_equal(np.ctypeslib.as_array(pj, shape=(nz.value,)), aj)
    assert np.array_equal(np.ctypeslib.as_array(pa, shape=(nz.value,)), aa)
    hostlib.chk(L.MatDestroy(C.byref(A)))


def test_error_convention_nonzero_code_no_abort():
    L = hostlib.lib()
    A = C.c_void_p(0)
    nnz = np.array([1, 1], np.int32)
    hostlib.chk(L.MatCreateSeqAIJ(2, 2, 2, 0, nnz.ctypes.data_as(C.c_void_p), C.byref(A)))
    x, y = hostlib.vec_from(np.ones(2)), hostlib.vec_from(np.ones(2))
    assert L.MatMult(A, x, y) == 73          # unassembled: PETSC_ERR_ARG_WRONGSTATE
    row, col, v = np.array([5], np.int32), np.array([0], np.int32), np.array([1.0])
    rc = L.MatSetValues(A, 1, row.ctypes.data_as(C.c_void_p), 1, col.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), 1)
    assert rc == 63                           # PETSC_ERR_ARG_OUTOFRANGE
    hostlib.chk(L.MatAssemblyBegin(A, 0)); hostlib.chk(L.MatAssemblyEnd(A, 0))
    assert L.MatMult(A, x, x) == 61           # x == y
    hostlib.vec_destroy(x); hostlib.vec_destroy(y)
    hostlib.chk(L.MatDestroy(C.byref(A)))


def test_no_gpu_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    s = hostlib.System(4)
    y = hostlib.vec_from(np.zeros(64))
    rc = hostlib.lib().MatMult(s.A, s.exact, y)
    assert rc == 92
    hostlib.vec_destroy(y)
    s.destroy()


def test_reference_driver_compiles_unmodified():
    """When the reference tree is mounted, its own main_ksp.cpp + helper.cpp build against the shim."""
    if not os.path.exists("/root/reference/src/main_ksp.cpp"):
        pytest.skip("reference tree not mounted")
    assert os.path.exists(os.path.join(hostlib.BIN, "ref_main_ksp"))
    out = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(hostlib.BIN, "ref_main_ksp")], capture_output=True, text=True).stdout
    for sym in ("KSPSolve", "MatSetValues", "DMDACreate3d", "MatZeroRowsColumns"):
        assert sym in out


REF_HELPER = os.path.join(hostlib.ROOT, "oracle", "_ref", "libref_helper.so")


@pytest.mark.skipif(not os.path.exists(REF_HELPER), reason="oracle/_ref/libref_helper.so is built only where the reference tree is mounted")
@pytest.mark.parametrize("dims", [(5, 5, 5), (12, 12, 12), (6, 4, 9), (30, 30, 30)])
def test_reference_own_problem_builder_gives_the_oracle_bits(dims):
    """The reference's OWN src/helper.cpp (createSystem -> generateRHS/generateExt/generateA/
    setRefPoint), compiled unmodified from where it lies into oracle/_ref, run on this host layer's
    DMDA / MatSetValues / MatZeroRowsColumns: the assembled CSR, the right-hand side and the exact
    solution are bit-identical to the oracle's restatement -- which pins both the restatement and
    the host layer to the reference's code for row A9 of the scope table."""
    import ctypes as C
    hostlib.lib()   # libb200aij / libb200petsc first
    R = C.CDLL(REF_HELPER)
    nx, ny, nz = (C.c_int(d) for d in dims)
    da, A, lhs, rhs, exact = (C.c_void_p(0) for _ in range(5))
    hostlib.chk(R.createSystem(C.byref(nx), C.byref(ny), C.byref(nz), C.byref(da), C.byref(A), C.byref(lhs), C.byref(rhs),
                               C.byref(exact)))
    n = dims[0] * dims[1] * dims[2]
    m, nn, nnz = C.c_int(0), C.c_int(0), C.c_int(0)
    pi, pj, pa = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
    hostlib.chk(hostlib.lib().MatSeqAIJGetCSRB200(A, C.byref(m), C.byref(nn), C.byref(nnz), C.byref(pi), C.byref(pj), C.byref(pa)))
    ai = np.ctypeslib.as_array(pi, shape=(m.value + 1,)).copy()
    aj = np.ctypeslib.as_array(pj, shape=(nnz.value,)).copy()
    aa = np.ctypeslib.as_array(pa, shape=(nnz.value,)).copy()
    p = oracle.poisson7(*dims)
    assert m.value == n and np.array_equal(ai, p["ai"]) and np.array_equal(aj, p["aj"])
    assert np.array_equal(aa, p["aa"])
    assert np.array_equal(hostlib.vec_array(rhs, n), p["rhs"])
    assert np.array_equal(hostlib.vec_array(exact, n), p["exact"])
    assert np.array_equal(hostlib.vec_array(lhs, n), np.zeros(n))
    hostlib.chk(R.destroySystem(C.byref(da), C.byref(A), C.byref(lhs), C.byref(rhs), C.byref(exact)))
