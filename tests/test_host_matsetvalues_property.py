"""Property tests (hypothesis) of the host-side MatSetValues / MatAssemblyEnd_SeqAIJ route against a
dictionary model: arbitrary insertion order, INSERT and ADD modes, repeated entries, negative
(ignored) indices, over- and under-preallocation.  Integer work must match exactly."""
import ctypes as C

import numpy as np
from hypothesis import given, settings, strategies as st

import hostlib
import oracle

INSERT, ADD = 1, 2


@st.composite
def programs(draw):
    m = draw(st.integers(1, 12))
    n = draw(st.integers(1, 12))
    nnz = [draw(st.integers(0, 6)) for _ in range(m)]
    ops = draw(st.lists(st.tuples(st.integers(-1, m - 1), st.integers(-1, n - 1),
                                  st.floats(-4, 4, allow_nan=False, width=32), st.sampled_from([INSERT, ADD])),
                        min_size=0, max_size=80))
    return m, n, nnz, ops


@settings(max_examples=120, deadline=None)
@given(programs())
def test_matsetvalues_matches_dictionary_model(prog):
    m, n, nnz, ops = prog
    L = hostlib.lib()
    A = C.c_void_p(0)
    pre = np.array(nnz, np.int32)
    hostlib.chk(L.MatCreateSeqAIJ(2, m, n, 0, pre.ctypes.data_as(C.c_void_p), C.byref(A)))
    model = {}
    for (i, j, v, mode) in ops:
        row, col, val = np.array([i], np.int32), np.array([j], np.int32), np.array([float(v)])
        hostlib.chk(L.MatSetValues(A, 1, row.ctypes.data_as(C.c_void_p), 1, col.ctypes.data_as(C.c_void_p),
                                   val.ctypes.data_as(C.c_void_p), mode))
        if i < 0 or j < 0:
            continue  # negative indices are ignored (this is how ghost neighbours are dropped)
        if mode == ADD and (i, j) in model:
            model[(i, j)] = model[(i, j)] + float(v)
        else:
            model[(i, j)] = float(v)
    hostlib.chk(L.MatAssemblyBegin(A, 0))
    hostlib.chk(L.MatAssemblyEnd(A, 0))
    mm, nn, nz = C.c_int(0), C.c_int(0), C.c_int(0)
    pi, pj, pa = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
    hostlib.chk(L.MatSeqAIJGetCSRB200(A, C.byref(mm), C.byref(nn), C.byref(nz), C.byref(pi), C.byref(pj), C.byref(pa)))
    assert (mm.value, nn.value, nz.value) == (m, n, len(model))
    ai = np.ctypeslib.as_array(pi, shape=(m + 1,))
    aj = np.ctypeslib.as_array(pj, shape=(max(nz.value, 1),))[:nz.value]
    aa = np.ctypeslib.as_array(pa, shape=(max(nz.value, 1),))[:nz.value]
    keys = sorted(model)
    assert list(aj) == [j for (_, j) in keys]
    assert list(np.diff(ai)) == [sum(1 for (i, _) in keys if i == r) for r in range(m)]
    assert list(aa) == [model[k] for k in keys]
    a, b, c, d, e = (C.c_int(0) for _ in range(5))
    hostlib.chk(L.MatSeqAIJGetInfoB200(A, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e)))
    lens = np.diff(ai)
    assert a.value == int((lens > 0).sum()) and b.value == (int(lens.max()) if m else 0)
    assert bool(c.value) == ((m - a.value) >= 0.6 * m)
    hostlib.chk(L.MatDestroy(C.byref(A)))


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 7), st.integers(2, 7), st.integers(2, 7))
def test_dmda_matrix_pattern_is_clipped_star(nx, ny, nz_):
    """DMDACreate3d + DMCreateMatrix + generateA on non-cubic grids: 7N^3-6N^2 generalised."""
    L = hostlib.lib()
    # non-cubic grids go through the options database like the reference's -da_grid_* flags
    for k, v in (("-da_grid_x", nx), ("-da_grid_y", ny), ("-da_grid_z", nz_)):
        hostlib.chk(L.PetscOptionsSetValue(None, k.encode(), str(v).encode()))
    da, A, lhs, rhs, ex = (C.c_void_p(0) for _ in range(5))
    hostlib.chk(L.b200_create_poisson_system(C.c_int(-4), C.byref(da), C.byref(A), C.byref(lhs), C.byref(rhs), C.byref(ex)))
    hostlib.chk(L.PetscOptionsClear(None))
    m, n, nz = C.c_int(0), C.c_int(0), C.c_int(0)
    pi, pj, pa = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
    hostlib.chk(L.MatSeqAIJGetCSRB200(A, C.byref(m), C.byref(n), C.byref(nz), C.byref(pi), C.byref(pj), C.byref(pa)))
    cells = nx * ny * nz_
    assert m.value == cells
    assert nz.value == 7 * cells - 2 * (ny * nz_ + nx * nz_ + nx * ny)
    aj = np.ctypeslib.as_array(pj, shape=(nz.value,))
    ai = np.ctypeslib.as_array(pi, shape=(cells + 1,))
    for r in range(cells):
        row = aj[ai[r]:ai[r + 1]]
        assert np.all(np.diff(row) > 0)
    # and bit for bit what the oracle's restatement of src/helper.cpp builds (values, rhs, exact)
    o = oracle.poisson7(nx, ny, nz_)
    aa = np.ctypeslib.as_array(pa, shape=(nz.value,))
    assert np.array_equal(ai, o["ai"]) and np.array_equal(aj, o["aj"]) and np.array_equal(aa, o["aa"])
    assert np.array_equal(hostlib.vec_array(rhs, cells), o["rhs"]) and np.array_equal(hostlib.vec_array(ex, cells), o["exact"])
    hostlib.chk(L.b200_destroy_poisson_system(C.byref(da), C.byref(A), C.byref(lhs), C.byref(rhs), C.byref(ex)))
