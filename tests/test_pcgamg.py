"""`-pc_type gamg` of the host layer (petsc-openacc_b200/host/src/pcgamg.cpp) for the reference's
solver options (configs/PETSc_SolverOptions_GAMG.info).

CPU tests: the host set-up (strength graph, aggregates, prolongator, Galerkin operator) against the
independent numpy/scipy restatement in oracle/gamg.py -- integers bit-exact, values to 1e-12 --
and structural properties of a smoothed-aggregation hierarchy.
GPU tests: the V-cycle on the device against the C restatement (orc_mg_apply) BIT-EXACT -- every
level operation is a MatMult / MatMultAdd / MatMultTranspose / fused sweep of the hot path -- and
the CG+GAMG solve against orc_cg_mg (iteration count, residual).

PCGAMG has no text in the reference (PETSc 3.7.6 is downloaded by its build): PARITY UNPINNED
against PETSc itself; these tests pin the product to this repository's own restatement."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

import gen
import hostlib
import oracle
from oracle import gamg

GAMG_OPTIONS = {
    "-ksp_type": "cg", "-ksp_atol": "1e-12", "-ksp_rtol": "1e-14", "-ksp_max_it": "10000",
    "-pc_type": "gamg", "-pc_gamg_type": "agg", "-pc_gamg_agg_nsmooths": "1", "-pc_gamg_threshold": "0.0",
    "-mg_coarse_ksp_type": "preonly", "-mg_coarse_pc_type": "bjacobi", "-mg_coarse_sub_pc_type": "jacobi",
    "-mg_levels_ksp_type": "richardson", "-mg_levels_ksp_max_it": "1", "-mg_levels_pc_type": "bjacobi",
    "-mg_levels_sub_pc_type": "jacobi",
}


def set_options(extra=None):
    L = hostlib.lib()
    hostlib.chk(L.PetscOptionsClear(None))
    for k, v in {**GAMG_OPTIONS, **(extra or {})}.items():
        hostlib.chk(L.PetscOptionsSetValue(None, k.encode(), str(v).encode()))


def mat_csr(A):
    L = hostlib.lib()
    m, n, nz = C.c_int(0), C.c_int(0), C.c_int(0)
    pi, pj, pa = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
    hostlib.chk(L.MatSeqAIJGetCSRB200(A, C.byref(m), C.byref(n), C.byref(nz), C.byref(pi), C.byref(pj), C.byref(pa)))
    ai = np.ctypeslib.as_array(pi, shape=(m.value + 1,)).copy()
    aj = np.ctypeslib.as_array(pj, shape=(max(nz.value, 1),))[:nz.value].copy()
    aa = np.ctypeslib.as_array(pa, shape=(max(nz.value, 1),))[:nz.value].copy()
    return ai, aj, aa, n.value


class Solver:
    """KSPCreate + KSPSetOperators + KSPSetFromOptions + KSPSetUp on a Mat of the host layer."""

    def __init__(self, A, extra=None):
        L = hostlib.lib()
        set_options(extra)
        self.ksp = C.c_void_p(0)
        hostlib.chk(L.KSPCreate(C.c_int(1), C.byref(self.ksp)))
        hostlib.chk(L.KSPSetOperators(self.ksp, A, A))
        hostlib.chk(L.KSPSetFromOptions(self.ksp))
        self.setup_rc = L.KSPSetUp(self.ksp)

    def levels(self):
        L = hostlib.lib()
        n = C.c_int(0)
        hostlib.chk(L.PCGAMGGetNumLevelsB200(self.ksp, C.byref(n)))
        out = []
        for l in range(n.value):
            A, P, dinv = C.c_void_p(0), C.c_void_p(0), C.c_void_p(0)
            agg, nagg, emax = C.POINTER(C.c_int)(), C.c_int(0), C.c_double(0.0)
            hostlib.chk(L.PCGAMGGetLevelB200(self.ksp, C.c_int(l), C.byref(A), C.byref(P), C.byref(dinv), C.byref(agg),
                                             C.byref(nagg), C.byref(emax)))
            ai, aj, aa, ncol = mat_csr(A)
            lv = dict(A=(ai, aj, aa), m=len(ai) - 1, P=None, agg=None, nagg=nagg.value, emax=emax.value,
                      dinv=hostlib.vec_array(dinv, len(ai) - 1))
            if P:
                pi, pj, pa, pn = mat_csr(P)
                lv["P"] = (pi, pj, pa)
                lv["pn"] = pn
                lv["agg"] = np.ctypeslib.as_array(agg, shape=(len(ai) - 1,)).copy()
            out.append(lv)
        return out

    def destroy(self):
        hostlib.chk(hostlib.lib().KSPDestroy(C.byref(self.ksp)))
        hostlib.chk(hostlib.lib().PetscOptionsClear(None))


def as_sp(t, n=None):
    ai, aj, aa = t
    return sp.csr_matrix((aa, aj, ai), shape=(len(ai) - 1, n or len(ai) - 1))


# ---- CPU: the set-up ------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [6, 11])
def test_setup_matches_the_python_restatement(N):
    s = hostlib.System(N)
    sv = Solver(s.A, {"-pc_gamg_b200_esteig": "gershgorin"})
    assert sv.setup_rc == 0
    got = sv.levels()
    ai, aj, aa = s.csr()
    want = gamg.hierarchy(ai, aj, aa, esteig="gershgorin")
    assert len(got) == len(want) >= 2
    for g, w in zip(got, want):
        assert g["m"] == w["A"].shape[0]
        assert g["nagg"] == w["nagg"]
        if w["P"] is None:
            assert g["P"] is None
            continue
        assert np.array_equal(g["agg"], w["agg"])                      # integer work: bit-exact
        assert g["emax"] == pytest.approx(w["emax"], rel=1e-14)
        # the product keeps the structural zeros of a sparse product (the zeroed row/column 0 of the
        # reference matrix keeps its pattern, src/helper.cpp:264-274); scipy drops them
        P = as_sp(g["P"], g["pn"])
        P.eliminate_zeros()
        assert np.array_equal(P.indptr, w["P"].indptr) and np.array_equal(P.indices, w["P"].indices)
        assert np.allclose(P.data, w["P"].data, rtol=1e-12, atol=0)
    for g, w in zip(got[1:], want[1:]):
        A = as_sp(g["A"])
        A.eliminate_zeros()
        assert abs(A - w["A"]).max() <= 1e-12 * abs(w["A"]).max()
        assert np.array_equal(A.indptr, w["A"].indptr) and np.array_equal(A.indices, w["A"].indices)
    sv.destroy()
    s.destroy()


def test_hierarchy_properties():
    """What makes it a smoothed-aggregation hierarchy, independent of any restatement."""
    N = 14
    s = hostlib.System(N)
    sv = Solver(s.A, {"-pc_gamg_b200_esteig": "gershgorin"})
    assert sv.setup_rc == 0
    lv = sv.levels()
    sizes = [l["m"] for l in lv]
    assert sizes[0] == N ** 3 and all(a > b for a, b in zip(sizes, sizes[1:]))
    assert sizes[-1] <= 50 < sizes[-2]                                   # -pc_gamg_coarse_eq_limit
    B = np.ones(sizes[0])
    for l, L in enumerate(lv[:-1]):
        A, P = as_sp(L["A"]), as_sp(L["P"], L["pn"])
        agg = L["agg"]
        # every vertex with a strong neighbour is in exactly one aggregate; ids are 0..nagg-1, all used
        assert agg.min() >= -1 and agg.max() == L["nagg"] - 1
        assert len(np.unique(agg[agg >= 0])) == L["nagg"]
        offdiag = A - sp.diags(A.diagonal())
        isolated = np.asarray((offdiag != 0).sum(axis=1)).ravel() == 0
        assert np.array_equal(agg < 0, isolated)
        # roots of the MIS are never adjacent (distance > 2 on the squared level)
        # near-null space is carried: P0 Bc = B on aggregated rows, hence P Bc = (I - w D^-1 A) B
        Bc = np.sqrt(np.bincount(agg[agg >= 0], weights=B[agg >= 0] ** 2, minlength=L["nagg"]))
        dinv = L["dinv"]
        want = np.where(agg >= 0, B, 0.0)
        want = want - (1.4 / L["emax"]) * dinv * (A @ want)
        assert np.allclose(P @ Bc, want, rtol=0, atol=1e-12 * max(1.0, np.abs(want).max()))
        # Galerkin: A_c = P^T A P, symmetric
        Ac = as_sp(lv[l + 1]["A"])
        G = (P.T @ (A @ P)).tocsr()
        assert abs(Ac - G).max() <= 1e-12 * abs(G).max()
        assert abs(Ac - Ac.T).max() <= 1e-12 * abs(Ac).max()
        assert np.array_equal(lv[l + 1]["dinv"], 1.0 / Ac.diagonal())
        B = Bc
    sv.destroy()
    s.destroy()


def test_threads_do_not_change_the_setup(monkeypatch):
    """The sparse products are summed in storage order per entry: any thread count, same bits."""
    N = 30  # 27,000 rows: above the threshold where the products go parallel
    s = hostlib.System(N)
    out = []
    for t in ("1", "5"):
        monkeypatch.setenv("B200_SETUP_THREADS", t)
        sv = Solver(s.A, {"-pc_gamg_b200_esteig": "gershgorin"})
        assert sv.setup_rc == 0
        out.append(sv.levels())
        sv.destroy()
    assert len(out[0]) == len(out[1])
    for a, b in zip(*out):
        for k in range(3):
            assert np.array_equal(a["A"][k], b["A"][k])
            if a["P"] is not None:
                assert np.array_equal(a["P"][k], b["P"][k])
    s.destroy()


def test_unsupported_multigrid_options_are_refused():
    s = hostlib.System(5)
    for extra in ({"-mg_levels_ksp_type": "chebyshev"}, {"-mg_levels_sub_pc_type": "ilu"},
                  {"-mg_coarse_pc_type": "lu"}, {"-pc_gamg_type": "geo"}, {"-pc_gamg_agg_nsmooths": "2"}):
        sv = Solver(s.A, {"-pc_gamg_b200_esteig": "gershgorin", **extra})
        assert sv.setup_rc == 56, extra  # PETSC_ERR_SUP
        sv.destroy()
    L = hostlib.lib()
    set_options({"-pc_type": "ilu"})
    ksp = C.c_void_p(0)
    hostlib.chk(L.KSPCreate(C.c_int(1), C.byref(ksp)))
    assert L.KSPSetFromOptions(ksp) == 56
    hostlib.chk(L.KSPDestroy(C.byref(ksp)))
    hostlib.chk(L.PetscOptionsClear(None))
    s.destroy()


def test_tridiagonal_emax_against_lapack():
    L = hostlib.lib()
    rng = np.random.default_rng(3)
    for n in (1, 2, 5, 10, 40):
        d = rng.uniform(0.5, 2.0, n)
        e = np.concatenate([[0.0], rng.uniform(-1.0, 1.0, n - 1)])
        T = np.diag(d) + np.diag(e[1:], 1) + np.diag(e[1:], -1)
        out = C.c_double(0.0)
        hostlib.chk(L.b200_tridiag_emax(C.c_int(n), d.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p), C.byref(out)))
        assert out.value == pytest.approx(np.linalg.eigvalsh(T)[-1], rel=1e-13)


def test_oracle_vcycle_is_a_symmetric_definite_operator():
    """The V-cycle with equal pre/post Jacobi smoothing must be symmetric and definite with the
    sign of A (the reference matrix is the NEGATIVE definite Laplacian, src/helper.cpp:188-233) --
    CG needs both.  Checked on the C restatement, which the GPU path is compared with bit for bit."""
    N = 7
    p = oracle.poisson7(N)
    lv = gamg.hierarchy(p["ai"], p["aj"], p["aa"])
    n = N ** 3
    M = np.column_stack([gamg.mg_apply(lv, np.eye(n)[:, i]) for i in range(n)])
    assert np.abs(M - M.T).max() <= 1e-12 * np.abs(M).max()
    ev = np.linalg.eigvalsh(0.5 * (M + M.T))
    assert ev[-1] < 0.0   # negative definite, like A
    A = lv[0]["A"].toarray()
    assert np.linalg.eigvalsh(0.5 * (A + A.T))[-1] < 0.0


def test_oracle_cg_mg_golden_iteration_counts():
    """Freezes the restatement: CG+V-cycle iteration counts on the reference problem
    (tests/golden/gamg.json, made by tests/golden/make_golden_gamg.py)."""
    with open(os.path.join(os.path.dirname(__file__), "golden", "gamg.json")) as f:
        gold = json.load(f)
    for case in gold["cases"]:
        p = oracle.poisson7(case["N"])
        lv = gamg.hierarchy(p["ai"], p["aj"], p["aa"], esteig=case["esteig"])
        assert [int(l["A"].shape[0]) for l in lv] == case["rows"]
        assert [int(l["A"].nnz) for l in lv] == case["nnz"]
        x, its, rn = gamg.cg_mg(lv, p["rhs"])
        assert its == case["its"]
        assert rn == pytest.approx(case["rnorm"], rel=1e-6)
        assert np.abs(x - p["exact"]).max() == pytest.approx(case["error_inf"], rel=1e-9)


# ---- GPU: the V-cycle and the solve ---------------------------------------------------------------
def _apply_pc(sv, r):
    L = hostlib.lib()
    vr, vz = hostlib.vec_from(r), hostlib.vec_from(np.full(len(r), np.nan))
    hostlib.chk(L.KSPApplyPCB200(sv.ksp, vr, vz))
    z = hostlib.vec_array(vz, len(r))
    hostlib.vec_destroy(vr)
    hostlib.vec_destroy(vz)
    return z


@pytest.mark.gpu
@pytest.mark.parametrize("N,sweeps,esteig", [(10, 1, "gershgorin"), (16, 1, "cg"), (12, 2, "cg"), (9, 3, "gershgorin")])
def test_vcycle_on_device_bit_exact_against_the_c_restatement(cuda, N, sweeps, esteig):
    s = hostlib.System(N)
    sv = Solver(s.A, {"-pc_gamg_b200_esteig": esteig, "-mg_levels_ksp_max_it": sweeps})
    assert sv.setup_rc == 0
    lv = sv.levels()
    assert len(lv) >= 2
    for seed in (21, 22):
        r = gen.uniform_pm1(N ** 3, seed)
        z = _apply_pc(sv, r)
        assert np.array_equal(z, gamg.mg_apply(lv, r, sweeps=sweeps))
    sv.destroy()
    s.destroy()


@pytest.mark.gpu
def test_device_eigenvalue_estimate_matches_the_restatement(cuda):
    N = 12
    s = hostlib.System(N)
    sv = Solver(s.A, {"-pc_gamg_b200_esteig": "cg"})
    assert sv.setup_rc == 0
    got = sv.levels()
    ai, aj, aa = s.csr()
    want = gamg.hierarchy(ai, aj, aa, esteig="cg")
    assert [g["m"] for g in got] == [w["A"].shape[0] for w in want]
    for g, w in zip(got[:-1], want[:-1]):
        assert 1.0 < g["emax"] < 2.5
        assert g["emax"] == pytest.approx(w["emax"], rel=1e-9)   # dot products reduce in another order
    sv.destroy()
    s.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("N", [12, 24])
def test_cg_gamg_solve_against_the_restatement(cuda, N):
    L = hostlib.lib()
    s = hostlib.System(N)
    sv = Solver(s.A, {"-pc_gamg_b200_esteig": "gershgorin"})
    assert sv.setup_rc == 0
    lv = sv.levels()
    hostlib.chk(L.KSPSolve(sv.ksp, s.rhs, s.lhs))
    its, reason, rn = C.c_int(0), C.c_int(0), C.c_double(0.0)
    hostlib.chk(L.KSPGetIterationNumber(sv.ksp, C.byref(its)))
    hostlib.chk(L.KSPGetConvergedReason(sv.ksp, C.byref(reason)))
    hostlib.chk(L.KSPGetResidualNorm(sv.ksp, C.byref(rn)))
    x = hostlib.vec_array(s.lhs, N ** 3)
    p = oracle.poisson7(N)
    xo, its_o, rn_o = gamg.cg_mg(lv, p["rhs"])
    assert reason.value > 0
    assert abs(its.value - its_o) <= 1, (its.value, its_o)      # the dots reduce in another order
    if its.value == its_o:
        assert abs(rn.value - rn_o) <= 1e-10 * max(1.0, rn_o / 1e-12) * 1e-2 + 1e-10
    assert np.abs(x - xo).max() <= 1e-9 * np.abs(xo).max()
    # far fewer iterations than Jacobi-CG: that is what the multigrid is for
    _, its_j, _ = oracle.cg_jacobi(p["ai"], p["aj"], p["aa"], p["rhs"])
    assert its.value * 2 < its_j
    # the reference's known-answer check (src/main_ksp.cpp:120-121): O(h^2) error
    assert np.abs(x - p["exact"]).max() < 6.5 / N ** 2
    sv.destroy()
    s.destroy()


@pytest.mark.gpu
def test_reference_driver_runs_its_own_gamg_options(cuda, tmp_path):
    """The reference's unmodified main_ksp.cpp (built against the host layer) with its own
    configs/PETSc_SolverOptions_GAMG.info text."""
    exe = os.path.join(hostlib.BIN, "ref_main_ksp")
    if not os.path.exists(exe):
        pytest.skip("ref_main_ksp is built only where the reference tree is mounted")
    cfg = tmp_path / "gamg.info"
    cfg.write_text("".join(f"{k} {v}\n" for k, v in GAMG_OPTIONS.items()))
    out = subprocess.run([exe, "-config", str(cfg), "-da_grid_x", "20", "-da_grid_y", "20", "-da_grid_z", "20"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    its = int([l for l in out.stdout.splitlines() if "iterations" in l.lower()][0].split(":")[1])
    p = oracle.poisson7(20)
    _, its_j, _ = oracle.cg_jacobi(p["ai"], p["aj"], p["aa"], p["rhs"])
    assert 0 < its * 2 < its_j


# ---- CPU: the set-up on matrices that are not the reference problem -------------------------------
def _graph_laplacian(n, extra_edges, rng, isolated=(), weights=(0.5, 2.0), shift=0.0):
    """Weighted graph Laplacian (+ shift I): a ring plus random chords; `isolated` vertices keep only
    their diagonal.  Symmetric M-matrix, ascending columns."""
    rows, cols, vals = [], [], []
    edges = {(i, (i + 1) % n) for i in range(n)}
    for _ in range(extra_edges):
        a, b = rng.integers(0, n, 2)
        if a != b:
            edges.add((min(a, b), max(a, b)))
    for a, b in edges:
        if a in isolated or b in isolated:
            continue
        w = rng.uniform(*weights)
        rows += [a, b]; cols += [b, a]; vals += [-w, -w]
    A = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    A.sum_duplicates()
    d = -np.asarray(A.sum(axis=1)).ravel() + shift
    d[list(isolated)] = 1.0
    A = (A + sp.diags(d)).tocsr()
    A.sort_indices()
    return A


@pytest.mark.parametrize("seed,n,extra,threshold", [(1, 400, 300, 0.0), (2, 900, 200, 0.0), (3, 700, 1500, 0.05), (4, 300, 0, 0.0)])
def test_setup_on_random_graph_laplacians(seed, n, extra, threshold):
    rng = np.random.default_rng(seed)
    isolated = tuple(int(v) for v in rng.choice(n, size=5, replace=False))
    A = _graph_laplacian(n, extra, rng, isolated=isolated, shift=1e-3)
    M = hostlib.mat_from_csr(A.indptr, A.indices, A.data, n)
    sv = Solver(M, {"-pc_gamg_b200_esteig": "gershgorin", "-pc_gamg_threshold": threshold})
    assert sv.setup_rc == 0
    got = sv.levels()
    want = gamg.hierarchy(A.indptr, A.indices, A.data, threshold=threshold, esteig="gershgorin")
    assert [g["m"] for g in got] == [w["A"].shape[0] for w in want]
    alone = np.nonzero(np.diff(A.indptr) == 1)[0]                    # diagonal-only rows join no aggregate
    assert set(alone) >= set(isolated) and np.array_equal(np.nonzero(got[0]["agg"] < 0)[0], alone)
    for g, w in zip(got[:-1], want[:-1]):
        assert np.array_equal(g["agg"], w["agg"])
        P = as_sp(g["P"], g["pn"])
        assert abs(P - w["P"]).max() <= 1e-12 * abs(w["P"]).max()
    for g, w in zip(got[1:], want[1:]):
        assert abs(as_sp(g["A"]) - w["A"]).max() <= 1e-12 * abs(w["A"]).max()
    # the restated CG + V-cycle converges on it, in fewer iterations than Jacobi-CG
    b = gen.uniform_pm1(n, 40 + seed)
    x, its, rn = gamg.cg_mg(got, b, rtol=1e-10, atol=1e-50)
    assert its > 0 and np.abs(A @ x - b).max() <= 1e-6 * np.abs(b).max()
    _, its_j, _ = oracle.cg_jacobi(A.indptr, A.indices, A.data, b, rtol=1e-10, atol=1e-50)
    assert its <= its_j
    sv.destroy()
    hostlib.chk(hostlib.lib().MatDestroy(C.byref(M)))


def test_setup_degenerate_operators():
    """Nothing to coarsen: a diagonal matrix (every vertex isolated) and a 1x1 operator give a
    one-level hierarchy whose "V-cycle" is the coarse Jacobi application."""
    for A in (sp.diags(np.arange(1.0, 8.0)).tocsr(), sp.csr_matrix(np.array([[4.0]]))):
        n = A.shape[0]
        M = hostlib.mat_from_csr(A.indptr, A.indices, A.data, n)
        sv = Solver(M, {"-pc_gamg_b200_esteig": "gershgorin"})
        assert sv.setup_rc == 0
        lv = sv.levels()
        assert len(lv) == 1 and lv[0]["P"] is None
        assert np.array_equal(lv[0]["dinv"], 1.0 / A.diagonal())
        r = np.arange(1.0, n + 1.0)
        assert np.array_equal(gamg.mg_apply(lv, r), r / A.diagonal())
        sv.destroy()
        hostlib.chk(hostlib.lib().MatDestroy(C.byref(M)))
