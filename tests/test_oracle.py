"""CPU tests of the oracle: against the committed golden checksums, against independent
restatements (numpy / scipy), and against the reference's only known-answer test (the analytic
solution of src/main_ksp.cpp:5-15,120-121)."""
import hashlib
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

import gen
import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "poisson7.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("case", GOLD["seq"], ids=lambda c: f"N{c['N']}")
def test_golden_seq(case):
    N = case["N"]
    p = oracle.poisson7(N)
    assert int(p["ai"][-1]) == case["nnz"] == 7 * N ** 3 - 6 * N ** 2
    for k in ("ai", "aj", "aa", "rhs", "exact"):
        assert sha(p[k]) == case[k], k
    assert p["scale"].hex() == case["scale"]
    x = gen.uniform_pm1(N ** 3, seed=0xB200)
    assert sha(oracle.matmult(p["ai"], p["aj"], p["aa"], x)) == case["y_rand"]
    assert sha(oracle.matmult(p["ai"], p["aj"], p["aa"], x, fma=True)) == case["y_rand_fma"]
    assert sha(oracle.matmult(p["ai"], p["aj"], p["aa"], p["exact"])) == case["y_exact"]
    assert sha(oracle.matmulttranspose(p["ai"], p["aj"], p["aa"], x, N ** 3)) == case["yt_rand"]


needs_ref = pytest.mark.skipif(oracle.ref_lib() is None,
                               reason="oracle/_ref is built only where the reference tree is mounted (make -C oracle)")


@needs_ref
@pytest.mark.parametrize("case", GOLD["seq"], ids=lambda c: f"N{c['N']}")
def test_reference_loop_text_reproduces_the_golden_checksums(case):
    """oracle/_ref = the reference's OWN row loops (the PETSc 3.7.6 loop on the old side of
    src/openacc-step1/MatMult_SeqAIJ.patch:19-32, and the author's host + device loops of
    src/openacc-step3/MatMult_SeqAIJ.patch:36-70) compiled from where they lie.  They give the bits
    the golden file holds -- i.e. the goldens are the reference's outputs, not only the port's."""
    N = case["N"]
    p = oracle.poisson7(N)
    x = gen.uniform_pm1(N ** 3, seed=0xB200)
    assert sha(oracle.ref_matmult(p["ai"], p["aj"], p["aa"], x)) == case["y_rand"]
    assert sha(oracle.ref_matmult(p["ai"], p["aj"], p["aa"], p["exact"])) == case["y_exact"]
    for host_rows in (0, 1, N ** 3 // 3, N ** 3):   # where the "transfer" ends the host loop
        assert sha(oracle.ref_matmult(p["ai"], p["aj"], p["aa"], x, "step3", host_rows)) == case["y_rand"]
        assert sha(oracle.ref_matmult(p["ai"], p["aj"], p["aa"], x, "step4", host_rows)) == case["y_rand"]
    assert sha(oracle.ref_matmult_mt(p["ai"], p["aj"], p["aa"], x, 7)) == case["y_rand"]


@needs_ref
def test_reference_blocked_loop_with_a_full_block():
    """Step 4 cuts the rows into blocks of 983,040: a matrix long enough for one whole block plus a
    remainder, host loop stopped early -- same bits as the plain loop."""
    ai, aj, aa = gen.poisson7_natural(110, refpoint=False)      # 1,331,000 rows
    x = gen.uniform_pm1(110 ** 3, 5)
    ref = oracle.matmult(ai, aj, aa, x)
    for host_rows in (0, 1000):
        assert np.array_equal(oracle.ref_matmult(ai, aj, aa, x, "step4", host_rows), ref)


@needs_ref
def test_port_equals_the_reference_loop_on_irregular_matrices():
    rng = np.random.default_rng(11)
    cases = [gen.random_csr(700, 500, 40, rng, empty_frac=0.3), gen.powerlaw(5000, lmax=800), gen.stencil27(7, seed=2)]
    for ai, aj, aa in cases:
        n = int(aj.max()) + 1 if len(aj) else 1
        x = gen.uniform_pm1(n, 9)
        ref = oracle.ref_matmult(ai, aj, aa, x)
        assert np.array_equal(oracle.matmult(ai, aj, aa, x), ref)
        assert np.array_equal(oracle.matmult_mt(ai, aj, aa, x, 3), ref)
        assert np.array_equal(oracle.ref_matmult(ai, aj, aa, x, "step3", len(ai) // 2), ref)
        # MatMultAdd restates the same loop with the accumulator started at y: adding to zero is the same sum
        assert np.array_equal(oracle.matmultadd(ai, aj, aa, x, np.zeros(len(ai) - 1)), ref)


def _pin_cases():
    rng = np.random.default_rng(23)
    p = oracle.poisson7(12)
    return {
        "poisson7_12": (p["ai"], p["aj"], p["aa"], 12 ** 3),
        "stencil27_7": (*gen.stencil27(7, seed=2), 7 ** 3),
        "powerlaw_5k": (*gen.powerlaw(5000, lmax=800), 5000),
        "ragged_nonsquare": (*gen.random_csr(700, 500, 40, rng, empty_frac=0.3), 500),
        "mostly_empty": (*gen.random_csr(2000, 300, 3, rng, empty_frac=0.9), 300),
    }


@needs_ref
@pytest.mark.parametrize("name", sorted(_pin_cases()))
def test_transpose_add_and_compressed_row_pinned_to_the_reference_loop(name):
    """MatMultAdd / MatMultTranspose[Add] / the compressed-row branch have no text in the reference;
    what it does hold is the MatMult row loop (oracle/_ref).  Each restatement is tied to that loop by
    feeding it an equivalent matrix built independently (scipy):
      * transpose: A^T as CSR with ascending columns -- row c of A^T lists column c's entries in
        ascending row order, the order in which the scatter loop y[aj[k]] += x[i]*aa[k] adds them;
      * add: [I | A] applied to [y; x] -- the row sum starts 0.0 + 1.0*y_i = y_i, then the same adds;
      * compressed row: the loop over the non-empty rows only, scattered through rindex."""
    ai, aj, aa, n = _pin_cases()[name]
    m = len(ai) - 1
    A = sp.csr_matrix((aa, aj, ai), shape=(m, n))
    x, xt = gen.uniform_pm1(n, 31), gen.uniform_pm1(m, 32)
    y0, z0 = gen.uniform_pm1(m, 33), gen.uniform_pm1(n, 34)
    # --- transpose ---------------------------------------------------------------------------
    T = A.T.tocsr()
    T.sort_indices()
    want_t = oracle.ref_matmult(T.indptr.astype(np.int32), T.indices.astype(np.int32), T.data, xt)
    assert np.array_equal(oracle.matmulttranspose(ai, aj, aa, xt, n), want_t)
    # --- add: [I | A] [y; x] -----------------------------------------------------------------
    def augmented(M, rows):
        I = sp.identity(rows, format="csr")
        G = sp.hstack([I, M]).tocsr()
        G.sort_indices()
        return G.indptr.astype(np.int32), G.indices.astype(np.int32), G.data
    gi, gj, ga = augmented(A, m)
    assert np.array_equal(oracle.matmultadd(ai, aj, aa, x, y0), oracle.ref_matmult(gi, gj, ga, np.concatenate([y0, x])))
    gi, gj, ga = augmented(T, n)
    assert np.array_equal(oracle.matmulttransposeadd(ai, aj, aa, xt, z0, n), oracle.ref_matmult(gi, gj, ga, np.concatenate([z0, xt])))
    # --- compressed row ----------------------------------------------------------------------
    nzr = int((np.diff(ai) > 0).sum())
    use, cpi, ridx = oracle.check_compressed_row(ai, nzr)
    if use:
        compact = oracle.ref_matmult(cpi, aj, aa, x)            # the loop over the non-empty rows
        want = np.zeros(m)
        want[ridx] = compact
        assert np.array_equal(oracle.matmult_cprow(m, cpi, ridx, aj, aa, x), want)
        assert np.array_equal(oracle.matmult(ai, aj, aa, x), want)
        ci, cj, ca = augmented(sp.csr_matrix((aa, aj, cpi), shape=(len(ridx), n)), len(ridx))
        wadd = y0.copy()
        wadd[ridx] = oracle.ref_matmult(ci, cj, ca, np.concatenate([y0[ridx], x]))
        assert np.array_equal(oracle.matmultadd_cprow(m, cpi, ridx, aj, aa, x, y0), wadd)
    else:
        assert name != "mostly_empty"


@pytest.mark.parametrize("case", GOLD["mpi"], ids=lambda c: f"N{c['N']}x{c['size']}")
def test_golden_mpi(case):
    N, size = case["N"], case["size"]
    assert list(oracle.dmda_decide(N, N, N, size)) == case["grid"]
    base = oracle.dmda_bases(N, N, N, size)
    assert [int(b) for b in base] == case["bases"]
    for r, g in enumerate(case["ranks"]):
        p = oracle.poisson7(N, size=size, rank=r)
        (Ai, Aj, Aa), (Bi, Bj, Ba) = oracle.mpiaij_split(p["ai"], p["aj"], p["aa"], p["rstart"], p["rend"])
        Bjc, garray = oracle.mpiaij_setup_multiply(Bj)
        assert (len(Aj), len(Bj), len(garray)) == (g["A_nnz"], g["B_nnz"], g["nghost"])
        for k, v in (("Ai", Ai), ("Aj", Aj), ("Aa", Aa), ("Bi", Bi), ("Bj", Bjc), ("Ba", Ba), ("garray", garray)):
            assert sha(v) == g[k], k
        assert [int(v) for v in oracle.scatter_recv_offsets(base, garray)] == g["recv_off"]


def test_survey_sizes():
    """process grids and the 300^3 value 1/(dx*dx) = 89999.99999999999."""
    for s, g in GOLD["decide_300"].items():
        assert list(oracle.dmda_decide(300, 300, 300, int(s))) == g
    assert 1.0 / ((1.0 / 300) * (1.0 / 300)) == 89999.99999999999
    # 8-rank layout of a 30^3 grid scales like the survey's 300^3 numbers (faces 3*15^2 etc.)
    p = oracle.poisson7(30, size=8, rank=0)
    (Ai, Aj, Aa), (Bi, Bj, Ba) = oracle.mpiaij_split(p["ai"], p["aj"], p["aa"], p["rstart"], p["rend"])
    Bjc, garray = oracle.mpiaij_setup_multiply(Bj)
    assert len(garray) == 3 * 15 * 15 and len(Bj) == 3 * 15 * 15
    assert int((np.diff(Bi) > 0).sum()) == 3 * 15 * 15 - 3 * 15 + 1
    assert len(Aj) == 7 * 15 ** 3 - 6 * 15 ** 2


@pytest.mark.parametrize("N", [3, 4, 7, 20])
def test_generator_matches_numpy_restatement(N):
    ai, aj, aa = gen.poisson7_natural(N)
    p = oracle.poisson7(N)
    assert np.array_equal(ai, p["ai"]) and np.array_equal(aj, p["aj"]) and np.array_equal(aa, p["aa"])


def test_known_answer_analytic_solution():
    """The reference's own check: the discrete solution converges to cos*cos*cos at O(h^2)."""
    errs = []
    for N in (8, 16, 32):
        p = oracle.poisson7(N)
        A = sp.csr_matrix((p["aa"], p["aj"], p["ai"]), shape=(N ** 3, N ** 3)).tocsc()
        u = spl.spsolve(A, p["rhs"])
        errs.append(np.abs(u - p["exact"]).max())
    assert errs[0] > errs[1] > errs[2]
    assert 3.0 < errs[0] / errs[1] < 5.0 and 3.0 < errs[1] / errs[2] < 5.0
    assert errs[2] < 0.01


def test_matmult_family_against_scipy_and_loops():
    rng = np.random.default_rng(3)
    for m, n, d, e in ((50, 40, 6, 0.2), (300, 300, 30, 0.0), (200, 10, 3, 0.8)):
        ai, aj, aa = gen.random_csr(m, n, d, rng, empty_frac=e)
        A = sp.csr_matrix((aa, aj, ai), shape=(m, n))
        x, y0, xt = rng.uniform(-1, 1, n), rng.uniform(-1, 1, m), rng.uniform(-1, 1, m)
        # python-loop restatement, strict left to right
        ref = np.zeros(m)
        for i in range(m):
            s = 0.0
            for k in range(ai[i], ai[i + 1]):
                s += aa[k] * x[aj[k]]
            ref[i] = s
        assert np.array_equal(oracle.matmult(ai, aj, aa, x), ref)
        np.testing.assert_allclose(oracle.matmult(ai, aj, aa, x), A @ x, rtol=0, atol=1e-13)
        np.testing.assert_allclose(oracle.matmult(ai, aj, aa, x, fma=True), ref, rtol=0, atol=1e-13)
        refadd = np.zeros(m)
        for i in range(m):
            s = y0[i]
            for k in range(ai[i], ai[i + 1]):
                s += aa[k] * x[aj[k]]
            refadd[i] = s
        assert np.array_equal(oracle.matmultadd(ai, aj, aa, x, y0), refadd)
        reft = np.zeros(n)
        for i in range(m):
            for k in range(ai[i], ai[i + 1]):
                reft[aj[k]] += xt[i] * aa[k]
        assert np.array_equal(oracle.matmulttranspose(ai, aj, aa, xt, n), reft)
        z = rng.uniform(-1, 1, n)
        refta = z.copy()
        for i in range(m):
            for k in range(ai[i], ai[i + 1]):
                refta[aj[k]] += xt[i] * aa[k]
        assert np.array_equal(oracle.matmulttransposeadd(ai, aj, aa, xt, z, n), refta)
        # compressed-row variants give the same numbers as the plain ones
        nzr = int((np.diff(ai) > 0).sum())
        use, cpi, ridx = oracle.check_compressed_row(ai, nzr)
        assert use == (m - nzr >= 0.6 * m)
        if use:
            assert np.array_equal(oracle.matmult_cprow(m, cpi, ridx, aj, aa, x), ref)
            assert np.array_equal(oracle.matmultadd_cprow(m, cpi, ridx, aj, aa, x, y0), refadd)
        assert oracle.matmult_flops(len(aj), nzr) == 2.0 * len(aj) - nzr


def test_multithreaded_baseline_identical():
    p = oracle.poisson7(20)
    x = gen.uniform_pm1(20 ** 3)
    ref = oracle.matmult(p["ai"], p["aj"], p["aa"], x)
    for t in (1, 3, 8):
        assert np.array_equal(oracle.matmult_mt(p["ai"], p["aj"], p["aa"], x, t), ref)


def test_assembly_end_compaction():
    """MatAssemblyEnd_SeqAIJ: rows stored with slack are packed; counters follow."""
    rng = np.random.default_rng(5)
    m = 200
    imax = rng.integers(0, 9, size=m).astype(np.int32)
    ailen = np.array([rng.integers(0, mx + 1) for mx in imax], dtype=np.int32)
    ai = np.zeros(m + 1, dtype=np.int32)
    np.cumsum(imax, out=ai[1:])
    aj = np.full(ai[-1], -7, dtype=np.int32)
    aa = np.full(ai[-1], np.nan)
    rows = []
    for i in range(m):
        c = np.sort(rng.choice(1000, size=ailen[i], replace=False)).astype(np.int32)
        v = rng.uniform(-1, 1, size=ailen[i])
        aj[ai[i]:ai[i] + ailen[i]] = c
        aa[ai[i]:ai[i] + ailen[i]] = v
        rows.append((c, v))
    unused = int(imax.sum() - ailen.sum())
    nz, nzr, rmax, fshift = oracle.assembly_end(ai, aj, aa, imax, ailen)
    assert nz == sum(len(c) for c, _ in rows) == ai[-1]
    assert nzr == sum(len(c) > 0 for c, _ in rows)
    assert rmax == max(len(c) for c, _ in rows)
    assert fshift == unused
    for i, (c, v) in enumerate(rows):
        assert np.array_equal(aj[ai[i]:ai[i + 1]], c) and np.array_equal(aa[ai[i]:ai[i + 1]], v)
    assert np.array_equal(imax, np.diff(ai)) and np.array_equal(ailen, np.diff(ai))


def test_mpiaij_reassembles_to_sequential_operator():
    """A x_local + B x_ghost over all ranks equals the 1-rank operator permuted to PETSc order."""
    N, size = 10, 8
    base = oracle.dmda_bases(N, N, N, size)
    n = N ** 3
    rng = np.random.default_rng(1)
    xg = rng.uniform(-1, 1, n)  # in PETSc (rank-major) ordering
    # natural -> PETSc ordering permutation
    perm = np.zeros(n, dtype=np.int64)
    for r in range(size):
        inf = oracle.dmda_info(N, N, N, size, r)
        idx = 0
        for k in range(inf["zs"], inf["zs"] + inf["zm"]):
            for j in range(inf["ys"], inf["ys"] + inf["ym"]):
                for i in range(inf["xs"], inf["xs"] + inf["xm"]):
                    perm[i + j * N + k * N * N] = base[r] + idx
                    idx += 1
    p1 = oracle.poisson7(N)
    A1 = sp.csr_matrix((p1["aa"], p1["aj"], p1["ai"]), shape=(n, n))
    xnat = xg[perm]
    ynat = A1 @ xnat
    for r in range(size):
        p = oracle.poisson7(N, size=size, rank=r)
        (Ai, Aj, Aa), (Bi, Bj, Ba) = oracle.mpiaij_split(p["ai"], p["aj"], p["aa"], p["rstart"], p["rend"])
        Bjc, garray = oracle.mpiaij_setup_multiply(Bj)
        assert np.all(np.diff(garray) > 0)
        y = oracle.matmult(Ai, Aj, Aa, xg[p["rstart"]:p["rend"]])
        y = oracle.matmultadd(Bi, Bjc, Ba, xg[garray], y)
        inv = np.argsort(perm)  # PETSc index -> natural index
        np.testing.assert_allclose(y, ynat[inv[p["rstart"]:p["rend"]]], rtol=0, atol=1e-9)
        # rhs/exact are the same fields in the other ordering
        assert np.array_equal(p["exact"], p1["exact"][inv[p["rstart"]:p["rend"]]])


def test_cg_jacobi_solves_reference_problem():
    p = oracle.poisson7(12)
    x, its, rn = oracle.cg_jacobi(p["ai"], p["aj"], p["aa"], p["rhs"], rtol=1e-12, atol=1e-50)
    assert its > 0
    A = sp.csr_matrix((p["aa"], p["aj"], p["ai"]))
    assert np.linalg.norm(A @ x - p["rhs"]) / np.linalg.norm(p["rhs"]) < 1e-9
    assert np.abs(x - p["exact"]).max() < 0.06


def test_golden_synthetic_workloads(pk):
    """27-point stencil, power-law matrix and the splitmix vector: numpy definition, product
    generator and oracle results against the committed checksums."""
    g = GOLD["synthetic"]
    ai, aj, aa = gen.stencil27(9)
    assert len(aj) == g["stencil27_9"]["nnz"] == (3 * 9 - 2) ** 3
    assert (sha(ai), sha(aj), sha(aa)) == (g["stencil27_9"]["ai"], g["stencil27_9"]["aj"], g["stencil27_9"]["aa"])
    p = g["powerlaw_20000_3000"]
    for src in (gen.powerlaw, pk.gen_powerlaw):
        ai, aj, aa = src(20000, lmax=3000)
        assert (len(aj), int(np.diff(ai).max())) == (p["nnz"], p["rmax"])
        assert (sha(ai), sha(aj), sha(aa)) == (p["ai"], p["aj"], p["aa"])
    x = gen.uniform_pm1(20000, seed=0xB200)
    assert sha(oracle.matmult(ai, aj, aa, x)) == p["y"]
    assert sha(oracle.matmulttranspose(ai, aj, aa, x, 20000)) == p["yt"]
    assert sha(gen.uniform_pm1(16, seed=0xB200)) == g["x_b200_16"] == sha(pk.gen_vector(16, 0xB200))
