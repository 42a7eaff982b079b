"""oracle/gamg.py -- TEST INFRASTRUCTURE ONLY.

An independent numpy/scipy restatement of the smoothed-aggregation set-up that `-pc_type gamg`
(configs/PETSc_SolverOptions_GAMG.info:6-8: gamg, agg, nsmooths 1, threshold 0.0) asks for, used
to check petsc-openacc_b200/host/src/pcgamg.cpp, plus ctypes wrappers of the C V-cycle / PCG
(orc_mg_apply, orc_cg_mg in seqaij_oracle.c).

PARITY UNPINNED: PCGAMG lives in PETSc 3.7.6, which the reference downloads at build time
(scripts/petsc.sh:38-40) and which is not under /root/reference; no file of the reference holds
its text and no log records an iteration count.  What is restated here is the published structure
of PCGAMG "agg" [P376]: strength graph -> greedy MIS aggregates (squared graph on the first level)
-> tentative prolongator from the near-null-space vector -> one damped-Jacobi smoothing step
(1.4/emax) -> Galerkin product, stopping at 50 coarse rows.  Two choices are this project's own
and are stated in the product source too: vertices are visited in natural order (PETSc permutes
them randomly) and PETSc's aggregate clean-up pass is not applied.

The aggregation loops are pure Python: small cases only.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _f64, _i32, _p, lib


def _csr(ai, aj, aa, n=None):
    m = len(ai) - 1
    return sp.csr_matrix((np.asarray(aa, dtype=np.float64), np.asarray(aj), np.asarray(ai)), shape=(m, n or m))


def strength_graph(ai, aj, aa, threshold=0.0):
    """PCGAMGGraph_AGG + PCGAMGFilterGraph: adjacency lists of the strong off-diagonal entries."""
    m = len(ai) - 1
    d = np.zeros(m)
    for r in range(m):
        for k in range(ai[r], ai[r + 1]):
            if aj[k] == r:
                d[r] = aa[k]
                break
    s = np.where(d != 0.0, 1.0 / np.sqrt(np.abs(np.where(d != 0.0, d, 1.0))), 1.0)
    adj = []
    for r in range(m):
        row = []
        for k in range(ai[r], ai[r + 1]):
            c = int(aj[k])
            if c != r and abs(aa[k]) * s[r] * s[c] > threshold:
                row.append(c)
        adj.append(row)
    return adj, d


def aggregate(adj, square):
    """Greedy MIS in natural order (of the squared graph when `square`); aggregate = root + the
    undecided vertices it removes; vertices without strong neighbours get -1."""
    m = len(adj)
    taken = np.zeros(m, dtype=bool)
    agg = np.full(m, -1, dtype=np.int32)
    nagg = 0
    for v in range(m):
        if taken[v]:
            continue
        taken[v] = True
        if not adj[v]:
            continue
        agg[v] = nagg
        ball = set(adj[v])
        if square:
            for w in adj[v]:
                ball.update(adj[w])
        for u in ball:
            if not taken[u]:
                taken[u] = True
                agg[u] = nagg
        nagg += 1
    return agg, nagg


def tentative(agg, nagg, B):
    """formProl0 with one near-null-space vector: P0[v, agg[v]] = B[v]/|B restricted to agg|."""
    m = len(agg)
    inside = agg >= 0
    norm = np.sqrt(np.bincount(agg[inside], weights=B[inside] ** 2, minlength=nagg))
    rows = np.nonzero(inside)[0]
    vals = B[rows] / norm[agg[rows]]
    P0 = sp.csr_matrix((vals, (rows, agg[rows])), shape=(m, nagg))
    return P0, norm


def gershgorin_emax(A, d):
    s = np.asarray(abs(A).sum(axis=1)).ravel()
    ok = d != 0.0
    return float(np.max(s[ok] / np.abs(d[ok])))


def lanczos_emax(A, d, its=10, seed=0x6A36):
    """KSPCG + PCJACOBI, KSP_NORM_NONE, `its` iterations on splitmix64 noise; largest eigenvalue of
    KSPCG's tridiagonal (d_i = b_i/a_{i-1} + 1/a_i, e_i = sqrt(b_i)/a_{i-1})."""
    m = A.shape[0]
    i = np.arange(m, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) ^ i
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    b = (z >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0
    dinv = np.where(d != 0.0, 1.0 / np.where(d != 0.0, d, 1.0), 1.0)
    r = b.copy()
    zz = dinv * r
    beta = zz @ r
    betaold, a = 1.0, 1.0
    p = zz.copy()
    dd, ee = [], []
    for it in range(its):
        bq, e = 0.0, 0.0
        if it > 0:
            bq = beta / betaold
            e = np.sqrt(abs(bq)) / a
            p = zz + bq * p
        betaold = beta
        w = A @ p
        a = beta / (p @ w)
        ee.append(e)
        dd.append(np.sqrt(abs(bq)) * e + 1.0 / a)
        r = r - a * w
        zz = dinv * r
        beta = zz @ r
    T = np.diag(dd) + np.diag(ee[1:], 1) + np.diag(ee[1:], -1)
    return float(np.linalg.eigvalsh(T)[-1])


def hierarchy(ai, aj, aa, threshold=0.0, nsmooths=1, coarse_eq_limit=50, max_levels=30, square_graph=1,
              esteig="gershgorin", est_its=10):
    """Returns a list of levels, finest first: dict(A=csr, P=csr or None, agg, nagg, emax)."""
    A = _csr(ai, aj, aa)
    A.sort_indices()
    B = np.ones(A.shape[0])
    levels = [dict(A=A, P=None, agg=None, nagg=0, emax=0.0)]
    while len(levels) < max_levels:
        l = len(levels) - 1
        A = levels[l]["A"]
        m = A.shape[0]
        if l > 0 and m <= coarse_eq_limit:
            break
        adj, d = strength_graph(A.indptr, A.indices, A.data, threshold)
        agg, nagg = aggregate(adj, l < square_graph)
        if nagg == 0 or nagg >= m:
            break
        P0, Bc = tentative(agg, nagg, B)
        emax = 0.0
        if nsmooths == 1:
            emax = gershgorin_emax(A, d) if esteig == "gershgorin" else lanczos_emax(A, d, est_its)
            dinv = np.where(d != 0.0, 1.0 / np.where(d != 0.0, d, 1.0), 1.0)
            P = (P0 - (1.4 / emax) * (sp.diags(dinv) @ (A @ P0))).tocsr()
        else:
            P = P0
        P.sort_indices()
        Ac = (P.T @ (A @ P)).tocsr()
        Ac.sort_indices()
        levels[l].update(P=P, agg=agg, nagg=nagg, emax=emax)
        levels.append(dict(A=Ac, P=None, agg=None, nagg=0, emax=0.0))
        B = Bc
    return levels


# ---- the C V-cycle / PCG over a given hierarchy ------------------------------------------------
class _Packed:
    """Arrays of per-level pointers for orc_mg_apply / orc_cg_mg; keeps the numpy arrays alive."""

    def __init__(self, levels):
        self.nlev = len(levels)
        self.keep = []
        self.m = np.array([lv["A"][0].shape[0] - 1 if isinstance(lv["A"], tuple) else lv["A"].shape[0] for lv in levels], dtype=np.int32)

        def triple(M):
            if M is None:
                z = (np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.float64))
            elif isinstance(M, tuple):
                z = (_i32(M[0]), _i32(M[1]), _f64(M[2]))
            else:
                z = (_i32(M.indptr), _i32(M.indices), _f64(M.data))
            self.keep.append(z)
            return z

        def ptrs(arrs):
            return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])

        A = [triple(lv["A"]) for lv in levels]
        P = [triple(lv["P"]) for lv in levels]
        self.ai, self.aj, self.aa = ptrs([t[0] for t in A]), ptrs([t[1] for t in A]), ptrs([t[2] for t in A])
        self.pi, self.pj, self.pa = ptrs([t[0] for t in P]), ptrs([t[1] for t in P]), ptrs([t[2] for t in P])

    def args(self, sweeps):
        return (C.c_int(self.nlev), C.c_int(sweeps), _p(self.m), self.ai, self.aj, self.aa, self.pi, self.pj, self.pa)


def mg_apply(levels, r, sweeps=1):
    """z = M^{-1} r, one V-cycle.  `levels`: dicts with A and P as scipy CSR or (ai, aj, aa) tuples."""
    pk = _Packed(levels)
    r = _f64(r)
    z = np.zeros(len(r))
    lib().orc_mg_apply(*pk.args(sweeps), _p(r), _p(z))
    return z


def cg_mg(levels, b, rtol=1e-14, atol=1e-12, max_it=10000, sweeps=1):
    pk = _Packed(levels)
    b = _f64(b)
    x = np.zeros(len(b))
    rn = C.c_double(0.0)
    lib().orc_cg_mg.restype = C.c_int
    its = lib().orc_cg_mg(*pk.args(sweeps), _p(b), _p(x), C.c_double(rtol), C.c_double(atol), C.c_int(max_it), C.byref(rn))
    return x, its, rn.value


# ---- row-partitioned variant (petsc-openacc_b200/dgamg.py) -------------------------------------------
def uncoupled_hierarchy(A, base, threshold=0.0, nsmooths=1, coarse_eq_limit=50, max_levels=30, square_graph=1):
    """Restatement of dgamg.setup on the GLOBAL matrix `A` (scipy CSR, rows in rank order, partition
    `base`): aggregates inside each rank's diagonal block, prolongator smoothed with that block
    (P = blockdiag(P_r)), emax = global Gershgorin bound, Galerkin product on the global matrices.
    Returns (levels, bases): levels as in `hierarchy`, bases[l] = row partition of level l."""
    A = A.tocsr()
    A.sort_indices()
    levels = [dict(A=A, P=None, agg=None, nagg=0, emax=0.0)]
    B = np.ones(A.shape[0])
    bases = [np.asarray(base, dtype=np.int64)]
    while len(levels) < max_levels:
        l = len(levels) - 1
        A, base = levels[l]["A"], bases[l]
        m = A.shape[0]
        if l > 0 and m <= coarse_eq_limit:
            break
        d = A.diagonal()
        emax = gershgorin_emax(A, d)
        blocks, Bcs, aggs, naggs = [], [], [], []
        for r in range(len(base) - 1):
            lo, hi = int(base[r]), int(base[r + 1])
            Ad = A[lo:hi, lo:hi].tocsr()
            Ad.sort_indices()
            adj, dd = strength_graph(Ad.indptr, Ad.indices, Ad.data, threshold)
            agg, nagg = aggregate(adj, l < square_graph)
            P0, Bc = tentative(agg, nagg, B[lo:hi])
            if nsmooths == 1 and nagg:
                dinv = np.where(dd != 0.0, 1.0 / np.where(dd != 0.0, dd, 1.0), 1.0)
                P = (P0 - (1.4 / emax) * (sp.diags(dinv) @ (Ad @ P0))).tocsr()
            else:
                P = P0.tocsr()
            blocks.append(P)
            Bcs.append(Bc)
            aggs.append(agg)
            naggs.append(nagg)
        if min(naggs) == 0 or sum(naggs) >= m:
            break
        P = sp.block_diag(blocks, format="csr")
        P.sort_indices()
        Ac = (P.T @ (A @ P)).tocsr()
        Ac.sort_indices()
        levels[l].update(P=P, agg=aggs, nagg=sum(naggs), emax=emax)
        levels.append(dict(A=Ac, P=None, agg=None, nagg=0, emax=0.0))
        B = np.concatenate(Bcs)
        bases.append(np.concatenate([[0], np.cumsum(naggs)]).astype(np.int64))
    return levels, bases
