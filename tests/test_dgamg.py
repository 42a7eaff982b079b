"""Row-partitioned multigrid set-up (petsc-openacc_b200/dgamg.py, BASELINE configs[2]): every rank
builds its rows of every level; the assembled hierarchy must equal the global restatement
oracle/gamg.py::uncoupled_hierarchy.  All ranks run as threads of this process (ThreadComm) and, once,
as two gloo processes (TorchComm); the device solve with all ranks on one device is the gpu test."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from oracle import gamg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def global_problem(N, size):
    base = oracle.dmda_bases(N, N, N, size)
    parts = [oracle.poisson7(N, size=size, rank=r) for r in range(size)]
    A = sp.vstack([sp.csr_matrix((p["aa"], p["aj"], p["ai"]), shape=(len(p["ai"]) - 1, N ** 3)) for p in parts]).tocsr()
    return A, base, parts


def assemble(levels_per_rank, l):
    """Global CSR of level l from the ranks' row blocks, and the block-diagonal prolongator."""
    rows = [lv[l].rows for lv in levels_per_rank]
    n = int(levels_per_rank[0][l].base[-1])
    A = sp.vstack([sp.csr_matrix((aa, aj, ai), shape=(len(ai) - 1, n)) for ai, aj, aa in rows]).tocsr()
    P = None
    if levels_per_rank[0][l].P is not None:
        P = sp.block_diag([sp.csr_matrix((lv[l].P[2], lv[l].P[1], lv[l].P[0]), shape=(len(lv[l].P[0]) - 1, lv[l].P[3]))
                           for lv in levels_per_rank], format="csr")
    return A, P


@pytest.mark.parametrize("N,size", [(10, 1), (10, 2), (12, 4), (12, 8), (9, 3)])
def test_distributed_setup_equals_the_global_restatement(pk, N, size):
    from petsc_openacc_b200 import dgamg
    A, base, parts = global_problem(N, size)

    def rank_main(comm):
        p = parts[comm.rank]
        return dgamg.setup(comm, base, p["ai"], p["aj"], p["aa"])

    per_rank = dgamg.ThreadComm.run(size, rank_main)
    want, bases = gamg.uncoupled_hierarchy(A, base)
    nlev = len(per_rank[0])
    assert all(len(lv) == nlev for lv in per_rank) and nlev == len(want) >= 3
    if size == 1:   # one rank: the row-partitioned rule is the one-process set-up of pcgamg.cpp
        seq = gamg.hierarchy(parts[0]["ai"], parts[0]["aj"], parts[0]["aa"], esteig="gershgorin")
        assert len(seq) == len(want)
        for a, b in zip(seq, want):
            assert abs(a["A"] - b["A"]).max() <= 1e-12 * abs(a["A"]).max()
    for l in range(nlev):
        assert np.array_equal(per_rank[0][l].base, bases[l])
        Al, Pl = assemble(per_rank, l)
        Al.eliminate_zeros()
        W = want[l]["A"].copy()
        W.eliminate_zeros()
        assert Al.shape == W.shape
        assert abs(Al - W).max() <= 1e-12 * abs(W).max()
        assert np.array_equal(Al.indptr, W.indptr) and np.array_equal(Al.indices, W.indices)
        if want[l]["P"] is not None:
            for r in range(size):
                assert np.array_equal(per_rank[r][l].agg, want[l]["agg"][r])      # integer work: bit-exact
                assert per_rank[r][l].emax == pytest.approx(want[l]["emax"], rel=1e-14)
            Pl.eliminate_zeros()
            assert abs(Pl - want[l]["P"]).max() <= 1e-12 * abs(want[l]["P"]).max()
        else:
            assert Pl is None
        dinv = np.concatenate([lv[l].dinv for lv in per_rank])
        assert np.array_equal(dinv, 1.0 / Al.diagonal())
    # what it is for: CG + V-cycle on this hierarchy (C restatement) converges like the one-rank hierarchy
    rhs = np.concatenate([p["rhs"] for p in parts])
    x, its, rn = gamg.cg_mg(want, rhs)
    p1 = oracle.poisson7(N)
    _, its1, _ = gamg.cg_mg(gamg.hierarchy(p1["ai"], p1["aj"], p1["aa"]), p1["rhs"])
    assert 0 < its <= its1 + 14
    for lv in per_rank:
        for L in lv:
            L.M.destroy()


WORKER = r'''
import os, sys, pickle
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np
import torch.distributed as dist
import oracle
from petsc_openacc_b200 import dgamg
dist.init_process_group("gloo")
comm = dgamg.TorchComm()
N = 10
base = oracle.dmda_bases(N, N, N, comm.size)
p = oracle.poisson7(N, size=comm.size, rank=comm.rank)
lv = dgamg.setup(comm, base, p["ai"], p["aj"], p["aa"])
out = [dict(base=L.base, rows=L.rows, P=L.P, agg=L.agg, emax=L.emax) for L in lv]
with open(os.path.join(sys.argv[1], f"rank{comm.rank}.pkl"), "wb") as f:
    pickle.dump(out, f)
dist.barrier()
dist.destroy_process_group()
'''


def test_distributed_setup_over_two_gloo_processes(pk, tmp_path):
    import pickle
    from petsc_openacc_b200 import dgamg
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29671", str(script), str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    got = [pickle.load(open(tmp_path / f"rank{r}.pkl", "rb")) for r in range(2)]
    N = 10
    A, base, parts = global_problem(N, 2)

    def rank_main(comm):
        p = parts[comm.rank]
        return dgamg.setup(comm, base, p["ai"], p["aj"], p["aa"])

    same = dgamg.ThreadComm.run(2, rank_main)                     # the threads and the processes agree bit for bit
    for r in range(2):
        assert len(got[r]) == len(same[r])
        for g, L in zip(got[r], same[r]):
            assert np.array_equal(g["base"], L.base)
            for k in range(3):
                assert np.array_equal(g["rows"][k], L.rows[k])
            if L.P is not None:
                for k in range(3):
                    assert np.array_equal(g["P"][k], L.P[k])
                assert np.array_equal(g["agg"], L.agg) and g["emax"] == L.emax
    for lv in same:
        for L in lv:
            L.M.destroy()


class OracleOps:
    """numpy/oracle stand-in for dgamg.DeviceOps: the same methods on numpy vectors, MatMult_MPIAIJ as
    PETSc orders it (diagonal block, then MatMultAdd with the ghost values), so that the orchestration
    of dgamg.Solver -- buffers, order of operations, the CG recurrence -- is checked on the CPU."""

    def __init__(self, comm, levels):
        self.comm = comm
        for L in levels:
            L.blkA, L.blkB = L.M.block(0), L.M.block(1)
            L.garr, L.roff = L.M.garray(), L.M.recv_offsets()

    def zeros(self, n):
        return np.zeros(n)

    def from_numpy(self, a):
        return np.array(a, dtype=np.float64)

    def to_numpy(self, t):
        return t.copy()

    def mult(self, L, x, y):
        allx = self.comm.allgather(x.copy())                    # the halo: ghost values from their owners
        lvec = np.zeros(max(len(L.garr), 1))
        for q in range(self.comm.size):
            g = L.garr[L.roff[q]:L.roff[q + 1]]
            lvec[L.roff[q]:L.roff[q + 1]] = allx[q][g - L.base[q]]
        Ai, Aj, Aa = L.blkA
        Bi, Bj, Ba = L.blkB
        t = oracle.matmult(Ai, Aj, Aa, x) if len(x) else np.zeros(0)
        y[:] = oracle.matmultadd(Bi, Bj, Ba, lvec, t) if len(x) else t

    def restrict(self, L, r, bc):
        Pi, Pj, Pa, nc = L.P
        bc[:] = oracle.matmulttranspose(Pi, Pj, Pa, r, nc)

    def interp_add(self, L, xc, x):
        Pi, Pj, Pa, nc = L.P
        x[:] = oracle.matmultadd(Pi, Pj, Pa, xc, x)

    def pointwise_mult(self, w, x, y):
        w[:] = x * y

    def aypx(self, y, a, x):
        y[:] = x + a * y

    def axpy(self, y, a, x):
        y[:] = y + a * x

    def copy(self, y, x):
        y[:] = x

    def set(self, x, a):
        x[:] = a

    def dots(self, pairs):
        parts = self.comm.allgather(np.array([oracle.vecdot(a, b) for a, b in pairs]))
        tot = np.zeros(len(pairs))
        for q in range(self.comm.size):
            tot = tot + parts[q]
        return [float(v) for v in tot]

    def destroy(self, levels):
        pass


@pytest.mark.parametrize("N,size,sweeps", [(10, 2, 1), (12, 4, 1), (12, 8, 2), (9, 3, 3)])
def test_solver_orchestration_with_the_oracle_backend(pk, N, size, sweeps):
    """dgamg.Solver (V-cycle + CG over the ranks) with the CPU stand-in for the device calls against
    the C restatement on the assembled global hierarchy: same preconditioner, same iteration count."""
    from petsc_openacc_b200 import dgamg
    A, base, parts = global_problem(N, size)
    want, _ = gamg.uncoupled_hierarchy(A, base)
    rhs = np.concatenate([p["rhs"] for p in parts])
    r0 = np.concatenate([np.sin(np.arange(len(p["rhs"])) + 3.0 * k) for k, p in enumerate(parts)])
    z_want = gamg.mg_apply(want, r0, sweeps=sweeps)
    xo, its_o, _ = gamg.cg_mg(want, rhs, sweeps=sweeps)

    def rank_main(comm):
        p = parts[comm.rank]
        lv = dgamg.setup(comm, base, p["ai"], p["aj"], p["aa"])
        sv = dgamg.Solver(comm, lv, sweeps=sweeps, ops=OracleOps(comm, lv))
        lo, hi = int(base[comm.rank]), int(base[comm.rank + 1])
        z = np.full(hi - lo, np.nan)
        sv.apply(r0[lo:hi].copy(), z)
        x = np.full(hi - lo, np.nan)
        its, reason, rn = sv.solve(p["rhs"].copy(), x)
        sv.destroy()
        return z, its, reason, x

    res = dgamg.ThreadComm.run(size, rank_main)
    z = np.concatenate([r[0] for r in res])
    # MatMult_MPIAIJ adds the ghost terms after the local ones: not the global CSR order, hence a tolerance
    assert np.abs(z - z_want).max() <= 1e-12 * np.abs(z_want).max()
    assert all(r[2] > 0 for r in res) and len({r[1] for r in res}) == 1
    assert abs(res[0][1] - its_o) <= 1
    x = np.concatenate([r[3] for r in res])
    assert np.abs(x - xo).max() <= 1e-9 * np.abs(xo).max()


@pytest.mark.gpu
@pytest.mark.parametrize("N,size", [(12, 2), (16, 4)])
def test_device_solve_in_process_ranks(pk, cuda, N, size):
    """All ranks on one device (threads, split-phase halo): the CG + V-cycle solve against the C
    restatement on the assembled global hierarchy."""
    torch = cuda
    from petsc_openacc_b200 import dgamg
    A, base, parts = global_problem(N, size)
    want, _ = gamg.uncoupled_hierarchy(A, base)
    rhs = np.concatenate([p["rhs"] for p in parts])
    xo, its_o, _ = gamg.cg_mg(want, rhs)

    def rank_main(comm):
        p = parts[comm.rank]
        lv = dgamg.setup(comm, base, p["ai"], p["aj"], p["aa"])
        sv = dgamg.Solver(comm, lv)
        b = torch.from_numpy(p["rhs"]).cuda()
        x = torch.zeros_like(b)
        its, reason, rn = sv.solve(b, x)
        out = (its, reason, x.cpu().numpy())
        sv.destroy()
        return out

    res = dgamg.ThreadComm.run(size, rank_main)
    assert all(r[1] > 0 for r in res) and len({r[0] for r in res}) == 1
    assert abs(res[0][0] - its_o) <= 1
    x = np.concatenate([r[2] for r in res])
    assert np.abs(x - xo).max() <= 1e-9 * np.abs(xo).max()


# ---- the host CSR kernels behind include/b200_gamg.h ---------------------------------------------------
def _rand(m, n, density, rng):
    M = sp.random(m, n, density=density, random_state=np.random.RandomState(int(rng.integers(1 << 30))), format="csr")
    M.sort_indices()
    return M


def test_host_csr_kernels_against_scipy(pk):
    from petsc_openacc_b200 import dgamg
    rng = np.random.default_rng(12)
    for (m, k, n) in [(1, 1, 1), (40, 30, 50), (300, 300, 300), (25000, 900, 700)]:   # the last one runs threaded
        X, Y, Z = _rand(m, k, min(1.0, 6.0 / k), rng), _rand(k, n, min(1.0, 5.0 / n), rng), _rand(m, k, min(1.0, 4.0 / k), rng)
        hX = dgamg.HostCsr.from_arrays(m, k, X.indptr, X.indices, X.data)
        hY = dgamg.HostCsr.from_arrays(k, n, Y.indptr, Y.indices, Y.data)
        hZ = dgamg.HostCsr.from_arrays(m, k, Z.indptr, Z.indices, Z.data)

        def back(h, shape):
            ai, aj, aa = h.arrays()
            assert all(np.all(np.diff(aj[ai[r]:ai[r + 1]]) > 0) for r in range(min(shape[0], 500)))   # ascending, unique
            return sp.csr_matrix((aa, aj, ai), shape=shape)

        P = back(hX.matmul(hY), (m, n))
        W = (X @ Y).tocsr()
        assert abs(P - W).max() <= 1e-13 * max(abs(W).max(), 1e-300) if W.nnz else P.nnz == 0 or abs(P).max() == 0
        T = back(hX.transpose(), (k, m))
        assert (T != X.T.tocsr()).nnz == 0
        S = back(hX.add(hZ), (m, k))
        assert abs(S - (X + Z)).max() <= 1e-15 * max(abs(X + Z).max(), 1e-300) if (X + Z).nnz else True
        assert np.allclose(hX.abs_row_sums(), np.asarray(abs(X).sum(axis=1)).ravel(), rtol=1e-14, atol=0)
        for h in (hX, hY, hZ):
            h.destroy()


def test_host_csr_rejects_bad_input(pk):
    from petsc_openacc_b200 import dgamg
    with pytest.raises(RuntimeError):
        dgamg.HostCsr.from_arrays(2, 3, [0, 2, 3], [1, 0, 2], [1.0, 2.0, 3.0])      # columns not ascending
    with pytest.raises(RuntimeError):
        dgamg.HostCsr.from_arrays(1, 2, [0, 1], [5], [1.0])                          # column out of range
    a = dgamg.HostCsr.from_arrays(2, 3, [0, 1, 2], [0, 2], [1.0, 2.0])
    b = dgamg.HostCsr.from_arrays(2, 2, [0, 1, 2], [0, 1], [1.0, 2.0])
    with pytest.raises(RuntimeError):
        a.matmul(b)                                                                  # 3 columns x 2 rows
    with pytest.raises(RuntimeError):
        a.add(b)
    a.destroy(); b.destroy()
