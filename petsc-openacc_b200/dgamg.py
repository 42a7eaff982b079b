"""dgamg.py -- CG + multigrid on the row-partitioned matrix, one process per GPU (BASELINE
configs[2]: "MatMult_MPIAIJ row-partitioned at 2/4/8 B200 ..., full CG+GAMG solve").

The reference gets this from PETSc's PCGAMG on a MATMPIAIJ matrix (src/helper.cpp:31,39 on more
than one rank; options configs/PETSc_SolverOptions_GAMG.info); PETSc is not available here, so
this is this repository's own row-partitioned variant of host/src/pcgamg.cpp -- PARITY UNPINNED:

  set-up (host, per rank, the C kernels of include/b200_gamg.h; the ranks meet in four small
  all-gathers per level):
    * aggregates are formed inside each rank's diagonal block ("uncoupled" aggregation) and the
      prolongator is smoothed with that block, P_r = (I - 1.4/emax D^-1 A_rr) P0_r, so P is block
      diagonal: restriction and interpolation need no communication;
    * emax = the global maximum of sum_j |a_ij| / |a_ii| over complete rows (Gershgorin);
    * Galerkin operator, row block r: P_r^T (A_rr P_r + A_r,ghost P_ghost) with the ghost rows of the
      neighbours' prolongators fetched once per level (the rows on each rank's send list);
    * every level is again a row-partitioned matrix (MpiAij) with its own garray / halo.
  solve (device): the V-cycle of pcgamg.cpp with MatMult_MPIAIJ (NVLink halo) in the residual and
  the Jacobi sweeps, MatMultTranspose / MatMultAdd on the local prolongator blocks, and KSPCG whose
  dot products are summed over the ranks in rank order (b200_mpiaij_allreduce_sum).

Process plumbing is a `Comm` with allgather/barrier: `TorchComm` (torch.distributed) for real runs,
`ThreadComm` (all ranks as threads of one process) for the CPU tests of the set-up and for the
single-GPU emulation.  STATUS: the set-up is tested on the CPU against oracle/gamg.py; the device
solve is tested with all ranks on one GPU (tests/test_dgamg.py) and ran on 2 and 8 B200s
(scripts/dgamg_worker.py; 300^3 on 8 GPUs: 109 iterations, 0.121 s, profiles/r02_dgamg_300_n8.log).
"""
import ctypes as C
import os
import threading

import numpy as np

from . import MODE_EXACT, MpiAij, Csr, check, lib as _aij  # noqa: F401  (libb200aij first: libb200petsc links to it)

_HERE = os.path.dirname(os.path.abspath(__file__))
_host = None


def host_lib():
    global _host
    if _host is None:
        _host = C.CDLL(os.path.join(_HERE, "libb200petsc.so"))
    return _host


def _chk(rc, what):
    if rc:
        raise RuntimeError(f"{what}: PetscErrorCode {rc}")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class HostCsr:
    """b200_hcsr_t: host CSR with the set-up's sparse kernels (include/b200_gamg.h)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_arrays(cls, m, n, ai, aj, aa):
        ai = np.ascontiguousarray(ai, dtype=np.int32)
        aj = np.ascontiguousarray(aj, dtype=np.int32)
        aa = np.ascontiguousarray(aa, dtype=np.float64)
        h = C.c_void_p(0)
        _chk(host_lib().b200_hcsr_create(C.byref(h), C.c_int32(m), C.c_int32(n), _ptr(ai), _ptr(aj), _ptr(aa)), "b200_hcsr_create")
        return cls(h)

    def shape(self):
        m, n, nz = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        _chk(host_lib().b200_hcsr_shape(self._h, C.byref(m), C.byref(n), C.byref(nz)), "b200_hcsr_shape")
        return m.value, n.value, nz.value

    def arrays(self):
        m, n, nz = self.shape()
        pi, pj, pa = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_double)()
        _chk(host_lib().b200_hcsr_arrays(self._h, C.byref(pi), C.byref(pj), C.byref(pa)), "b200_hcsr_arrays")
        ai = np.ctypeslib.as_array(pi, shape=(m + 1,)).copy()
        aj = np.ctypeslib.as_array(pj, shape=(max(nz, 1),))[:nz].copy()
        aa = np.ctypeslib.as_array(pa, shape=(max(nz, 1),))[:nz].copy()
        return ai, aj, aa

    def _binary(self, fn, other):
        out = C.c_void_p(0)
        _chk(getattr(host_lib(), fn)(self._h, other._h, C.byref(out)), fn)
        return HostCsr(out)

    def matmul(self, other):
        return self._binary("b200_hcsr_spgemm", other)

    def add(self, other):
        return self._binary("b200_hcsr_add", other)

    def transpose(self):
        out = C.c_void_p(0)
        _chk(host_lib().b200_hcsr_transpose(self._h, C.byref(out)), "b200_hcsr_transpose")
        return HostCsr(out)

    def abs_row_sums(self):
        out = np.zeros(max(self.shape()[0], 1))
        _chk(host_lib().b200_hcsr_abs_row_sums(self._h, _ptr(out)), "b200_hcsr_abs_row_sums")
        return out[:self.shape()[0]]

    def coarsen(self, B, threshold, square, emax):
        m = self.shape()[0]
        B = np.ascontiguousarray(B, dtype=np.float64)
        agg = np.zeros(max(m, 1), np.int32)
        Bc = np.zeros(max(m, 1))
        nagg, P = C.c_int32(0), C.c_void_p(0)
        _chk(host_lib().b200_gamg_coarsen_block(self._h, _ptr(B), C.c_double(threshold), C.c_int(int(square)), C.c_double(emax),
                                                _ptr(agg), C.byref(nagg), C.byref(P), _ptr(Bc)), "b200_gamg_coarsen_block")
        return agg[:m], nagg.value, HostCsr(P), Bc[:nagg.value].copy()

    def destroy(self):
        if self._h:
            host_lib().b200_hcsr_destroy(self._h)
            self._h = None


# ---- process plumbing ------------------------------------------------------------------------------
class TorchComm:
    """torch.distributed (nccl or gloo): one process per rank."""
    inprocess = False

    def __init__(self):
        import torch.distributed as dist
        self._dist = dist
        self.size, self.rank = dist.get_world_size(), dist.get_rank()

    def allgather(self, obj):
        out = [None] * self.size
        self._dist.all_gather_object(out, obj)
        return out

    def barrier(self):
        self._dist.barrier()


class ThreadComm:
    """All ranks as threads of one process: `ThreadComm.run(size, fn)` calls fn(comm) on every rank."""
    inprocess = True

    def __init__(self, size, rank, shared):
        self.size, self.rank, self._s = size, rank, shared

    def allgather(self, obj):
        s = self._s
        s["slots"][self.rank] = obj
        s["barrier"].wait()
        out = list(s["slots"])
        s["barrier"].wait()
        return out

    def barrier(self):
        self._s["barrier"].wait()

    @staticmethod
    def run(size, fn):
        shared = {"slots": [None] * size, "barrier": threading.Barrier(size)}
        results, errors = [None] * size, [None] * size

        def body(r):
            try:
                results[r] = fn(ThreadComm(size, r, shared))
            except BaseException as e:  # noqa: BLE001  (re-raised below; the others must not wait forever)
                errors[r] = e
                shared["barrier"].abort()

        th = [threading.Thread(target=body, args=(r,)) for r in range(size)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        real = [e for e in errors if e is not None and not isinstance(e, threading.BrokenBarrierError)]
        if real or any(errors):
            raise (real or [e for e in errors if e is not None])[0]
        return results


# ---- the hierarchy -----------------------------------------------------------------------------------
class Level:
    def __init__(self):
        self.M = None          # MpiAij of this level (rows of this rank, global columns)
        self.base = None       # row partition of this level
        self.P = None          # (ai, aj, aa, ncoarse_local): prolongator block of this rank, local coarse columns
        self.agg = None
        self.nagg = 0
        self.emax = 0.0
        self.dinv = None       # host reciprocal diagonal (PCJACOBI) of this rank's rows
        self.rows = None       # (ai, aj_global, aa) of this rank's rows, kept for tests / the oracle


def wire(comm, M):
    """VecScatter set-up of one level: every rank learns what the others need from it."""
    garrays = comm.allgather(M.garray())
    for q in range(comm.size):
        M.set_peer_garray(q, garrays[q])


def setup(comm, base, ai, aj_global, aa, threshold=0.0, nsmooths=1, coarse_eq_limit=50, max_levels=30, square_graph=1):
    """Builds the row-partitioned hierarchy from this rank's rows (global column ids).  Returns the
    list of `Level`s, finest first.  Collective over `comm`."""
    size, rank = comm.size, comm.rank
    levels = []
    base = np.ascontiguousarray(base, dtype=np.int32)
    ai = np.ascontiguousarray(ai, dtype=np.int32)
    aj = np.ascontiguousarray(aj_global, dtype=np.int32)
    aa = np.ascontiguousarray(aa, dtype=np.float64)
    Bvec = np.ones(len(ai) - 1)
    while True:
        L = Level()
        L.base, L.rows = base, (ai, aj, aa)
        L.M = MpiAij(size, rank, base, ai, aj, aa)
        wire(comm, L.M)
        nloc = L.M.nloc
        Ai, Aj, Aa = L.M.block(0)
        Bi, Bj, Ba = L.M.block(1)
        Ad = HostCsr.from_arrays(nloc, nloc, Ai, Aj, Aa)
        d = np.zeros(nloc)                      # diagonal of this rank's rows (it lies in the diagonal block)
        rows_of = np.repeat(np.arange(nloc, dtype=np.int32), np.diff(Ai))
        on_diag = Aj == rows_of
        d[rows_of[on_diag]] = Aa[on_diag]
        L.dinv = np.where(d != 0.0, 1.0 / np.where(d != 0.0, d, 1.0), 1.0)
        levels.append(L)
        nglobal = int(base[-1])
        if len(levels) >= max_levels or (len(levels) > 1 and nglobal <= coarse_eq_limit):
            Ad.destroy()
            break
        # global Gershgorin bound of lambda_max(D^-1 A) over complete rows (diagonal + off-diagonal block)
        ng = max(L.M.nghost, 1)                 # (an empty ghost list is kept as one unused column)
        Bo = HostCsr.from_arrays(nloc, ng, Bi, Bj, Ba)
        rowsum = Ad.abs_row_sums() + Bo.abs_row_sums()
        ok = d != 0.0
        local_emax = float(np.max(rowsum[ok] / np.abs(d[ok]))) if ok.any() else 0.0
        emax = max(comm.allgather(local_emax))
        agg, nagg, P, Bc = Ad.coarsen(Bvec, threshold, len(levels) - 1 < square_graph, emax if nsmooths else 0.0)
        naggs = comm.allgather(nagg)
        cbase = np.concatenate([[0], np.cumsum(naggs)]).astype(np.int32)
        ncoarse = int(cbase[-1])
        if min(naggs) == 0 or ncoarse >= nglobal:   # a rank with nothing to aggregate: stop here
            for h in (Ad, Bo, P):
                h.destroy()
            break
        Pi, Pj, Pa = P.arrays()
        L.P, L.agg, L.nagg, L.emax = (Pi, Pj, Pa, nagg), agg, nagg, emax
        # ghost rows of the neighbours' prolongators, with GLOBAL coarse columns
        outgoing = {}
        for q in range(size):
            if q == rank:
                continue
            idx, _ = L.M.send_list(q)
            if len(idx):
                lens = (Pi[idx + 1] - Pi[idx]).astype(np.int32)
                take = np.concatenate([np.arange(Pi[i], Pi[i + 1]) for i in idx] or [np.zeros(0, np.int64)]).astype(np.int64)
                outgoing[q] = (lens, (Pj[take] + cbase[rank]).astype(np.int32), Pa[take])
        inbox = comm.allgather(outgoing)
        garray, roff = L.M.garray(), L.M.recv_offsets()
        glen = np.zeros(len(garray), np.int32)
        gcols, gvals = [None] * size, [None] * size
        for q in range(size):
            part = inbox[q].get(rank) if q != rank else None
            n_from_q = int(roff[q + 1] - roff[q])
            if n_from_q:
                if part is None or len(part[0]) != n_from_q:
                    raise RuntimeError(f"rank {rank}: expected {n_from_q} prolongator rows from rank {q}")
                glen[roff[q]:roff[q + 1]] = part[0]
                gcols[q], gvals[q] = part[1], part[2]
        Gi = np.zeros(ng + 1, np.int32)
        Gi[1:len(glen) + 1] = np.cumsum(glen)
        Gi[len(glen) + 1:] = Gi[len(glen)]
        Gj = np.concatenate([c for c in gcols if c is not None] or [np.zeros(0, np.int32)])
        Ga = np.concatenate([v for v in gvals if v is not None] or [np.zeros(0)])
        Pghost = HostCsr.from_arrays(ng, ncoarse, Gi, Gj, Ga)
        Pglob = HostCsr.from_arrays(nloc, ncoarse, Pi, (Pj + cbase[rank]).astype(np.int32), Pa)
        AP = Ad.matmul(Pglob).add(Bo.matmul(Pghost))          # rows of this rank, global coarse columns
        PT = P.transpose()
        Ac = PT.matmul(AP)                                     # row block of the Galerkin operator
        ci, cj, ca = Ac.arrays()
        for h in (Ad, Bo, P, Pghost, Pglob, AP, PT, Ac):
            h.destroy()
        base, ai, aj, aa, Bvec = cbase, ci, cj, ca, Bc
    return levels


# ---- device solve ------------------------------------------------------------------------------------
class DeviceOps:
    """Everything the solve does on the GPU, one method per PETSc call.  Vectors are torch CUDA
    tensors of this rank's rows.  (tests/test_dgamg.py substitutes a numpy/oracle stand-in with the
    same methods to check the orchestration of `Solver` on the CPU.)"""

    def __init__(self, comm, levels, mode=MODE_EXACT):
        import torch
        self.torch, self.comm, self.mode = torch, comm, mode
        self.dev = torch.device("cuda", torch.cuda.current_device())
        for L in levels:
            L.M.upload()
        if comm.inprocess:
            ptrs = [comm.allgather(L.M.window_ptr()) for L in levels]
            for L, pl in zip(levels, ptrs):
                for q in range(comm.size):
                    if q != comm.rank:
                        if len(L.M.send_list(q)[0]):
                            L.M.set_peer_window(q, pl[q])
                        L.M.set_rank_window(q, ptr=pl[q])
        else:
            handles = [comm.allgather(L.M.ipc_handle()) for L in levels]
            for L, hl in zip(levels, handles):
                for q in range(comm.size):
                    if q != comm.rank and len(L.M.send_list(q)[0]):
                        L.M.open_peer_window(q, hl[q])
                for q in range(comm.size):
                    L.M.set_rank_window(q, handle=hl[q] if q != comm.rank else None)
        for L in levels:
            L.Pdev = None
            if L.P is not None:
                Pi, Pj, Pa, nc = L.P
                L.Pdev = Csr(Pi, Pj, Pa, n=max(nc, 1))
                L.Pdev.build_transpose()
        self.allreduce_level = levels[0]
        self.scal = torch.zeros(3, dtype=torch.float64, device=self.dev)

    def zeros(self, n):
        return self.torch.zeros(max(n, 1), dtype=self.torch.float64, device=self.dev)[:n]

    def from_numpy(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.dev)

    def to_numpy(self, t):
        return t.cpu().numpy()

    # y = A x on level L (MatMult_MPIAIJ); split-phase when all ranks share one device and one stream
    def mult(self, L, x, y):
        if self.comm.inprocess:
            L.M.mult_begin(x)
            self.comm.barrier()
            L.M.mult_finish(x, y, self.mode)
            self.comm.barrier()
        else:
            L.M.mult(x, y, self.mode)

    def restrict(self, L, r, bc):          # MatRestrict: b_c = P^T r, this rank's block
        if L.M.nloc:
            L.Pdev.mult_transpose(r, bc, self.mode)

    def interp_add(self, L, xc, x):        # MatInterpolateAdd: x = x + P x_c, this rank's block
        if L.M.nloc and len(xc):
            L.Pdev.mult_add(xc, x, x, self.mode)

    def pointwise_mult(self, w, x, y):
        from . import vec_pointwise_mult
        vec_pointwise_mult(w, x, y)

    def aypx(self, y, a, x):
        from . import vec_aypx
        vec_aypx(y, a, x)

    def axpy(self, y, a, x):
        from . import vec_axpy
        vec_axpy(y, a, x)

    def copy(self, y, x):
        from . import vec_copy
        vec_copy(y, x)

    def set(self, x, a):
        from . import vec_set
        vec_set(x, a)

    def dots(self, pairs):
        """Global dot products, partial sums added in rank order (same bits on every rank)."""
        from . import vec_dot
        for i, (a, b) in enumerate(pairs):
            vec_dot(a, b, self.scal[i:i + 1])
        if self.comm.inprocess:
            # one device, one stream: the peer-window all-reduce would wait on kernels queued behind it
            parts = self.comm.allgather(self.scal[:len(pairs)].cpu().numpy().copy())
            tot = np.zeros(len(pairs))
            for q in range(self.comm.size):
                tot = tot + parts[q]
            return [float(v) for v in tot]
        self.allreduce_level.M.allreduce_sum(self.scal[:len(pairs)])
        return [float(v) for v in self.scal[:len(pairs)].cpu()]

    def destroy(self, levels):
        for L in levels:
            if getattr(L, "Pdev", None) is not None:
                L.Pdev.destroy()
                L.Pdev = None


class Solver:
    """KSPCG preconditioned by the V-cycle of a row-partitioned hierarchy.  Collective over `comm`."""

    def __init__(self, comm, levels, sweeps=1, mode=MODE_EXACT, ops=None):
        self.comm, self.levels, self.sweeps = comm, levels, int(sweeps)
        self.ops = ops if ops is not None else DeviceOps(comm, levels, mode)
        for L in levels:
            n = L.M.nloc
            L.d_dinv = self.ops.from_numpy(L.dinv)
            L.d_b, L.d_x, L.d_t, L.d_r = (self.ops.zeros(n) for _ in range(4))

    def _residual(self, L, b, x, r):
        self.ops.mult(L, x, r)
        self.ops.aypx(r, -1.0, b)                       # r = b - A x        (VecAYPX(r, -1, b))

    def _sweep(self, L, b, x):
        self._residual(L, b, x, L.d_r)
        self.ops.pointwise_mult(L.d_r, L.d_dinv, L.d_r)  # z = dinv .* r
        self.ops.axpy(x, 1.0, L.d_r)                     # x = x + z

    def _cycle(self, l, b, x):
        L = self.levels[l]
        self.ops.pointwise_mult(x, b, L.d_dinv)          # coarse solve, or the first sweep from the zero guess
        if l + 1 == len(self.levels):
            return
        for _ in range(self.sweeps - 1):
            self._sweep(L, b, x)
        self._residual(L, b, x, L.d_t)
        C_ = self.levels[l + 1]
        self.ops.restrict(L, L.d_t, C_.d_b)
        self._cycle(l + 1, C_.d_b, C_.d_x)
        self.ops.interp_add(L, C_.d_x, x)
        for _ in range(self.sweeps):
            self._sweep(L, b, x)

    def apply(self, r, z):
        """z = M^-1 r: one V-cycle."""
        self._cycle(0, r, z)

    def solve(self, b, x, rtol=1e-14, atol=1e-12, max_it=10000):
        """KSPSolve_CG [P376] with the V-cycle; b, x = this rank's rows.  Returns (its, reason, rnorm)."""
        ops, L0 = self.ops, self.levels[0]
        r, z, p, w = (ops.zeros(L0.M.nloc) for _ in range(4))
        ops.set(x, 0.0)
        ops.copy(r, b)
        self.apply(r, z)
        zz, beta = ops.dots([(z, z), (z, r)])
        dp = rnorm0 = zz ** 0.5
        ttol = max(rtol * rnorm0, atol)
        it, betaold, reason = 0, 1.0, 0
        if dp < ttol:
            return 0, (3 if dp < atol else 2), dp
        while it < max_it:
            if beta == 0.0:
                reason = 3
                break
            if it == 0:
                ops.copy(p, z)
            else:
                ops.aypx(p, beta / betaold, z)
            betaold = beta
            ops.mult(L0, p, w)
            (dpi,) = ops.dots([(p, w)])
            a = beta / dpi
            ops.axpy(x, a, p)
            ops.axpy(r, -a, w)
            self.apply(r, z)
            zz, beta = ops.dots([(z, z), (z, r)])
            dp = zz ** 0.5
            it += 1
            if dp < ttol:
                reason = 3 if dp < atol else 2
                break
        if reason == 0:
            reason = -3
        return it, reason, dp

    def destroy(self):
        self.ops.destroy(self.levels)
        for L in self.levels:
            L.M.destroy()
