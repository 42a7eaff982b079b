"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

The CPU restatement of the reference's SeqAIJ hot path (see seqaij_oracle.c for the
reference file:line each function follows and for the "parity unpinned" statement).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "seqaij_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_matmult_flops.restype = C.c_double
        _lib.orc_poisson7_diag_scale.restype = C.c_double
        _lib.orc_poisson7_exact0.restype = C.c_double
        _lib.orc_vecdot.restype = C.c_double
        _lib.orc_vecnorm2.restype = C.c_double
        _lib.orc_vecnorm_inf.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def gen_vector(n, seed=0xB200):
    """x[i] = uniform [-1,1) from splitmix64(seed ^ i): the synthetic x of SURVEY 8(d), in numpy, so
    that the reference arm of bench.py needs nothing of the product package."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) ^ np.arange(n, dtype=np.uint64)) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0


# ---- oracle/_ref: the reference's own MatMult loops, compiled from its patch files -------------
_REF_SO = os.path.join(_HERE, "_ref", "libref_matmult.so")
_ref = None


def ref_lib():
    """CDLL of oracle/_ref/libref_matmult.so (ref_harness.c + the loop text extract_ref_loops.py cuts
    from the reference tree at build time), or None where it was never built."""
    global _ref
    if _ref is None and os.path.exists(_REF_SO):
        _ref = C.CDLL(_REF_SO)
    return _ref


def ref_matmult(ai, aj, aa, x, variant="original", host_rows=0):
    """y = A x by the reference's loop text.  variant "original": PETSc 3.7.6's loop as the step-1
    patch shows it; "step3": the author's host loop for the first `host_rows` rows, then the device
    loop (src/openacc-step3/MatMult_SeqAIJ.patch:36-70) compiled as plain C."""
    L = ref_lib()
    if L is None:
        raise RuntimeError("oracle/_ref/libref_matmult.so is not built (make -C oracle with the reference tree mounted)")
    ai, aj, aa, x = _i32(ai), _i32(aj), _f64(aa), _f64(x)
    m = len(ai) - 1
    y = np.empty(m, dtype=np.float64)
    if variant == "original":
        L.ref_matmult_original(C.c_int(m), _p(ai), _p(aj), _p(aa), _p(x), _p(y))
    elif variant in ("step3", "step4"):   # step4: the device loop in blocks of 983,040 rows
        f = L.ref_matmult_step3 if variant == "step3" else L.ref_matmult_step4
        f(C.c_int(m), _p(ai), _p(aj), _p(aa), _p(x), _p(y), C.c_int(host_rows))
    else:
        raise ValueError(variant)
    return y


def ref_matmult_mt(ai, aj, aa, x, nthreads, y=None):
    """The reference's loop on `nthreads` row blocks (one MPI rank per core in its runs)."""
    L = ref_lib()
    if L is None:
        raise RuntimeError("oracle/_ref/libref_matmult.so is not built")
    m = len(ai) - 1
    if y is None:
        y = np.empty(m, dtype=np.float64)
    L.ref_matmult_mt(C.c_int(nthreads), C.c_int(m), _p(ai), _p(aj), _p(aa), _p(x), _p(y))
    return y


# ---- SpMV family -----------------------------------------------------------------------------
def matmult(ai, aj, aa, x, fma=False):
    ai, aj, aa, x = _i32(ai), _i32(aj), _f64(aa), _f64(x)
    m = len(ai) - 1
    y = np.empty(m, dtype=np.float64)
    f = lib().orc_matmult_fma if fma else lib().orc_matmult
    f(C.c_int(m), _p(ai), _p(aj), _p(aa), _p(x), _p(y))
    return y


def matmult_mt(ai, aj, aa, x, nthreads, fma=False, y=None):
    m = len(ai) - 1
    if y is None:
        y = np.empty(m, dtype=np.float64)
    lib().orc_matmult_mt(C.c_int(nthreads), C.c_int(int(fma)), C.c_int(m), _p(ai), _p(aj), _p(aa),
                         _p(x), _p(y))
    return y


def matmult_cprow(m, cpi, ridx, aj, aa, x):
    cpi, ridx, aj, aa, x = _i32(cpi), _i32(ridx), _i32(aj), _f64(aa), _f64(x)
    y = np.empty(m, dtype=np.float64)
    lib().orc_matmult_cprow(C.c_int(m), C.c_int(len(ridx)), _p(cpi), _p(ridx), _p(aj), _p(aa),
                            _p(x), _p(y))
    return y


def matmultadd(ai, aj, aa, x, y, fma=False):
    ai, aj, aa, x, y = _i32(ai), _i32(aj), _f64(aa), _f64(x), _f64(y)
    m = len(ai) - 1
    z = np.empty(m, dtype=np.float64)
    f = lib().orc_matmultadd_fma if fma else lib().orc_matmultadd
    f(C.c_int(m), _p(ai), _p(aj), _p(aa), _p(x), _p(y), _p(z))
    return z


def matmultadd_cprow(m, cpi, ridx, aj, aa, x, y):
    cpi, ridx, aj, aa, x, y = _i32(cpi), _i32(ridx), _i32(aj), _f64(aa), _f64(x), _f64(y)
    z = np.empty(m, dtype=np.float64)
    lib().orc_matmultadd_cprow(C.c_int(m), C.c_int(len(ridx)), _p(cpi), _p(ridx), _p(aj), _p(aa),
                               _p(x), _p(y), _p(z))
    return z


def matmulttranspose(ai, aj, aa, x, n, fma=False):
    ai, aj, aa, x = _i32(ai), _i32(aj), _f64(aa), _f64(x)
    m = len(ai) - 1
    y = np.empty(n, dtype=np.float64)
    f = lib().orc_matmulttranspose_fma if fma else lib().orc_matmulttranspose
    f(C.c_int(m), C.c_int(n), _p(ai), _p(aj), _p(aa), _p(x), _p(y))
    return y


def matmulttransposeadd(ai, aj, aa, x, z, n):
    ai, aj, aa, x, z = _i32(ai), _i32(aj), _f64(aa), _f64(x), _f64(z)
    m = len(ai) - 1
    y = np.empty(n, dtype=np.float64)
    lib().orc_matmulttransposeadd(C.c_int(m), C.c_int(n), _p(ai), _p(aj), _p(aa), _p(x), _p(z),
                                  _p(y))
    return y


def residual(ai, aj, aa, x, b):
    ai, aj, aa, x, b = _i32(ai), _i32(aj), _f64(aa), _f64(x), _f64(b)
    m = len(ai) - 1
    r = np.empty(m, dtype=np.float64)
    lib().orc_residual(C.c_int(m), _p(ai), _p(aj), _p(aa), _p(x), _p(b), _p(r))
    return r


def jacobi_sweep(ai, aj, aa, x, b, dinv):
    ai, aj, aa, x, b, dinv = _i32(ai), _i32(aj), _f64(aa), _f64(x), _f64(b), _f64(dinv)
    m = len(ai) - 1
    xn = np.empty(m, dtype=np.float64)
    lib().orc_jacobi_sweep(C.c_int(m), _p(ai), _p(aj), _p(aa), _p(x), _p(b), _p(dinv), _p(xn))
    return xn


def row_abs_sum(ai, aj, aa, x):
    ai, aj, aa, x = _i32(ai), _i32(aj), _f64(aa), _f64(x)
    m = len(ai) - 1
    s = np.empty(m, dtype=np.float64)
    lib().orc_row_abs_sum(C.c_int(m), _p(ai), _p(aj), _p(aa), _p(x), _p(s))
    return s


def matmult_flops(nz, nonzerorowcnt):
    return lib().orc_matmult_flops(C.c_int(nz), C.c_int(nonzerorowcnt))


# ---- assembly --------------------------------------------------------------------------------
def assembly_end(ai, aj, aa, imax, ailen):
    """In-place compaction; returns (nz, nonzerorowcnt, rmax, fshift)."""
    m = len(ai) - 1
    nz, nzr, rmax = C.c_int(0), C.c_int(0), C.c_int(0)
    fshift = lib().orc_assembly_end(C.c_int(m), _p(ai), _p(aj), _p(aa), _p(imax), _p(ailen),
                                    C.byref(nz), C.byref(nzr), C.byref(rmax))
    return nz.value, nzr.value, rmax.value, fshift


def check_compressed_row(ai, nonzerorowcnt, ratio=0.6):
    ai = _i32(ai)
    m = len(ai) - 1
    cpi = np.zeros(nonzerorowcnt + 1, dtype=np.int32)
    ridx = np.zeros(max(nonzerorowcnt, 1), dtype=np.int32)
    nrows = C.c_int(0)
    use = lib().orc_check_compressed_row(C.c_int(m), _p(ai), C.c_int(nonzerorowcnt),
                                         C.c_double(ratio), _p(cpi), _p(ridx), C.byref(nrows))
    return bool(use), cpi[:nrows.value + 1].copy(), ridx[:nrows.value].copy()


# ---- generator -------------------------------------------------------------------------------
def dmda_decide(M, N, P, size):
    m, n, p = C.c_int(0), C.c_int(0), C.c_int(0)
    lib().orc_dmda_decide(C.c_int(M), C.c_int(N), C.c_int(P), C.c_int(size), C.byref(m),
                          C.byref(n), C.byref(p))
    return m.value, n.value, p.value


def dmda_info(M, N, P, size, rank):
    out = np.zeros(9, dtype=np.int32)
    lib().orc_dmda_info(C.c_int(M), C.c_int(N), C.c_int(P), C.c_int(size), C.c_int(rank), _p(out))
    return dict(zip("m n p xs ys zs xm ym zm".split(), (int(v) for v in out)))


def dmda_bases(M, N, P, size):
    base = np.zeros(size + 1, dtype=np.int32)
    lib().orc_dmda_bases(C.c_int(M), C.c_int(N), C.c_int(P), C.c_int(size), _p(base))
    return base


def poisson7(M, N=None, P=None, size=1, rank=0, refpoint=True):
    """The reference problem (src/helper.cpp) for one rank: rows with GLOBAL column ids.

    Returns dict(ai, aj, aa, rhs, exact, rstart, rend, scale)."""
    N = M if N is None else N
    P = M if P is None else P
    info = dmda_info(M, N, P, size, rank)
    nloc = info["xm"] * info["ym"] * info["zm"]
    ai = np.zeros(nloc + 1, dtype=np.int32)
    aj = np.zeros(7 * nloc, dtype=np.int32)
    aa = np.zeros(7 * nloc, dtype=np.float64)
    nz = lib().orc_poisson7_rows(C.c_int(M), C.c_int(N), C.c_int(P), C.c_int(size), C.c_int(rank),
                                 _p(ai), _p(aj), _p(aa))
    aj, aa = aj[:nz].copy(), aa[:nz].copy()
    rhs = np.zeros(nloc, dtype=np.float64)
    exact = np.zeros(nloc, dtype=np.float64)
    lib().orc_poisson7_vectors(C.c_int(M), C.c_int(N), C.c_int(P), C.c_int(size), C.c_int(rank),
                               _p(rhs), _p(exact))
    base = dmda_bases(M, N, P, size)
    scale = None
    if refpoint:
        scale = lib().orc_poisson7_diag_scale(C.c_int(M), C.c_int(N), C.c_int(P), C.c_int(size))
        e0 = lib().orc_poisson7_exact0(C.c_int(M), C.c_int(N), C.c_int(P))
        lib().orc_poisson7_refpoint(C.c_int(nloc), C.c_int(int(base[rank])), _p(ai), _p(aj),
                                    _p(aa), _p(rhs), C.c_double(e0), C.c_double(scale))
    return dict(ai=ai, aj=aj, aa=aa, rhs=rhs, exact=exact, rstart=int(base[rank]),
                rend=int(base[rank + 1]), scale=scale, info=info)


# ---- MPIAIJ ----------------------------------------------------------------------------------
def mpiaij_split(ai, aj, aa, cstart, cend):
    ai, aj, aa = _i32(ai), _i32(aj), _f64(aa)
    nloc, nz = len(ai) - 1, len(aj)
    Ai = np.zeros(nloc + 1, np.int32); Aj = np.zeros(nz, np.int32); Aa = np.zeros(nz)
    Bi = np.zeros(nloc + 1, np.int32); Bj = np.zeros(nz, np.int32); Ba = np.zeros(nz)
    bnz = C.c_int(0)
    anz = lib().orc_mpiaij_split(C.c_int(nloc), C.c_int(cstart), C.c_int(cend), _p(ai), _p(aj),
                                 _p(aa), _p(Ai), _p(Aj), _p(Aa), _p(Bi), _p(Bj), _p(Ba),
                                 C.byref(bnz))
    b = bnz.value
    return (Ai, Aj[:anz].copy(), Aa[:anz].copy()), (Bi, Bj[:b].copy(), Ba[:b].copy())


def mpiaij_setup_multiply(Bj_global):
    Bj = _i32(Bj_global).copy()
    garray = np.zeros(max(len(Bj), 1), dtype=np.int32)
    ng = lib().orc_mpiaij_setup_multiply(C.c_int(len(Bj)), _p(Bj), _p(garray))
    return Bj, garray[:ng].copy()


def scatter_recv_offsets(base, garray):
    base, garray = _i32(base), _i32(garray)
    size = len(base) - 1
    off = np.zeros(size + 1, dtype=np.int32)
    lib().orc_scatter_recv_offsets(C.c_int(size), _p(base), C.c_int(len(garray)), _p(garray),
                                   _p(off))
    return off


# ---- CG --------------------------------------------------------------------------------------
def vecdot(x, y):
    x, y = _f64(x), _f64(y)
    return lib().orc_vecdot(C.c_int(len(x)), _p(x), _p(y))


def cg_jacobi(ai, aj, aa, b, rtol=1e-14, atol=1e-12, max_it=10000):
    ai, aj, aa, b = _i32(ai), _i32(aj), _f64(aa), _f64(b)
    m = len(ai) - 1
    x = np.zeros(m)
    rn = C.c_double(0.0)
    lib().orc_cg_jacobi.restype = C.c_int
    its = lib().orc_cg_jacobi(C.c_int(m), _p(ai), _p(aj), _p(aa), _p(b), _p(x), C.c_double(rtol),
                              C.c_double(atol), C.c_int(max_it), C.byref(rn))
    return x, its, rn.value
