"""Product-side host code against the oracle: the benchmark generator (b200_gen_poisson7) must
give bit-identical CSR / rhs / exact for every rank of every process grid."""
import ctypes as C

import numpy as np
import pytest

import gen
import oracle


def product_poisson(pk, M, size, rank, refpoint=True):
    out = np.zeros(12, np.int32)
    pk.check(pk.lib.b200_gen_poisson7_info(M, M, M, size, rank, out.ctypes.data_as(C.c_void_p)))
    nloc, nnz = int(out[9]), int(out[11])
    ai, aj, aa = np.zeros(nloc + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
    rhs, ex = np.zeros(nloc), np.zeros(nloc)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    pk.check(pk.lib.b200_gen_poisson7(M, M, M, size, rank, int(refpoint), p(ai), p(aj), p(aa), p(rhs), p(ex)))
    return dict(ai=ai, aj=aj, aa=aa, rhs=rhs, exact=ex, info=out)


@pytest.mark.parametrize("M,size", [(5, 1), (50, 1), (10, 2), (12, 4), (13, 8), (9, 6), (20, 3), (11, 7)])
def test_poisson_generator_bit_exact(pk, M, size):
    bases = np.zeros(size + 1, np.int32)
    pk.check(pk.lib.b200_gen_poisson7_bases(M, M, M, size, bases.ctypes.data_as(C.c_void_p)))
    assert np.array_equal(bases, oracle.dmda_bases(M, M, M, size))
    for r in range(size):
        a = product_poisson(pk, M, size, r)
        o = oracle.poisson7(M, size=size, rank=r)
        assert int(a["info"][11]) == len(o["aj"])
        for k in ("ai", "aj", "aa", "rhs", "exact"):
            assert np.array_equal(a[k], o[k]), (k, r)
        inf = oracle.dmda_info(M, M, M, size, r)
        assert [int(v) for v in a["info"][:9]] == [inf[k] for k in "m n p xs ys zs xm ym zm".split()]


def test_gen_vector_matches_definition(pk):
    assert np.array_equal(pk.gen_vector(1000, 0xB200), gen.uniform_pm1(1000, 0xB200))
    x = pk.gen_vector(100000, 5)
    assert x.min() >= -1.0 and x.max() < 1.0 and abs(x.mean()) < 0.02


@pytest.mark.parametrize("m,lmax", [(2000, 300), (50000, 10000), (300, 1000)])
def test_powerlaw_generator_matches_definition(pk, m, lmax):
    ai, aj, aa = pk.gen_powerlaw(m, lmax=lmax)
    ri, rj, ra = gen.powerlaw(m, lmax=lmax)
    assert np.array_equal(ai, ri) and np.array_equal(aj, rj) and np.array_equal(aa, ra)
    lens = np.diff(ai)
    assert lens.min() >= 1 and lens.max() <= min(lmax, m)
