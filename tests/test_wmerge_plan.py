"""The plan of k_wmerge (the warp-granular exact-order kernel for skewed matrices), on the CPU: the
chunks are walked exactly the way the kernel walks them -- products staged per chunk, the rows of a
chunk summed left to right one after the other, a long row carried from piece to piece -- and the
result must be the oracle's bits; plus the invariants the kernel relies on."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import gen
import oracle

WM_CAP = 128


def walk_like_k_wmerge(ai, aj, aa, x, chunks, blk, y0=None, fma=False):
    m = len(ai) - 1
    y = np.full(m, np.nan)
    carry = None
    for b in range(len(blk) - 1):
        for c in range(blk[b], blk[b + 1]):
            r0, kind, k0, k1 = (int(v) for v in chunks[c])
            a, xv = aa[k0:k1], x[aj[k0:k1]]
            prod = a * xv                                      # rounded products (EXACT stages these)
            if kind >= 0:
                for r in range(r0, r0 + kind):
                    s = 0.0 if y0 is None else y0[r]
                    for k in range(ai[r] - k0, ai[r + 1] - k0):
                        s = float(np.float64(s) + prod[k]) if not fma else oracle_fma(a[k], xv[k], s)
                    y[r] = s
            else:
                if kind == -2:
                    carry = 0.0 if y0 is None else y0[r0]
                for k in range(k1 - k0):
                    carry = float(np.float64(carry) + prod[k]) if not fma else oracle_fma(a[k], xv[k], carry)
                if kind == -3:
                    y[r0] = carry
    return y


def oracle_fma(a, b, c):
    import math
    return math.fma(a, b, c) if hasattr(math, "fma") else float(np.float64(a) * np.float64(b) + np.float64(c))


def check_invariants(ai, chunks, blk):
    m, nz = len(ai) - 1, int(ai[-1])
    assert blk[0] == 0 and blk[-1] == len(chunks) and np.all(np.diff(blk) > 0 if len(chunks) else True)
    row, k = 0, 0
    for c, (r0, kind, k0, k1) in enumerate(chunks):
        assert 0 <= k1 - k0 <= WM_CAP and k0 == k                      # contiguous, never more than a warp's slice
        if kind >= 0:
            assert r0 == row and 1 <= kind <= WM_CAP and k0 == ai[r0] and k1 == ai[r0 + kind]
            row += kind
        else:
            assert r0 == row and ai[r0 + 1] - ai[r0] > WM_CAP
            assert (kind == -2) == (k0 == ai[r0]) and (kind == -3) == (k1 == ai[r0 + 1])
            if kind == -3:
                row += 1
        k = k1
    assert row == m and k == nz
    starts = set(int(b) for b in blk[:-1])
    for c, ch in enumerate(chunks):
        if ch[1] in (-1, -3):
            assert c not in starts, "a work block begins in the middle of a long row"


CASES = {
    "powerlaw": lambda: gen.powerlaw(4000, lmax=3000, seed=3),
    "poisson": lambda: (lambda p: (p["ai"], p["aj"], p["aa"]))(oracle.poisson7(9)),
    "mostly_empty": lambda: gen.random_csr(3000, 200, 3, np.random.default_rng(1), empty_frac=0.9),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_walk_gives_the_oracle_bits(pk, name):
    ai, aj, aa = CASES[name]()
    n = int(aj.max()) + 1 if len(aj) else 1
    x, y0 = gen.uniform_pm1(n, 3), gen.uniform_pm1(len(ai) - 1, 4)
    chunks, blk = pk.wmerge_plan(ai)
    check_invariants(ai, chunks, blk)
    assert np.array_equal(walk_like_k_wmerge(ai, aj, aa, x, chunks, blk), oracle.matmult(ai, aj, aa, x))
    assert np.array_equal(walk_like_k_wmerge(ai, aj, aa, x, chunks, blk, y0), oracle.matmultadd(ai, aj, aa, x, y0))


def test_row_lengths_around_the_chunk_capacity(pk):
    """0, 1, 127, 128 (still a whole-row chunk), 129, 256, 257 and 1000 entries, and 300 empty rows in a row."""
    lens = [0, 1, 127, 128, 129, 0, 256, 257, 3, 1000, 128, 128, 1] + [0] * 300 + [5, 128]
    ai = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    rng = np.random.default_rng(5)
    ncol = 1200
    aj = np.concatenate([np.sort(rng.choice(ncol, size=l, replace=False)) for l in lens]).astype(np.int32)
    aa = gen.uniform_pm1(len(aj), 6)
    chunks, blk = pk.wmerge_plan(ai)
    check_invariants(ai, chunks, blk)
    kinds = [int(c[1]) for c in chunks]
    assert kinds.count(-2) == 4 and kinds.count(-3) == 4        # the rows of 129, 256, 257 and 1000 entries
    x = gen.uniform_pm1(ncol, 7)
    assert np.array_equal(walk_like_k_wmerge(ai, aj, aa, x, chunks, blk), oracle.matmult(ai, aj, aa, x))


@settings(max_examples=40, deadline=None)
@given(st.lists(st.one_of(st.integers(0, 6), st.integers(100, 140), st.integers(250, 700)), min_size=0, max_size=60))
def test_invariants_for_arbitrary_row_lengths(pk, lens):
    ai = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    chunks, blk = pk.wmerge_plan(ai)
    check_invariants(ai, chunks, blk)


def test_column_blocks_continue_every_row_left_to_right(pk):
    """A = [A_0 | A_1 | ...]: no blocking up to 1.5 blocks of x; otherwise every row is cut at ascending
    column bounds, the pieces are contiguous and in order, and summing them block after block -- each
    block starting from the previous block's y, as MatMultAdd does -- gives the oracle's bits."""
    ai, aj, aa = gen.powerlaw(3000, lmax=2000, seed=9)
    n = 3000
    assert pk.colblock_split(ai, aj, n, n * 8)[0] == 0 and pk.colblock_split(ai, aj, n, int(n * 8 / 1.5))[0] == 0
    nb, split = pk.colblock_split(ai, aj, n, 4096)
    assert nb == -(-n * 8 // 4096) and split.shape == (nb, 3000)
    x = gen.uniform_pm1(n, 2)
    y = np.zeros(3000)
    lo = ai[:-1].astype(np.int64)
    for b in range(nb):
        hi = split[b].astype(np.int64)
        assert np.all(hi >= lo) and np.all(hi <= ai[1:])
        cend = n if b + 1 == nb else n * (b + 1) // nb
        for r in range(3000):
            seg = aj[lo[r]:hi[r]]
            assert np.all(seg < cend) and (hi[r] == ai[r + 1] or aj[hi[r]] >= cend)
            s = y[r]
            for k in range(lo[r], hi[r]):
                s = float(np.float64(s) + np.float64(aa[k]) * np.float64(x[aj[k]]))
            y[r] = s
        lo = hi
    assert np.array_equal(lo, ai[1:])
    assert np.array_equal(y, oracle.matmult(ai, aj, aa, x))
