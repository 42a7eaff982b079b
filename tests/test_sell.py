"""The optional SELL-32-sigma copy (north star: "optional SELL-C-sigma or ELL-sliced copy for
stencil matrices"): b200_sell_pack / b200_csr_build_sell / B200_KERNEL_SELL.

CPU: the host packing, checked by walking the packed arrays exactly the way k_sell does (lane l of a
chunk reads base + 32 k + l, skips padding, adds left to right) -- the result must be the oracle's
bits, for every sigma.
GPU: the kernel itself against the oracle."""
import numpy as np
import pytest

import gen
import oracle


def walk_like_k_sell(m, cs, perm, val, col, x, y0=None):
    """numpy/Python restatement of k_sell's traversal (unfused multiply-add = MODE_EXACT)."""
    y = np.full(m, np.nan)
    nchunks = len(cs) - 1
    for c in range(nchunks):
        base, length = int(cs[c]), (int(cs[c + 1]) - int(cs[c])) // 32
        for lane in range(32):
            row = int(perm[c * 32 + lane])
            if row < 0:
                # a padding slot owns no entries
                assert all(col[base + k * 32 + lane] == -1 for k in range(length))
                continue
            s = np.float64(0.0 if y0 is None else y0[row])
            for k in range(length):
                at = base + k * 32 + lane
                if col[at] >= 0:
                    s = s + np.float64(val[at]) * np.float64(x[col[at]])
            y[row] = s
    return y


def cases():
    rng = np.random.default_rng(5)
    p = oracle.poisson7(9)
    out = {"poisson7_9": (p["ai"], p["aj"], p["aa"], 9 ** 3)}
    ai, aj, aa = gen.stencil27(6, seed=4)
    out["stencil27_6"] = (ai, aj, aa, 6 ** 3)
    ai, aj, aa = gen.random_csr(333, 250, 20, rng, empty_frac=0.25)
    out["random_333x250"] = (ai, aj, aa, 250)
    ai, aj, aa = gen.powerlaw(700, lmax=300)
    out["powerlaw_700"] = (ai, aj, aa, 700)
    out["one_row"] = (np.array([0, 3], np.int32), np.array([0, 2, 5], np.int32), np.array([1.5, -2.0, 0.25]), 6)
    out["empty"] = (np.zeros(41, np.int32), np.zeros(0, np.int32), np.zeros(0), 7)
    return out


CASES = cases()


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("sigma", [1, 32, 128, 100000])
def test_pack_walk_gives_the_oracle_bits(pk, name, sigma):
    ai, aj, aa, n = CASES[name]
    m = len(ai) - 1
    cs, perm, val, col = pk.sell_pack(ai, aj, aa, sigma)
    lens = np.diff(ai)
    # layout invariants: every row exactly once; windows keep their rows; chunk length = longest row
    rows = perm[perm >= 0]
    assert np.array_equal(np.sort(rows), np.arange(m))
    if sigma == 1:
        assert np.array_equal(perm[:m], np.arange(m))
    for w in range(0, m, sigma):
        win = perm[w:min(m, w + sigma)]
        assert np.array_equal(np.sort(win), np.arange(w, min(m, w + sigma)))
        assert np.all(np.diff(lens[win]) <= 0)                       # decreasing length inside a window
    for c in range(len(cs) - 1):
        slot_rows = perm[c * 32:(c + 1) * 32]
        longest = max([lens[r] for r in slot_rows if r >= 0], default=0)
        assert int(cs[c + 1]) - int(cs[c]) == 32 * longest
    assert int((col >= 0).sum()) == len(aj)
    x = gen.uniform_pm1(n, 3)
    assert np.array_equal(walk_like_k_sell(m, cs, perm, val, col, x), oracle.matmult(ai, aj, aa, x))
    y0 = gen.uniform_pm1(m, 4)
    assert np.array_equal(walk_like_k_sell(m, cs, perm, val, col, x, y0), oracle.matmultadd(ai, aj, aa, x, y0))


def test_padding_overhead_of_the_reference_matrix(pk):
    """SELL-32-1 on the 7-point matrix pads a chunk to 7 entries per row: under 2 % at 20^3 and
    shrinking with the grid (boundary rows are the only short ones)."""
    p = oracle.poisson7(20)
    cs, perm, val, col = pk.sell_pack(p["ai"], p["aj"], p["aa"], 1)
    assert len(val) <= 1.12 * len(p["aj"])
    cs2, _, val2, _ = pk.sell_pack(p["ai"], p["aj"], p["aa"], 4096)
    assert len(val2) <= len(val)                                     # sorting can only reduce padding


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("sigma", [1, 64])
def test_sell_kernel_bit_exact(pk, cuda, name, sigma):
    torch = cuda
    ai, aj, aa, n = CASES[name]
    m = len(ai) - 1
    A = pk.Csr(ai, aj, aa, n=n)
    A.build_sell(sigma)
    info = A.info()
    assert info.sell_sigma == sigma and info.sell_chunks == (m + 31) // 32
    A.set_kernel(pk.KERNEL_SELL)
    x, y0 = gen.uniform_pm1(n, 3), gen.uniform_pm1(m, 4)
    dx, dy0 = torch.from_numpy(x).cuda(), torch.from_numpy(y0).cuda()
    dy = torch.full((m,), float("nan"), dtype=torch.float64, device="cuda")
    for mode, fma in ((pk.MODE_EXACT, False), (pk.MODE_EXACT_FMA, True), (pk.MODE_FAST, True)):
        A.mult(dx, dy, mode)
        assert np.array_equal(dy.cpu().numpy(), oracle.matmult(ai, aj, aa, x, fma=fma)), (name, sigma, mode)
        A.mult_add(dx, dy0, dy, mode)
        assert np.array_equal(dy.cpu().numpy(), oracle.matmultadd(ai, aj, aa, x, y0, fma=fma)), (name, sigma, mode)
    if m == n:
        out = torch.full((m,), float("nan"), dtype=torch.float64, device="cuda")
        A.residual(dx, dy0, out, pk.MODE_EXACT)
        assert np.array_equal(out.cpu().numpy(), oracle.residual(ai, aj, aa, x, y0))
    A.update_values(aa * 2.0)                      # drops the copy and the override
    assert A.info().sell_chunks == 0
    A.mult(dx, dy, pk.MODE_EXACT)
    assert np.array_equal(dy.cpu().numpy(), oracle.matmult(ai, aj, aa * 2.0, x))
    A.destroy()
