"""Randomised shape sweep of the kernel plans: many small and medium matrices whose mean row length,
number of distinct diagonals, emptiness and skew are drawn so that every plan branch is hit
(stream with 128/256 consumer threads, byte-coded or int32 indices, 1- and 2-deep rings, exact-order
merge, long rows, compressed row), every operation against the oracle bit for bit.

Written after the multigrid levels exposed a plan-dependent fault in k_stream (byte codes x 160-thread
CTAs) that no hand-picked case had hit (sorted last on purpose)."""

import numpy as np
import pytest

import gen
import oracle

pytestmark = pytest.mark.gpu


def banded(m, n, mean, ndiag, rng, empty_frac=0.0):
    """Rows with ~mean entries drawn from a fixed set of ndiag diagonals (so the byte-code plan
    applies when ndiag <= 256), ascending columns."""
    diags = np.sort(rng.choice(np.arange(-(m - 1), n), size=min(ndiag, m + n - 1), replace=False))
    ai = np.zeros(m + 1, dtype=np.int32)
    cols = []
    for r in range(m):
        ok = diags[(diags + r >= 0) & (diags + r < n)]
        k = 0 if rng.random() < empty_frac else min(len(ok), max(1, int(rng.poisson(mean))))
        c = np.sort(rng.choice(ok, size=k, replace=False)) + r if k else np.zeros(0, dtype=np.int64)
        cols.append(c)
        ai[r + 1] = ai[r] + len(c)
    aj = np.concatenate(cols).astype(np.int32) if cols else np.zeros(0, np.int32)
    return ai, aj, rng.uniform(-1, 1, size=len(aj))


def sweep_cases():
    rng = np.random.default_rng(2026)
    out = []
    for mean in (2, 6, 13, 24, 48, 90, 200):
        for ndiag in (5, 150, 161, 200, 256, 257, 2000):
            m = int(rng.integers(40, 1500))
            n = m if rng.random() < 0.6 else int(rng.integers(20, 1500))
            out.append((f"banded_m{m}_n{n}_mean{mean}_d{ndiag}", banded(m, n, mean, ndiag, rng, empty_frac=float(rng.choice([0, 0.1, 0.7])))))
    for lmax in (40, 400, 4000):
        ai, aj, aa = gen.powerlaw(3000, lmax=lmax, seed=lmax)
        out.append((f"powerlaw_lmax{lmax}", (ai, aj, aa)))
    return out


SWEEP = sweep_cases()


@pytest.mark.parametrize("name,mat", SWEEP, ids=[c[0] for c in SWEEP])
def test_every_operation_bit_exact(pk, cuda, name, mat):
    torch = cuda
    ai, aj, aa = mat
    m = len(ai) - 1
    n = int(name.split("_n")[1].split("_")[0]) if "_n" in name else m
    A = pk.Csr(ai, aj, aa, n=n)
    x, xt, y0, z0 = gen.uniform_pm1(n, 1), gen.uniform_pm1(m, 2), gen.uniform_pm1(m, 3), gen.uniform_pm1(n, 4)
    dx, dxt, dy0, dz0 = (torch.from_numpy(v).cuda() for v in (x, xt, y0, z0))
    dy = torch.full((m,), float("nan"), dtype=torch.float64, device="cuda")
    dyt = torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
    for mode, fma in ((pk.MODE_EXACT, False), (pk.MODE_EXACT_FMA, True)):
        A.mult(dx, dy, mode)
        assert np.array_equal(dy.cpu().numpy(), oracle.matmult(ai, aj, aa, x, fma=fma)), (name, mode, pk.KERNEL_NAMES[A.info().kernel_exact])
        A.mult_add(dx, dy0, dy, mode)
        assert np.array_equal(dy.cpu().numpy(), oracle.matmultadd(ai, aj, aa, x, y0, fma=fma)), (name, mode)
    A.mult_transpose(dxt, dyt, pk.MODE_EXACT)
    assert np.array_equal(dyt.cpu().numpy(), oracle.matmulttranspose(ai, aj, aa, xt, n)), name
    A.mult_transpose_add(dxt, dz0, dyt, pk.MODE_EXACT)
    assert np.array_equal(dyt.cpu().numpy(), oracle.matmulttransposeadd(ai, aj, aa, xt, z0, n)), name
    A.residual(dx, dy0, dy, pk.MODE_EXACT)
    assert np.array_equal(dy.cpu().numpy(), oracle.residual(ai, aj, aa, x, y0)), name
    # FAST: the stated tolerance, whatever kernel the plan picked
    A.mult(dx, dy, pk.MODE_FAST)
    ref = oracle.matmult(ai, aj, aa, x)
    assert np.all(np.abs(dy.cpu().numpy() - ref) <= 1e-13 * oracle.row_abs_sum(ai, aj, aa, x) + 0.0), name
    torch.cuda.synchronize()
    A.destroy()
