"""BASELINE.json's full sizes on the GPU: the 300^3 matrix (configs[1]) bit-exact against the oracle
and through size-independent properties; its 8-rank decomposition (configs[2]) on one device."""
import numpy as np
import pytest

import gen
import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def poisson300(pk, cuda):
    g = pk.gen_poisson7(300, 1, 0, vectors=True)
    A = pk.Csr(g["ai"], g["aj"], g["aa"])
    yield g, A
    A.destroy()


def test_300_cubed_matmult_bit_exact(pk, cuda, poisson300):
    torch = cuda
    g, A = poisson300
    m = A.m
    assert (m, A.nz) == (27_000_000, 188_460_000)
    assert g["aa"][g["ai"][2] - 1] == 89999.99999999999          # 1/(dx*dx) = last entry of row 1, SURVEY 7.1b
    x = pk.gen_vector(m, 0xB200)
    dx = torch.from_numpy(x).cuda()
    dy = torch.empty(m, dtype=torch.float64, device="cuda")
    ref = oracle.matmult(g["ai"], g["aj"], g["aa"], x)
    A.mult(dx, dy, pk.MODE_EXACT)
    y_stream = dy.cpu().numpy()
    assert np.array_equal(y_stream, ref)
    A.set_kernel(pk.KERNEL_ROW)
    A.mult(dx, dy, pk.MODE_EXACT)
    A.set_kernel(pk.KERNEL_AUTO)
    assert np.array_equal(dy.cpu().numpy(), ref)                  # two independent kernels agree
    A.mult(dx, dy, pk.MODE_FAST)
    bound = 1e-13 * oracle.row_abs_sum(g["ai"], g["aj"], g["aa"], x)
    assert np.all(np.abs(dy.cpu().numpy() - ref) <= bound)
    # host-vector entry (pipelined in ~1M-row blocks) gives the same bits
    hx, hy = pk.PinnedArray(m), pk.PinnedArray(m)
    hx.array[:] = x
    A.mult_host(hx.array, hy.array, pk.MODE_EXACT)
    assert np.array_equal(hy.array, ref)
    hx.free(); hy.free()


def test_300_cubed_properties(pk, cuda, poisson300):
    torch = cuda
    g, A = poisson300
    m = A.m
    x = torch.from_numpy(pk.gen_vector(m, 1)).cuda()
    z = torch.from_numpy(pk.gen_vector(m, 2)).cuda()
    ax, az, axz = (torch.empty(m, dtype=torch.float64, device="cuda") for _ in range(3))
    A.mult(x, ax, pk.MODE_EXACT)
    A.mult(z, az, pk.MODE_EXACT)
    # linearity within rounding: A(x + 2z) = Ax + 2Az
    A.mult(x + 2.0 * z, axz, pk.MODE_EXACT)
    scale = 7 * 90000.0 * 3.0
    assert float((axz - (ax + 2.0 * az)).abs().max()) <= 1e-12 * scale
    # symmetry of the operator away from the reference point: x'(A z) = z'(A x) with A^T = A
    # except row/column 0; use MatMultTranspose for the exact identity x'(A z) = (A^T x)'z
    atx = torch.empty(m, dtype=torch.float64, device="cuda")
    A.mult_transpose(x, atx, pk.MODE_EXACT)
    lhs, rhs = float(torch.dot(x, az)), float(torch.dot(atx, z))
    assert abs(lhs - rhs) <= 1e-9 * abs(lhs) + 1e-3
    # Neumann rows sum to zero: A * ones = 0 up to the rounding of the diagonal, except near cell 0
    ones = torch.ones(m, dtype=torch.float64, device="cuda")
    A.mult(ones, ax, pk.MODE_EXACT)
    r = ax.cpu().numpy()
    touched = {0, 1, 300, 90000}
    mask = np.ones(m, bool)
    mask[list(touched)] = False
    assert np.abs(r[mask]).max() <= 1e-9
    # the discrete operator applied to the exact solution reproduces the right-hand side to O(h^2)
    ex = torch.from_numpy(g["exact"]).cuda()
    A.mult(ex, ax, pk.MODE_EXACT)
    res = (ax.cpu().numpy() - g["rhs"])
    assert np.abs(res[mask]).max() < 0.05 * np.abs(g["rhs"]).max()


def test_300_cubed_eight_rank_decomposition_on_one_gpu(pk, cuda):
    """configs[2] sizes: per rank 3,375,000 rows, 23,490,000 + 67,500 non-zeros, 67,500 ghosts;
    A x_local + B x_ghost bit-exact for rank 0 and rank 7 (all eight pushes launched first)."""
    torch = cuda
    size, N = 8, 300
    ranks, base = [], None
    for r in range(size):
        g = pk.gen_poisson7(N, size, r)
        base = g["base"]
        M = pk.MpiAij(size, r, base, g["ai"], g["aj"], g["aa"])
        assert (M.nloc, M.annz, M.bnnz, M.nghost, M.brows, M.nsrc) == (3_375_000, 23_490_000, 67_500, 67_500, 67_051, 3)
        ranks.append(M)
    garrays = [M.garray() for M in ranks]
    for M in ranks:
        for q in range(size):
            M.set_peer_garray(q, garrays[q])
        M.upload()
    for M in ranks:
        for q in range(size):
            if q != M.rank:
                M.set_peer_window(q, ranks[q].window_ptr())
    xg = pk.gen_vector(N ** 3, 0xB200)
    xs = [torch.from_numpy(xg[base[r]:base[r + 1]].copy()).cuda() for r in range(size)]
    ys = [torch.empty(ranks[r].nloc, dtype=torch.float64, device="cuda") for r in range(size)]
    for r, M in enumerate(ranks):
        M.mult_begin(xs[r])
    for r, M in enumerate(ranks):
        M.mult_finish(xs[r], ys[r], pk.MODE_EXACT)
    torch.cuda.synchronize()
    for r in (0, 7):
        M = ranks[r]
        M.check()
        Ai, Aj, Aa = M.block(0)
        Bi, Bj, Ba = M.block(1)
        ref = oracle.matmultadd(Bi, Bj, Ba, xg[garrays[r]], oracle.matmult(Ai, Aj, Aa, xg[base[r]:base[r + 1]]))
        assert np.array_equal(ys[r].cpu().numpy(), ref)
    for M in ranks:
        M.destroy()


def test_powerlaw_1m_merge_and_transpose(pk, cuda):
    torch = cuda
    ai, aj, aa = pk.gen_powerlaw(1_000_000)   # product generator (== tests/gen.py::powerlaw, tests/test_generator.py)
    m = len(ai) - 1
    A = pk.Csr(ai, aj, aa)
    assert pk.KERNEL_NAMES[A.info().kernel_fast] == "merge"
    x = gen.uniform_pm1(m, 3)
    dx = torch.from_numpy(x).cuda()
    dy = torch.empty(m, dtype=torch.float64, device="cuda")
    ref = oracle.matmult(ai, aj, aa, x)
    bound = 1e-13 * oracle.row_abs_sum(ai, aj, aa, x)
    A.mult(dx, dy, pk.MODE_FAST)
    assert np.all(np.abs(dy.cpu().numpy() - ref) <= bound)
    A.mult(dx, dy, pk.MODE_EXACT)
    assert np.array_equal(dy.cpu().numpy(), ref)
    A.mult_transpose(dx, dy, pk.MODE_EXACT)
    assert np.array_equal(dy.cpu().numpy(), oracle.matmulttranspose(ai, aj, aa, x, m))
    A.destroy()


def test_stencil27_200_cubed_bit_exact(pk, cuda):
    """BASELINE configs[3] at full size: 8,000,000 rows, (3*200-2)^3 = 213,847,192 non-zeros."""
    torch = cuda
    ai, aj, aa = gen.stencil27(200)
    m = len(ai) - 1
    assert (m, len(aj)) == (8_000_000, 213_847_192)
    A = pk.Csr(ai, aj, aa)
    assert pk.KERNEL_NAMES[A.info().kernel_exact] == "stream" and A.info().index8_diagonals == 27
    x = pk.gen_vector(m, 0xB200)
    dx = torch.from_numpy(x).cuda()
    dy = torch.empty(m, dtype=torch.float64, device="cuda")
    ref = oracle.matmult(ai, aj, aa, x)
    A.mult(dx, dy, pk.MODE_EXACT)
    assert np.array_equal(dy.cpu().numpy(), ref)
    A.mult(dx, dy, pk.MODE_EXACT_FMA)
    assert np.array_equal(dy.cpu().numpy(), oracle.matmult(ai, aj, aa, x, fma=True))
    # box stencil: off-diagonals -1, diagonal = number of neighbours -> every row sums to zero
    A.mult(torch.ones(m, dtype=torch.float64, device="cuda"), dy, pk.MODE_EXACT)
    assert float(dy.abs().max()) == 0.0
    A.destroy()


def test_powerlaw_10m_bit_exact(pk, cuda):
    """BASELINE configs[4] at full size: 10 M rows, lengths 1..10,000, MatMult and MatMultTranspose."""
    torch = cuda
    ai, aj, aa = pk.gen_powerlaw(10_000_000)
    m = len(ai) - 1
    lens = np.diff(ai)
    assert lens.min() >= 1 and lens.max() == 10_000 and 9.0 < lens.mean() < 10.5
    A = pk.Csr(ai, aj, aa)
    assert pk.KERNEL_NAMES[A.info().kernel_exact] == "merge"
    x = pk.gen_vector(m, 3)
    dx = torch.from_numpy(x).cuda()
    dy = torch.empty(m, dtype=torch.float64, device="cuda")
    ref = oracle.matmult(ai, aj, aa, x)
    A.mult(dx, dy, pk.MODE_EXACT)
    assert np.array_equal(dy.cpu().numpy(), ref)
    A.mult(dx, dy, pk.MODE_FAST)
    assert np.all(np.abs(dy.cpu().numpy() - ref) <= 1e-13 * oracle.row_abs_sum(ai, aj, aa, x))
    A.mult_transpose(dx, dy, pk.MODE_EXACT)
    assert np.array_equal(dy.cpu().numpy(), oracle.matmulttranspose(ai, aj, aa, x, m))
    A.destroy()
