"""GPU parity: the CUDA path (through the C ABI) against the oracle on the same seeded inputs.

Bars (north star):
  * EXACT / EXACT_FMA modes: bit-exact against oracle.matmult / matmult(fma=True);
  * FAST mode: |y - y_ref| <= 1e-13 * sum_j |a_ij x_j| per row;
  * integer work (plans, compressed-row index) bit-exact.
"""
import os

import numpy as np
import pytest

import gen
import oracle

pytestmark = pytest.mark.gpu

TOL = 1e-13


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _run(pk, torch, A, x, mode, kernel=None, add=None):
    if kernel is not None:
        A.set_kernel(kernel)
    dx = _dev(torch, x)
    dy = torch.full((A.m,), float("nan"), dtype=torch.float64, device="cuda")
    if add is None:
        A.mult(dx, dy, mode)
    else:
        A.mult_add(dx, _dev(torch, add), dy, mode)
    torch.cuda.synchronize()
    if kernel is not None:
        A.set_kernel(pk.KERNEL_AUTO)
    return dy.cpu().numpy()


def _cases():
    rng = np.random.default_rng(7)
    cases = {}
    p = oracle.poisson7(20)
    cases["poisson7_20"] = (p["ai"], p["aj"], p["aa"], 20 ** 3)
    p = oracle.poisson7(50)
    cases["poisson7_50"] = (p["ai"], p["aj"], p["aa"], 50 ** 3)
    ai, aj, aa = gen.stencil27(16, seed=3)
    cases["stencil27_16"] = (ai, aj, aa, 16 ** 3)
    ai, aj, aa = gen.powerlaw(20000, lmax=3000)
    cases["powerlaw_20k"] = (ai, aj, aa, 20000)
    ai, aj, aa = gen.random_csr(1000, 700, 40, rng, empty_frac=0.3)
    cases["random_ragged"] = (ai, aj, aa, 700)
    ai, aj, aa = gen.random_csr(5000, 300, 3, rng, empty_frac=0.9)
    cases["mostly_empty"] = (ai, aj, aa, 300)
    # long rows: several times the merge tile capacity (2048), neighbours of every length
    lens = np.array([1, 0, 5000, 3, 2048, 2049, 0, 0, 7, 4096, 1, 10000, 2, 2047, 300] * 3)
    ai = np.zeros(len(lens) + 1, np.int32)
    np.cumsum(lens, out=ai[1:])
    ncol = 12000
    aj = np.concatenate([np.sort(rng.choice(ncol, size=l, replace=False)) for l in lens]).astype(np.int32)
    cases["long_rows"] = (ai, aj, rng.uniform(-1, 1, size=len(aj)), ncol)
    cases["single_row"] = (np.array([0, 3], np.int32), np.array([0, 2, 4], np.int32), np.array([1.5, -2.0, 0.25]), 5)
    cases["all_empty"] = (np.zeros(11, np.int32), np.zeros(0, np.int32), np.zeros(0), 4)
    return cases


CASES = _cases()


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("fma", [False, True])
def test_exact_modes_bit_exact(pk, cuda, name, fma):
    ai, aj, aa, n = CASES[name]
    x = gen.uniform_pm1(n, seed=0xB200)
    ref = oracle.matmult(ai, aj, aa, x, fma=fma)
    A = pk.Csr(ai, aj, aa, n=n)
    mode = pk.MODE_EXACT_FMA if fma else pk.MODE_EXACT
    info = A.info()
    kernels = [None, pk.KERNEL_ROW]
    if info.stream_tiles:
        kernels.append(pk.KERNEL_STREAM)
    if info.compressedrow_use:
        kernels.append(pk.KERNEL_CPROW)
    for k in kernels:
        y = _run(pk, cuda, A, x, mode, kernel=k)
        assert np.array_equal(y, ref), f"{name} kernel={k} fma={fma}: max diff {np.abs(y - ref).max()}"
    A.destroy()


@pytest.mark.parametrize("name", sorted(CASES))
def test_fast_mode_within_row_bound(pk, cuda, name):
    ai, aj, aa, n = CASES[name]
    x = gen.uniform_pm1(n, seed=0xB200)
    ref = oracle.matmult(ai, aj, aa, x)
    bound = TOL * oracle.row_abs_sum(ai, aj, aa, x)
    A = pk.Csr(ai, aj, aa, n=n)
    info = A.info()
    kernels = [None, pk.KERNEL_ROW, pk.KERNEL_VECTOR]
    if info.stream_tiles:
        kernels.append(pk.KERNEL_STREAM)
    if info.nz:
        kernels.append(pk.KERNEL_MERGE)
    for k in kernels:
        y = _run(pk, cuda, A, x, pk.MODE_FAST, kernel=k)
        assert np.all(np.abs(y - ref) <= bound), f"{name} kernel={k}: {np.max(np.abs(y - ref) - bound)}"
    if info.nz:
        # the split-row merge kernel stays covered even where FAST now prefers the exact-order kernels
        os.environ["B200_MERGE_SPLIT"] = "1"
        ys = _run(pk, cuda, A, x, pk.MODE_FAST, kernel=pk.KERNEL_MERGE)
        os.environ.pop("B200_MERGE_SPLIT")
        assert np.all(np.abs(ys - ref) <= bound), f"{name} split-row merge"
        # merge path: deterministic from run to run (no atomics), also for MatMultAdd
        y1 = _run(pk, cuda, A, x, pk.MODE_FAST, kernel=pk.KERNEL_MERGE)
        y2 = _run(pk, cuda, A, x, pk.MODE_FAST, kernel=pk.KERNEL_MERGE)
        assert np.array_equal(y1, y2)
        y0 = gen.uniform_pm1(len(ai) - 1, seed=11)
        z = _run(pk, cuda, A, x, pk.MODE_FAST, kernel=pk.KERNEL_MERGE, add=y0)
        refa = oracle.matmultadd(ai, aj, aa, x, y0)
        assert np.all(np.abs(z - refa) <= bound + 1e-15 * np.abs(y0))
    A.destroy()


def test_plan_picks_kernel_from_histogram(pk, cuda):
    p = oracle.poisson7(16)
    A = pk.Csr(p["ai"], p["aj"], p["aa"])
    assert pk.KERNEL_NAMES[A.info().kernel_fast] == "stream" and pk.KERNEL_NAMES[A.info().kernel_exact] == "stream"
    A.destroy()
    ai, aj, aa = gen.powerlaw(30000, lmax=5000)
    A = pk.Csr(ai, aj, aa)
    i = A.info()
    assert pk.KERNEL_NAMES[i.kernel_fast] == "merge" and i.merge_tiles > 0
    assert pk.KERNEL_NAMES[i.kernel_exact] == "merge"     # exact order at merge speed (whole-row tiles)
    assert sum(i.hist) == 30000
    A.destroy()
    rng = np.random.default_rng(0)
    ai, aj, aa = gen.random_csr(4000, 100, 2, rng, empty_frac=0.9)
    A = pk.Csr(ai, aj, aa, n=100)
    assert pk.KERNEL_NAMES[A.info().kernel_fast] == "cprow"
    A.destroy()


@pytest.mark.parametrize("name", sorted(CASES))
def test_multadd_bit_exact(pk, cuda, name):
    ai, aj, aa, n = CASES[name]
    m = len(ai) - 1
    x = gen.uniform_pm1(n, seed=1)
    y0 = gen.uniform_pm1(m, seed=2)
    y0[::7] = -0.0
    ref = oracle.matmultadd(ai, aj, aa, x, y0)
    A = pk.Csr(ai, aj, aa, n=n)
    info = A.info()
    kernels = [None, pk.KERNEL_ROW] + ([pk.KERNEL_STREAM] if info.stream_tiles else []) + (
        [pk.KERNEL_CPROW] if info.compressedrow_use else [])
    for k in kernels:
        z = _run(pk, cuda, A, x, pk.MODE_EXACT, kernel=k, add=y0)
        assert np.array_equal(z, ref), f"{name} kernel={k}"
        assert np.array_equal(np.signbit(z), np.signbit(ref))
    # in place (z aliases y), the MatMult_MPIAIJ use: yy = B*lvec + yy
    dx = cuda.from_numpy(x).cuda()
    dy = cuda.from_numpy(y0.copy()).cuda()
    A.mult_add(dx, dy, dy, pk.MODE_EXACT)
    assert np.array_equal(dy.cpu().numpy(), ref)
    A.destroy()


@pytest.mark.parametrize("name", sorted(CASES))
def test_transpose(pk, cuda, name):
    ai, aj, aa, n = CASES[name]
    m = len(ai) - 1
    x = gen.uniform_pm1(m, seed=5)
    A = pk.Csr(ai, aj, aa, n=n)
    dx = cuda.from_numpy(x).cuda()
    dy = cuda.full((n,), float("nan"), dtype=cuda.float64, device="cuda")
    ref = oracle.matmulttranspose(ai, aj, aa, x, n)
    A.mult_transpose(dx, dy, pk.MODE_EXACT)
    assert np.array_equal(dy.cpu().numpy(), ref)
    ref_f = oracle.matmulttranspose(ai, aj, aa, x, n, fma=True)
    A.mult_transpose(dx, dy, pk.MODE_EXACT_FMA)
    assert np.array_equal(dy.cpu().numpy(), ref_f)
    # transpose-add
    z = gen.uniform_pm1(n, seed=6)
    refa = oracle.matmulttransposeadd(ai, aj, aa, x, z, n)
    dz = cuda.from_numpy(z).cuda()
    A.mult_transpose_add(dx, dz, dy, pk.MODE_EXACT)
    assert np.array_equal(dy.cpu().numpy(), refa)
    # fast mode within the column bound
    A.mult_transpose(dx, dy, pk.MODE_FAST)
    absb = np.zeros(n)
    np.add.at(absb, aj, np.abs(aa * np.repeat(x, np.diff(ai))))
    assert np.all(np.abs(dy.cpu().numpy() - ref) <= TOL * absb)
    A.destroy()


def test_host_vector_entry(pk, cuda):
    p = oracle.poisson7(30)
    A = pk.Csr(p["ai"], p["aj"], p["aa"])
    x = gen.uniform_pm1(A.n)
    y = A.mult_host(x, mode=pk.MODE_EXACT)
    assert np.array_equal(y, oracle.matmult(p["ai"], p["aj"], p["aa"], x))
    y0 = gen.uniform_pm1(A.m, seed=9)
    z = A.mult_add_host(x, y0, mode=pk.MODE_EXACT)
    assert np.array_equal(z, oracle.matmultadd(p["ai"], p["aj"], p["aa"], x, y0))
    yt = A.mult_transpose_host(x, mode=pk.MODE_EXACT)
    assert np.array_equal(yt, oracle.matmulttranspose(p["ai"], p["aj"], p["aa"], x, A.n))
    A.destroy()


def test_plan_integers(pk, cuda):
    """compressed-row verdict and statistics are bit-exact against the oracle's restatement."""
    rng = np.random.default_rng(11)
    for empty in (0.0, 0.5, 0.61, 0.95):
        ai, aj, aa = gen.random_csr(2000, 500, 5, rng, empty_frac=empty)
        A = pk.Csr(ai, aj, aa, n=500)
        info = A.info()
        lens = np.diff(ai)
        assert info.nonzerorowcnt == int((lens > 0).sum())
        assert info.rmax == int(lens.max())
        use, cpi, ridx = oracle.check_compressed_row(ai, info.nonzerorowcnt)
        assert bool(info.compressedrow_use) == use
        if use:
            assert info.cprow_nrows == len(ridx)
        A.destroy()


def test_update_values(pk, cuda):
    p = oracle.poisson7(12)
    A = pk.Csr(p["ai"], p["aj"], p["aa"])
    aa2 = p["aa"] * 0.5 + 1.0
    A.update_values(aa2)
    x = gen.uniform_pm1(A.n)
    assert np.array_equal(A.mult_host(x, mode=pk.MODE_EXACT), oracle.matmult(p["ai"], p["aj"], aa2, x))
    A.destroy()


def test_launches_counted(pk, cuda):
    p = oracle.poisson7(10)
    A = pk.Csr(p["ai"], p["aj"], p["aa"])
    l0 = pk.launch_count()
    A.mult_host(gen.uniform_pm1(A.n))
    assert pk.launch_count() > l0
    A.destroy()


@pytest.mark.parametrize("case", ["poisson", "stencil27", "scattered"])
def test_host_vector_pipeline_blocks(pk, cuda, monkeypatch, case):
    """The row-blocked, three-stream host path (x chunks up / tiles / y rows down) gives the same
    bits as the single-shot path, for banded and for non-banded column patterns."""
    monkeypatch.setenv("B200_HOST_BLOCK_ROWS", "2048")
    rng = np.random.default_rng(3)
    if case == "poisson":
        p = oracle.poisson7(30)
        ai, aj, aa = p["ai"], p["aj"], p["aa"]
    elif case == "stencil27":
        ai, aj, aa = gen.stencil27(24, seed=9)
    else:
        ai, aj, aa = gen.random_csr(20000, 20000, 6, rng)
    n = len(ai) - 1
    A = pk.Csr(ai, aj, aa)
    assert A.info().stream_tiles > 8
    x, y0 = gen.uniform_pm1(n, 3), gen.uniform_pm1(n, 4)
    hx, hy = pk.PinnedArray(n), pk.PinnedArray(n)
    hx.array[:] = x
    for mode, fma in ((pk.MODE_EXACT, False), (pk.MODE_EXACT_FMA, True)):
        A.mult_host(hx.array, hy.array, mode)
        assert np.array_equal(hy.array, oracle.matmult(ai, aj, aa, x, fma=fma))
        hy.array[:] = y0
        A.mult_add_host(hx.array, hy.array, hy.array, mode)   # in place
        assert np.array_equal(hy.array, oracle.matmultadd(ai, aj, aa, x, y0, fma=fma))
    monkeypatch.setenv("B200_HOST_PIPELINE", "0")
    y1 = A.mult_host(x, mode=pk.MODE_EXACT)
    assert np.array_equal(y1, oracle.matmult(ai, aj, aa, x))
    hx.free(); hy.free()
    A.destroy()


def test_byte_codes_with_more_diagonals_than_cta_threads(pk, cuda, monkeypatch):
    """Regression (found by the multigrid levels, N=12 Galerkin operator: 218 rows, 169 diagonals):
    a matrix with long rows gets 128 consumer threads (160-thread CTAs); every one of the up to 256
    diagonal-table entries must still be loaded.  Codes >= 160 used to read an unset table slot."""
    m = 230
    rng = np.random.default_rng(7)
    ai = np.zeros(m + 1, dtype=np.int32)
    cols = []
    for r in range(m):
        c = np.sort(rng.choice(m, size=24, replace=False))   # dense-ish coarse operator: ~2m-1 > 160 diagonals
        cols.append(c)
        ai[r + 1] = ai[r] + len(c)
    aj = np.concatenate(cols).astype(np.int32)
    aa = rng.uniform(-1, 1, size=len(aj))
    # keep it under 257 distinct diagonals: clip the band
    keep = np.abs(aj - np.repeat(np.arange(m), 24)) <= 120
    lens = np.add.reduceat(keep.astype(np.int32), ai[:-1])
    ai = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    aj, aa = aj[keep], aa[keep]
    monkeypatch.setenv("B200_STREAM_THREADS", "128")   # what the plan picks for rows this long anyway
    A = pk.Csr(ai, aj, aa)
    monkeypatch.delenv("B200_STREAM_THREADS")
    info = A.info()
    assert 160 < info.index8_diagonals <= 256, info.index8_diagonals
    assert info.stream_tiles == 2                      # 128 rows per tile
    assert pk.KERNEL_NAMES[info.kernel_exact] == "stream"
    x, y0 = gen.uniform_pm1(m, 5), gen.uniform_pm1(m, 6)
    for mode in (pk.MODE_EXACT, pk.MODE_EXACT_FMA, pk.MODE_FAST):
        ref = oracle.matmult(ai, aj, aa, x, fma=(mode != pk.MODE_EXACT))
        assert np.array_equal(_run(pk, cuda, A, x, mode), ref)
    torch = cuda
    dx, db = torch.from_numpy(x).cuda(), torch.from_numpy(y0).cuda()
    out = torch.full((m,), float("nan"), dtype=torch.float64, device="cuda")
    A.residual(dx, db, out, pk.MODE_EXACT)
    assert np.array_equal(out.cpu().numpy(), oracle.residual(ai, aj, aa, x, y0))
    A.destroy()


@pytest.mark.parametrize("threads", ["128", "256"])
@pytest.mark.parametrize("index8", ["1", "0"])
def test_tiles_of_empty_rows_with_unaligned_code_offset(pk, cuda, monkeypatch, threads, index8):
    """Regression (round-1 review): a stream tile made only of empty rows whose first non-zero
    offset is 4 mod 16 copies no values and no codes; the producer must not make the stage's
    barrier wait for the 16 bytes of the rounded code range (the consumers would spin for ever).
    Rows 200..999 are empty and ai[200] = 596 = 37*16 + 4, so every tile inside that run, for both
    CTA sizes, is such a tile."""
    m, n = 3000, 3002
    lens = np.full(m, 3, np.int64)
    lens[:4] = 2
    lens[200:1000] = 0
    ai = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    assert ai[200] % 16 == 4 and ai[200] == ai[1000]
    aj = np.concatenate([np.arange(r, r + lens[r]) for r in range(m)]).astype(np.int32)
    aa = gen.uniform_pm1(len(aj), 17)
    monkeypatch.setenv("B200_STREAM_THREADS", threads)
    monkeypatch.setenv("B200_INDEX8", index8)
    A = pk.Csr(ai, aj, aa, n=n)
    info = A.info()
    assert pk.KERNEL_NAMES[info.kernel_exact] == "stream" and (info.index8_diagonals == 3) == (index8 == "1")
    x, y0 = gen.uniform_pm1(n, 5), gen.uniform_pm1(m, 6)
    for mode in (pk.MODE_EXACT, pk.MODE_EXACT_FMA):
        fma = mode != pk.MODE_EXACT
        assert np.array_equal(_run(pk, cuda, A, x, mode), oracle.matmult(ai, aj, aa, x, fma=fma))
        assert np.array_equal(_run(pk, cuda, A, x, mode, add=y0), oracle.matmultadd(ai, aj, aa, x, y0, fma=fma))
    A.destroy()


@pytest.mark.parametrize("name", ["powerlaw_20k", "long_rows", "random_ragged"])
def test_column_blocks_of_the_skewed_plan_keep_the_bits(pk, cuda, monkeypatch, name):
    """When x does not fit L2 the k_wmerge plan is cut into column blocks, y = A_0 x, y += A_1 x, ...;
    columns ascend inside a row, so that is still the reference's left-to-right sum.  Forced here on
    small matrices (B200_COLBLOCK_KB); the 10 M-row matrix of test_gpu_fullsize takes it by itself."""
    ai, aj, aa, n = CASES[name]
    m = len(ai) - 1
    monkeypatch.setenv("B200_COLBLOCK_KB", "16")
    A = pk.Csr(ai, aj, aa, n=n)
    monkeypatch.delenv("B200_COLBLOCK_KB")
    info = A.info()
    x, y0 = gen.uniform_pm1(n, 21), gen.uniform_pm1(m, 22)
    if pk.KERNEL_NAMES[info.kernel_exact] != "merge":
        A.set_kernel(pk.KERNEL_MERGE)       # (a plan that is not skewed enough: no blocks, plain k_wmerge)
    for mode, fma in ((pk.MODE_EXACT, False), (pk.MODE_EXACT_FMA, True), (pk.MODE_FAST, None)):
        y = _run(pk, cuda, A, x, mode)
        if fma is None:
            ref = oracle.matmult(ai, aj, aa, x)
            assert np.all(np.abs(y - ref) <= TOL * oracle.row_abs_sum(ai, aj, aa, x))
        else:
            assert np.array_equal(y, oracle.matmult(ai, aj, aa, x, fma=fma)), (name, mode)
            assert np.array_equal(_run(pk, cuda, A, x, mode, add=y0), oracle.matmultadd(ai, aj, aa, x, y0, fma=fma)), (name, mode)
    if pk.KERNEL_NAMES[info.kernel_exact] == "merge":
        assert info.device_bytes > 1.8 * (len(aj) * 12)          # the blocks are a second copy of aj / aa
    A.update_values(aa * 0.5)                                  # drops the blocks, same plan without them
    assert np.array_equal(_run(pk, cuda, A, x, pk.MODE_EXACT), oracle.matmult(ai, aj, aa * 0.5, x))
    A.destroy()


def test_compressed_index_plan_and_equivalence(pk, cuda, monkeypatch):
    """Stencil matrices stream 1-byte diagonal codes; results are the same bits as with int32
    column indices, and matrices with more than 256 diagonals keep int32."""
    p = oracle.poisson7(20)
    x = gen.uniform_pm1(8000, 3)
    ref = oracle.matmult(p["ai"], p["aj"], p["aa"], x)
    A8 = pk.Csr(p["ai"], p["aj"], p["aa"])
    assert A8.info().index8_diagonals == 7
    y8 = _run(pk, cuda, A8, x, pk.MODE_EXACT, kernel=pk.KERNEL_STREAM)
    monkeypatch.setenv("B200_INDEX8", "0")
    A32 = pk.Csr(p["ai"], p["aj"], p["aa"])
    assert A32.info().index8_diagonals == 0
    y32 = _run(pk, cuda, A32, x, pk.MODE_EXACT, kernel=pk.KERNEL_STREAM)
    assert np.array_equal(y8, ref) and np.array_equal(y32, ref)
    monkeypatch.delenv("B200_INDEX8")
    ai, aj, aa = gen.stencil27(12, seed=1)
    A = pk.Csr(ai, aj, aa)
    assert A.info().index8_diagonals == 27
    A.destroy()
    rng = np.random.default_rng(1)
    ai, aj, aa = gen.random_csr(3000, 3000, 8, rng)
    A = pk.Csr(ai, aj, aa)
    assert A.info().index8_diagonals == 0     # thousands of distinct diagonals
    xr = gen.uniform_pm1(3000, 4)
    assert np.array_equal(_run(pk, cuda, A, xr, pk.MODE_EXACT), oracle.matmult(ai, aj, aa, xr))
    A.destroy(); A8.destroy(); A32.destroy()


def test_edge_shapes(pk, cuda):
    """Zero rows, zero columns, one giant row, a diagonal, strongly non-square: every entry point."""
    torch = cuda
    rng = np.random.default_rng(8)
    shapes = {}
    shapes["zero_rows"] = (np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0), 5)
    n_big = 150_000
    cols = np.sort(rng.choice(n_big, size=100_000, replace=False)).astype(np.int32)
    shapes["one_giant_row"] = (np.array([0, 0, len(cols), len(cols)], np.int32), cols, rng.uniform(-1, 1, len(cols)), n_big)
    d = 5000
    shapes["diagonal"] = (np.arange(d + 1, dtype=np.int32), np.arange(d, dtype=np.int32), rng.uniform(-1, 1, d), d)
    ai, aj, aa = gen.random_csr(40, 9000, 30, rng)
    shapes["wide"] = (ai, aj, aa, 9000)
    ai, aj, aa = gen.random_csr(9000, 40, 8, rng)
    shapes["tall"] = (ai, aj, aa, 40)
    for name, (ai, aj, aa, n) in shapes.items():
        m = len(ai) - 1
        A = pk.Csr(ai, aj, aa, n=n)
        x, xt, y0 = gen.uniform_pm1(n, 1), gen.uniform_pm1(m, 2), gen.uniform_pm1(m, 3)
        dx, dxt = torch.from_numpy(x).cuda(), torch.from_numpy(xt).cuda()
        dy = torch.full((m,), float("nan"), dtype=torch.float64, device="cuda")
        dyt = torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
        for mode, fma in ((pk.MODE_EXACT, False), (pk.MODE_EXACT_FMA, True)):
            A.mult(dx, dy, mode)
            assert np.array_equal(dy.cpu().numpy(), oracle.matmult(ai, aj, aa, x, fma=fma)), name
            A.mult_add(dx, torch.from_numpy(y0).cuda(), dy, mode)
            assert np.array_equal(dy.cpu().numpy(), oracle.matmultadd(ai, aj, aa, x, y0, fma=fma)), name
            A.mult_transpose(dxt, dyt, mode)
            assert np.array_equal(dyt.cpu().numpy(), oracle.matmulttranspose(ai, aj, aa, xt, n, fma=fma)), name
        A.mult(dx, dy, pk.MODE_FAST)
        ref = oracle.matmult(ai, aj, aa, x)
        assert np.all(np.abs(dy.cpu().numpy() - ref) <= 1e-13 * oracle.row_abs_sum(ai, aj, aa, x)), name
        assert np.array_equal(A.mult_host(x, mode=pk.MODE_EXACT), ref), name
        A.destroy()


def test_argument_contract_on_device(pk, cuda):
    """MatMult requires x != y; bad modes and null vectors are errors, not crashes."""
    torch = cuda
    p = oracle.poisson7(6)
    A = pk.Csr(p["ai"], p["aj"], p["aa"])
    v = torch.zeros(A.m, dtype=torch.float64, device="cuda")
    w = torch.zeros(A.m, dtype=torch.float64, device="cuda")
    with pytest.raises(pk.B200Error) as e:
        A.mult(v, v)
    assert e.value.code == 60
    with pytest.raises(pk.B200Error):
        A.mult(v, w, mode=7)
    with pytest.raises(pk.B200Error):
        A.mult_add(v, w, v)          # z aliases x
    with pytest.raises(pk.B200Error):
        A.set_kernel(99)
    A.mult_add(v, w, w)              # z aliases y: allowed (MatMult_MPIAIJ's yy = yy + B lvec)
    A.destroy()


@pytest.mark.parametrize("name", ["poisson7_50", "stencil27_16", "powerlaw_20k", "random_ragged"])
def test_fused_residual_and_jacobi_sweep(pk, cuda, name):
    """r = b - A x and xnew = x + dinv.*(b - A x) in one pass: the same bits as PETSc's separate
    MatMult, VecAYPX, VecPointwiseMult, VecAXPY (oracle.residual / oracle.jacobi_sweep)."""
    torch = cuda
    ai, aj, aa, n = CASES[name]
    m = len(ai) - 1
    if m != n:   # the sweep needs a square operator; the residual does not
        x, b = gen.uniform_pm1(n, 1), gen.uniform_pm1(m, 2)
        A = pk.Csr(ai, aj, aa, n=n)
        r = torch.empty(m, dtype=torch.float64, device="cuda")
        A.residual(torch.from_numpy(x).cuda(), torch.from_numpy(b).cuda(), r, pk.MODE_EXACT)
        assert np.array_equal(r.cpu().numpy(), oracle.residual(ai, aj, aa, x, b))
        A.destroy()
        return
    x, b = gen.uniform_pm1(n, 1), gen.uniform_pm1(m, 2)
    dinv = 1.0 / (1.5 + gen.uniform_pm1(m, 3))
    A = pk.Csr(ai, aj, aa, n=n)
    dx, db, dd = (torch.from_numpy(v).cuda() for v in (x, b, dinv))
    out = torch.full((m,), float("nan"), dtype=torch.float64, device="cuda")
    kernels = [None, pk.KERNEL_ROW] + ([pk.KERNEL_STREAM] if A.info().stream_tiles else [])
    for k in kernels:
        if k is not None:
            A.set_kernel(k)
        A.residual(dx, db, out, pk.MODE_EXACT)
        assert np.array_equal(out.cpu().numpy(), oracle.residual(ai, aj, aa, x, b)), (name, k)
        A.jacobi_sweep(dx, db, dd, out, pk.MODE_EXACT)
        assert np.array_equal(out.cpu().numpy(), oracle.jacobi_sweep(ai, aj, aa, x, b, dinv)), (name, k)
        A.set_kernel(pk.KERNEL_AUTO)
    with pytest.raises(pk.B200Error):
        A.jacobi_sweep(dx, db, dd, dx, pk.MODE_EXACT)   # in place is an error: other rows still read x
    A.destroy()


def test_create_from_device_arrays(pk, cuda):
    torch = cuda
    p = oracle.poisson7(14)
    d_ai, d_aj, d_aa = (torch.from_numpy(p[k]).cuda() for k in ("ai", "aj", "aa"))
    A = pk.Csr.from_device(d_ai, d_aj, d_aa, 14 ** 3, 14 ** 3)
    assert A.nz == len(p["aj"]) and A.info().stream_tiles > 0 and A.info().index8_diagonals == 7
    x = gen.uniform_pm1(14 ** 3, 2)
    dy = torch.empty(14 ** 3, dtype=torch.float64, device="cuda")
    A.mult(torch.from_numpy(x).cuda(), dy, pk.MODE_EXACT)
    assert np.array_equal(dy.cpu().numpy(), oracle.matmult(p["ai"], p["aj"], p["aa"], x))
    A.destroy()


def test_transpose_host_and_device_builds_agree(pk, cuda, monkeypatch):
    """The device build (histogram + scan + stable radix sort) and the host counting sort give the
    same transpose, hence the same bits."""
    torch = cuda
    for name in ("poisson7_20", "powerlaw_20k", "random_ragged", "long_rows", "all_empty"):
        ai, aj, aa, n = CASES[name]
        m = len(ai) - 1
        x = gen.uniform_pm1(m, 5)
        ref = oracle.matmulttranspose(ai, aj, aa, x, n)
        outs = []
        for host in ("0", "1"):
            monkeypatch.setenv("B200_TRANSPOSE_HOST", host)
            A = pk.Csr(ai, aj, aa, n=n)
            dy = torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
            A.mult_transpose(torch.from_numpy(x).cuda(), dy, pk.MODE_EXACT)
            outs.append(dy.cpu().numpy())
            A.destroy()
        assert np.array_equal(outs[0], ref) and np.array_equal(outs[1], ref), name


def test_transpose_by_atomics(pk, cuda, monkeypatch):
    """FAST-mode A^T x without the explicit transpose copy (fp64 RED.ADD): within the column bound."""
    torch = cuda
    monkeypatch.setenv("B200_TRANSPOSE_ATOMIC", "1")
    ai, aj, aa, n = CASES["random_ragged"]
    m = len(ai) - 1
    A = pk.Csr(ai, aj, aa, n=n)
    x, z = gen.uniform_pm1(m, 5), gen.uniform_pm1(n, 6)
    dx, dz = torch.from_numpy(x).cuda(), torch.from_numpy(z).cuda()
    dy = torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
    absb = np.zeros(n)
    np.add.at(absb, aj, np.abs(aa * np.repeat(x, np.diff(ai))))
    A.mult_transpose(dx, dy, pk.MODE_FAST)
    assert not A.info().has_transpose
    assert np.all(np.abs(dy.cpu().numpy() - oracle.matmulttranspose(ai, aj, aa, x, n)) <= TOL * absb)
    A.mult_transpose_add(dx, dz, dy, pk.MODE_FAST)
    assert np.all(np.abs(dy.cpu().numpy() - oracle.matmulttransposeadd(ai, aj, aa, x, z, n)) <= TOL * (absb + np.abs(z)))
    A.mult_transpose(dx, dy, pk.MODE_EXACT)          # EXACT always goes through the explicit copy
    assert A.info().has_transpose
    assert np.array_equal(dy.cpu().numpy(), oracle.matmulttranspose(ai, aj, aa, x, n))
    A.destroy()


def test_vector_kernels_direct(pk, cuda):
    torch = cuda
    for n in (1, 31, 1000, 1_000_003):
        a, b = gen.uniform_pm1(n, 1), gen.uniform_pm1(n, 2)
        da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        out = torch.zeros(1, dtype=torch.float64, device="cuda")
        pk.vec_dot(da, db, out)
        assert abs(out.item() - float(np.dot(a, b))) <= 1e-12 * float(np.abs(a * b).sum()) + 1e-300
        first = out.item()
        pk.vec_dot(da, db, out)
        assert out.item() == first                      # deterministic reduction
        pk.vec_norm2(da, out)
        assert abs(out.item() - float(np.linalg.norm(a))) <= 1e-13 * float(np.linalg.norm(a)) + 1e-300
        pk.vec_norm_inf(da, out)
        assert out.item() == float(np.abs(a).max())
        pk.vec_sum(da, out)
        assert abs(out.item() - float(a.sum())) <= 1e-12 * float(np.abs(a).sum()) + 1e-300
        dw = torch.empty_like(da)
        pk.vec_pointwise_mult(dw, da, db)
        assert np.array_equal(dw.cpu().numpy(), a * b)
        pk.vec_copy(dw, da)
        pk.vec_axpy(dw, 0.5, db)
        assert np.allclose(dw.cpu().numpy(), a + 0.5 * b, rtol=0, atol=1e-15)
        pk.vec_aypx(dw, -1.0, db)
        assert np.allclose(dw.cpu().numpy(), b - (a + 0.5 * b), rtol=0, atol=1e-15)
        pk.vec_set(dw, 3.25)
        assert np.all(dw.cpu().numpy() == 3.25)


def test_registered_pageable_host_vectors(pk, cuda):
    """cudaHostRegister route for Vec arrays PETSc allocated itself (real-PETSc route)."""
    import ctypes as C
    p = oracle.poisson7(20)
    A = pk.Csr(p["ai"], p["aj"], p["aa"])
    x = np.ascontiguousarray(gen.uniform_pm1(A.n, 1))
    y = np.empty(A.m)
    pk.check(pk.lib.b200_host_register(x.ctypes.data_as(C.c_void_p), C.c_size_t(x.nbytes)))
    pk.check(pk.lib.b200_host_register(y.ctypes.data_as(C.c_void_p), C.c_size_t(y.nbytes)))
    A.mult_host(x, y, pk.MODE_EXACT)
    assert np.array_equal(y, oracle.matmult(p["ai"], p["aj"], p["aa"], x))
    pk.check(pk.lib.b200_host_unregister(x.ctypes.data_as(C.c_void_p)))
    pk.check(pk.lib.b200_host_unregister(y.ctypes.data_as(C.c_void_p)))
    A.destroy()


@pytest.mark.parametrize("name", ["poisson7_20", "stencil27_16", "powerlaw_20k", "random_ragged", "mostly_empty", "long_rows"])
def test_no_write_outside_the_output_vector(pk, cuda, name):
    """compute-sanitizer is closed on this pool: guard bands around y (and around lvec-sized scratch)
    must keep their sentinel through every kernel and epilogue."""
    torch = cuda
    ai, aj, aa, n = CASES[name]
    m = len(ai) - 1
    A = pk.Csr(ai, aj, aa, n=n)
    G = 4096
    sentinel = -7.25e300
    buf = torch.full((m + 2 * G,), sentinel, dtype=torch.float64, device="cuda")
    y = buf[G:G + m]
    bt = torch.full((n + 2 * G,), sentinel, dtype=torch.float64, device="cuda")
    yt = bt[G:G + n]
    x = torch.from_numpy(gen.uniform_pm1(n, 1)).cuda()
    xt = torch.from_numpy(gen.uniform_pm1(m, 2)).cuda()
    y0 = torch.from_numpy(gen.uniform_pm1(m, 3)).cuda()
    info = A.info()
    kernels = [pk.KERNEL_ROW, pk.KERNEL_VECTOR, pk.KERNEL_MERGE]
    if info.stream_tiles:
        kernels.append(pk.KERNEL_STREAM)
    if info.compressedrow_use:
        kernels.append(pk.KERNEL_CPROW)
    for k in kernels:
        A.set_kernel(k)
        mode = pk.MODE_FAST if k in (pk.KERNEL_VECTOR, pk.KERNEL_MERGE) else pk.MODE_EXACT
        A.mult(x, y, mode)
        A.mult_add(x, y0, y, mode)
    A.set_kernel(pk.KERNEL_AUTO)
    A.mult_transpose(xt, yt, pk.MODE_EXACT)
    if m == n:
        A.residual(x, y0, y, pk.MODE_EXACT)
        A.jacobi_sweep(x, y0, y0, y, pk.MODE_EXACT)
    torch.cuda.synchronize()
    for b, k in ((buf, m), (bt, n)):
        assert bool((b[:G] == sentinel).all()) and bool((b[G + k:] == sentinel).all()), name
    assert not bool((y == sentinel).any()) and not bool((yt == sentinel).any())
    A.destroy()


def test_gpu_results_match_committed_golden_checksums(pk, cuda):
    """The CUDA path against tests/golden/poisson7.json directly (no oracle call at run time):
    50^3 (BASELINE configs[0]) y = A x for the seeded x and for x = exact, y = A^T x, EXACT and
    EXACT_FMA; the power-law fixture."""
    import hashlib
    import json
    torch = cuda
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "poisson7.json")))

    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()

    case = [c for c in gold["seq"] if c["N"] == 50][0]
    g = pk.gen_poisson7(50, vectors=True)                       # product generator
    assert (sha(g["ai"]), sha(g["aj"]), sha(g["aa"]), sha(g["rhs"]), sha(g["exact"])) == tuple(case[k] for k in ("ai", "aj", "aa", "rhs", "exact"))
    A = pk.Csr(g["ai"], g["aj"], g["aa"])
    n = 50 ** 3
    x = torch.from_numpy(pk.gen_vector(n, 0xB200)).cuda()
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    A.mult(x, y, pk.MODE_EXACT)
    assert sha(y.cpu().numpy()) == case["y_rand"]
    A.mult(x, y, pk.MODE_EXACT_FMA)
    assert sha(y.cpu().numpy()) == case["y_rand_fma"]
    A.mult(torch.from_numpy(g["exact"]).cuda(), y, pk.MODE_EXACT)
    assert sha(y.cpu().numpy()) == case["y_exact"]
    A.mult_transpose(x, y, pk.MODE_EXACT)
    assert sha(y.cpu().numpy()) == case["yt_rand"]
    A.destroy()
    p = gold["synthetic"]["powerlaw_20000_3000"]
    ai, aj, aa = pk.gen_powerlaw(20000, lmax=3000)
    A = pk.Csr(ai, aj, aa)
    x = torch.from_numpy(pk.gen_vector(20000, 0xB200)).cuda()
    y = torch.empty(20000, dtype=torch.float64, device="cuda")
    A.mult(x, y, pk.MODE_EXACT)
    assert sha(y.cpu().numpy()) == p["y"]
    A.mult_transpose(x, y, pk.MODE_EXACT)
    assert sha(y.cpu().numpy()) == p["yt"]
    A.destroy()
