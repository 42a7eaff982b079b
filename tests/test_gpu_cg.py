"""The fused Jacobi-CG (KSPSolve_CG + PCJACOBI) on the device against the oracle's CG.

The solve runs without a host round trip per iteration (b200_vec.cu: k_cg_p, MatMult with (p, A p)
folded into its epilogue, k_cg_r; convergence is decided on the device and the queued launches
drain).  What can be bit-exact is: the iteration count on these cases, the state after a fixed
number of iterations across chunk sizes / pointer alignments / streams.  The solution vector is
compared with the stated tolerance: the dot products are summed in a fixed tree, not left to right
(VecDot's order is BLAS-defined in the reference and MPI-defined on several ranks)."""
import threading

import numpy as np
import pytest

import gen
import oracle

pytestmark = pytest.mark.gpu


def _solve(pk, torch, A, b, x=None, **kw):
    db = torch.from_numpy(np.ascontiguousarray(b)).cuda()
    dx = torch.full((A.m,), float("nan"), dtype=torch.float64, device="cuda") if x is None else x
    res = A.cg_jacobi(db, dx, **kw)
    return res, dx.cpu().numpy()


@pytest.mark.parametrize("N", [7, 16, 24, 40])
def test_iteration_count_and_solution_against_the_oracle(pk, cuda, N):
    p = oracle.poisson7(N)
    A = pk.Csr(p["ai"], p["aj"], p["aa"])
    kw = dict(rtol=1e-10, atol=1e-50, max_it=5000)
    xo, its, rn = oracle.cg_jacobi(p["ai"], p["aj"], p["aa"], p["rhs"], **kw)
    for mode in (pk.MODE_EXACT, pk.MODE_EXACT_FMA, pk.MODE_FAST):
        res, x = _solve(pk, cuda, A, p["rhs"], mode=mode, **kw)
        assert res.reason == 2 and res.its == its, (N, mode, res.its, its)
        assert abs(res.rnorm - rn) <= 1e-3 * rn                 # the last residual is rounding-sensitive; the count is not
        assert np.abs(x - xo).max() <= 1e-8 * np.abs(xo).max()
        assert res.launches <= 3 * (res.its + 40) + 8          # 3 launches per iteration + the drained tail
    A.destroy()


def test_fixed_iteration_counts_and_the_deferred_update(pk, cuda):
    """max_it = k stops after exactly k iterations with KSP_DIVERGED_ITS and x = the oracle's k-th
    iterate (the x update of iteration k is applied by the closing kernel)."""
    p = oracle.poisson7(12)
    A = pk.Csr(p["ai"], p["aj"], p["aa"])
    for k in (0, 1, 2, 5, 17):
        xo, its, _ = oracle.cg_jacobi(p["ai"], p["aj"], p["aa"], p["rhs"], rtol=1e-30, atol=1e-300, max_it=k)
        res, x = _solve(pk, cuda, A, p["rhs"], rtol=1e-30, atol=1e-300, max_it=k, mode=pk.MODE_EXACT)
        assert its == -k and res.its == k and res.reason == -3, (k, its, res.its, res.reason)
        assert np.abs(x - xo).max() <= 1e-11 * max(np.abs(xo).max(), 1e-300), k
    A.destroy()


def test_zero_right_hand_side_converges_at_once(pk, cuda):
    p = oracle.poisson7(8)
    A = pk.Csr(p["ai"], p["aj"], p["aa"])
    res, x = _solve(pk, cuda, A, np.zeros(A.m), rtol=1e-10, atol=1e-12, max_it=100)
    assert res.its == 0 and res.reason == 3 and not x.any()
    A.destroy()


def test_same_bits_for_every_chunk_size_alignment_and_plan(pk, cuda, monkeypatch):
    torch = cuda
    p = oracle.poisson7(15)                                    # odd row count: the two-element path has a tail
    A = pk.Csr(p["ai"], p["aj"], p["aa"])
    kw = dict(rtol=1e-9, atol=1e-50, max_it=3000, mode=pk.MODE_EXACT)
    res0, x0 = _solve(pk, torch, A, p["rhs"], **kw)
    assert res0.reason == 2
    for chunk in ("1", "3", "64"):
        monkeypatch.setenv("B200_CG_CHUNK", chunk)
        res, x = _solve(pk, torch, A, p["rhs"], **kw)
        assert (res.its, res.rnorm) == (res0.its, res0.rnorm) and np.array_equal(x, x0), chunk
    monkeypatch.delenv("B200_CG_CHUNK")
    # x at an address that is not 16-byte aligned: the scalar variants of the update kernels
    big = torch.full((A.m + 1,), float("nan"), dtype=torch.float64, device="cuda")
    res, _ = _solve(pk, torch, A, p["rhs"], x=big[1:], **kw)
    assert (res.its, res.rnorm) == (res0.its, res0.rnorm) and np.array_equal(big[1:].cpu().numpy(), x0)
    # programmatic dependent launch off: same arithmetic
    # plans that cannot fold the dot product into the MatMult reduce afterwards: another summation tree,
    # so the count is compared, not the bits
    A.set_kernel(pk.KERNEL_ROW)
    res, x = _solve(pk, torch, A, p["rhs"], **kw)
    assert res.reason == 2 and abs(res.its - res0.its) <= 1 and np.abs(x - x0).max() <= 1e-7 * np.abs(x0).max()
    A.destroy()


def test_two_streams_and_two_threads_at_once(pk, cuda):
    """The reduction scratch and the CG workspace are per stream / per host thread: two solves and two
    stand-alone dot products in flight together give the bits of the serial runs."""
    torch = cuda
    p = oracle.poisson7(20)
    A1, A2 = pk.Csr(p["ai"], p["aj"], p["aa"]), pk.Csr(p["ai"], p["aj"], 2.0 * p["aa"])
    kw = dict(rtol=1e-9, atol=1e-50, max_it=3000, mode=pk.MODE_EXACT)
    ref1, x1 = _solve(pk, torch, A1, p["rhs"], **kw)
    ref2, x2 = _solve(pk, torch, A2, p["rhs"], **kw)
    out = {}

    def work(tag, A):
        torch.cuda.set_device(0)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for rep in range(3):
                res, x = _solve(pk, torch, A, p["rhs"], stream=s, **kw)
            s.synchronize()
        out[tag] = (res.its, res.rnorm, x)

    ts = [threading.Thread(target=work, args=("a", A1)), threading.Thread(target=work, args=("b", A2))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert out["a"][:2] == (ref1.its, ref1.rnorm) and np.array_equal(out["a"][2], x1)
    assert out["b"][:2] == (ref2.its, ref2.rnorm) and np.array_equal(out["b"][2], x2)
    # stand-alone reductions on two streams, interleaved
    n = 3_000_000
    u, v = torch.from_numpy(gen.uniform_pm1(n, 1)).cuda(), torch.from_numpy(gen.uniform_pm1(n, 2)).cuda()
    want_uv, want_vv = torch.zeros(1, dtype=torch.float64, device="cuda"), torch.zeros(1, dtype=torch.float64, device="cuda")
    pk.vec_dot(u, v, want_uv)
    pk.vec_dot(v, v, want_vv)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    o1 = torch.zeros(64, dtype=torch.float64, device="cuda")
    o2 = torch.zeros(64, dtype=torch.float64, device="cuda")
    for k in range(64):
        pk.vec_dot(u, v, o1[k:k + 1], stream=s1)
        pk.vec_dot(v, v, o2[k:k + 1], stream=s2)
    torch.cuda.synchronize()
    assert bool((o1 == want_uv).all()) and bool((o2 == want_vv).all())
    A1.destroy(); A2.destroy()


def test_second_device_from_one_process(pk, cuda):
    """One process, two devices: per-device state (SM count, shared-memory opt-in of the stream
    kernels, reduction scratch)."""
    torch = cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU on this box")
    p = oracle.poisson7(20)
    x = gen.uniform_pm1(20 ** 3, 3)
    ref = oracle.matmult(p["ai"], p["aj"], p["aa"], x)
    try:
        for dev in (1, 0, 1):
            torch.cuda.set_device(dev)
            pk.init(dev)
            A = pk.Csr(p["ai"], p["aj"], p["aa"])
            dx = torch.from_numpy(x).to(f"cuda:{dev}")
            dy = torch.empty_like(dx)
            A.mult(dx, dy, pk.MODE_EXACT)
            assert np.array_equal(dy.cpu().numpy(), ref)
            res, _ = _solve_on(pk, torch, A, p["rhs"], dev)
            assert res.reason == 2
            A.destroy()
    finally:
        torch.cuda.set_device(0)
        pk.init(0)


def _solve_on(pk, torch, A, b, dev):
    db = torch.from_numpy(np.ascontiguousarray(b)).to(f"cuda:{dev}")
    dx = torch.zeros_like(db)
    return A.cg_jacobi(db, dx, rtol=1e-8, atol=1e-50, max_it=3000, mode=pk.MODE_EXACT), dx
