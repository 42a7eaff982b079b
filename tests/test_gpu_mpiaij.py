"""GPU parity of the MatMult_MPIAIJ path on ONE device: all ranks of a process grid live in this
process; every rank's push is launched before any rank's off-diagonal kernel waits (the deadlock-free
order on a single GPU), through the same kernels and windows the multi-process path uses."""
import numpy as np
import pytest

import gen
import oracle

pytestmark = pytest.mark.gpu


def setup(pk, N, size):
    ranks, gens = [], []
    for r in range(size):
        g = pk.gen_poisson7(N, size, r)
        M = pk.MpiAij(size, r, g["base"], g["ai"], g["aj"], g["aa"])
        ranks.append(M)
        gens.append(g)
    garrays = [M.garray() for M in ranks]
    for M in ranks:
        for q in range(size):
            M.set_peer_garray(q, garrays[q])
        M.upload()
    for M in ranks:
        for q in range(size):
            if q != M.rank:
                M.set_peer_window(q, ranks[q].window_ptr())
    return ranks, gens, garrays


@pytest.mark.parametrize("N,size", [(12, 2), (12, 4), (16, 8), (13, 8), (10, 3), (8, 1)])
@pytest.mark.parametrize("mode_name", ["exact", "fma"])
@pytest.mark.parametrize("fused", [False, True])
def test_mpiaij_matmult_bit_exact(pk, cuda, N, size, mode_name, fused):
    torch = cuda
    mode = pk.MODE_EXACT if mode_name == "exact" else pk.MODE_EXACT_FMA
    fma = mode_name == "fma"
    ranks, gens, garrays = setup(pk, N, size)
    n = N ** 3
    base = gens[0]["base"]
    for it in range(3):  # three MatMults: both lvec buffers and the sequence flags are exercised
        xg = gen.uniform_pm1(n, seed=100 + it)
        xs = [torch.from_numpy(xg[base[r]:base[r + 1]].copy()).cuda() for r in range(size)]
        ys = [torch.full((ranks[r].nloc,), float("nan"), dtype=torch.float64, device="cuda") for r in range(size)]
        for r, M in enumerate(ranks):
            M.mult_begin(xs[r])
        for r, M in enumerate(ranks):
            if fused:   # A x and B lvec in one launch (the kernel b200_mpiaij_mult uses)
                M.mult_finish(xs[r], ys[r], mode)
            else:
                M.mult_local(xs[r], ys[r], mode)
                M.mult_end(ys[r], mode)
        torch.cuda.synchronize()
        for r, M in enumerate(ranks):
            M.check()
            Ai, Aj, Aa = M.block(0)
            Bi, Bj, Ba = M.block(1)
            ref = oracle.matmult(Ai, Aj, Aa, xg[base[r]:base[r + 1]], fma=fma)
            ref = oracle.matmultadd(Bi, Bj, Ba, xg[garrays[r]], ref, fma=fma) if M.nghost else ref
            assert np.array_equal(ys[r].cpu().numpy(), ref), (N, size, r, it)
    for M in ranks:
        M.destroy()


def test_pack_and_caller_owned_transport(pk, cuda):
    """The NCCL-transport variant: pack -> (copy stands in for send/recv) -> y += B lvec."""
    torch = cuda
    N, size = 12, 4
    ranks, gens, garrays = setup(pk, N, size)
    base = gens[0]["base"]
    xg = gen.uniform_pm1(N ** 3, seed=7)
    xs = [torch.from_numpy(xg[base[r]:base[r + 1]].copy()).cuda() for r in range(size)]
    lvecs = [torch.zeros(max(M.nghost, 1), dtype=torch.float64, device="cuda") for M in ranks]
    for r, M in enumerate(ranks):
        for q in range(size):
            idx, off = M.send_list(q)
            if len(idx) == 0:
                continue
            buf = torch.empty(len(idx), dtype=torch.float64, device="cuda")
            M.pack(q, xs[r], buf)
            lvecs[q][off:off + len(idx)] = buf
    for r, M in enumerate(ranks):
        assert np.array_equal(lvecs[r].cpu().numpy()[:M.nghost], xg[garrays[r]])
        y = torch.empty(M.nloc, dtype=torch.float64, device="cuda")
        M.mult_local(xs[r], y, pk.MODE_EXACT)
        M.mult_add_ghost(lvecs[r], y, pk.MODE_EXACT)
        Ai, Aj, Aa = M.block(0)
        Bi, Bj, Ba = M.block(1)
        ref = oracle.matmultadd(Bi, Bj, Ba, xg[garrays[r]], oracle.matmult(Ai, Aj, Aa, xg[base[r]:base[r + 1]]))
        assert np.array_equal(y.cpu().numpy(), ref)
    for M in ranks:
        M.destroy()


def test_flag_timeout_is_reported_not_hung(pk, cuda, monkeypatch):
    """A missing peer push must end in B200_ERR_TIMEOUT, never in a hung GPU."""
    torch = cuda
    monkeypatch.setenv("B200_MPIAIJ_TIMEOUT_MS", "50")
    ranks, gens, garrays = setup(pk, 8, 2)
    M = ranks[0]
    x = torch.zeros(M.nloc, dtype=torch.float64, device="cuda")
    y = torch.zeros(M.nloc, dtype=torch.float64, device="cuda")
    M.mult_begin(x)          # rank 1 never pushes
    M.mult_local(x, y, pk.MODE_EXACT)
    M.mult_end(y, pk.MODE_EXACT)
    torch.cuda.synchronize()
    with pytest.raises(pk.B200Error) as e:
        M.check()
    assert e.value.code == 77
    for m in ranks:
        m.destroy()


def _torchrun(nproc, script, *args, timeout=600):
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(root, "tests", script), *[str(a) for a in args]]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    import json
    return json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])


def test_two_gpus_distributed_cg_and_halo_stress(pk, cuda):
    """Two real ranks (torchrun, one process per GPU, NVLink peer windows).
    1. KSPCG + PCJACOBI (fused MatMult_MPIAIJ with (p, A p) folded in + peer-window all-reduces, no host
       round trip): same iteration count on every rank and equal to the single-rank oracle CG's, error O(h^2).
    2. Halo stress (tests/mpiaij_stress_worker.py): thousands of unsynchronised MatMults with a
       different x each, skewed ranks, every result checked; the host-vector pipeline against the oracle.
    The 8-GPU runs of the same workers are logged under profiles/."""
    torch = cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    N = 24
    r = _torchrun(2, "mpiaij_cg_worker.py", N, "1e-10", timeout=300)
    p = oracle.poisson7(N)
    _, its, _ = oracle.cg_jacobi(p["ai"], p["aj"], p["aa"], p["rhs"], rtol=1e-10, atol=1e-50, max_it=20000)
    assert r["allreduce_ok"] and r["reason"] > 0
    assert len(set(r["its_all"])) == 1 and r["its"] == its, (r, its)
    assert r["linf_err"] < 0.02
    s = _torchrun(2, "mpiaij_stress_worker.py", 40, 3000)
    assert s["first_result_equals_oracle"] and s["mismatches"] == 0, s


def test_host_vector_pipeline_on_one_rank(pk, cuda, monkeypatch):
    """b200_mpiaij_mult_host on a one-rank 'partition' (no ghosts): the row-blocked pipeline of uploads,
    block kernels and downloads gives the bits of the oracle (several blocks forced on a small grid)."""
    monkeypatch.setenv("B200_HOST_BLOCK_ROWS", "8192")
    g = pk.gen_poisson7(40, 1, 0)
    M = pk.MpiAij(1, 0, g["base"], g["ai"], g["aj"], g["aa"])
    M.upload()
    x = gen.uniform_pm1(M.nloc, 9)
    hx, hy = pk.PinnedArray(M.nloc), pk.PinnedArray(M.nloc)
    hx.array[:] = x
    ref = oracle.matmult(g["ai"], g["aj"], g["aa"], x)
    for mode, fma in ((pk.MODE_EXACT, False), (pk.MODE_EXACT_FMA, True)):
        hy.array[:] = np.nan
        M.mult_host(hx.array, hy.array, mode)
        assert np.array_equal(hy.array, oracle.matmult(g["ai"], g["aj"], g["aa"], x, fma=fma))
    monkeypatch.setenv("B200_MPIAIJ_HOST_PIPELINE", "0")
    hy.array[:] = np.nan
    M.mult_host(hx.array, hy.array, pk.MODE_EXACT)
    assert np.array_equal(hy.array, ref)
    M.destroy()
