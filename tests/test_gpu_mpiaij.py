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


def test_distributed_cg_two_gpus(pk, cuda):
    """KSPCG + PCJACOBI over two ranks (fused MatMult_MPIAIJ + peer-window all-reduce): same
    iteration count on every rank, within one of the single-rank oracle CG, error at O(h^2)."""
    import json
    import os
    import subprocess
    import sys
    torch = cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    N = 24
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(root, "tests", "mpiaij_cg_worker.py"), str(N), "1e-10"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    p = oracle.poisson7(N)
    _, its, _ = oracle.cg_jacobi(p["ai"], p["aj"], p["aa"], p["rhs"], rtol=1e-10, atol=1e-50, max_it=20000)
    assert r["allreduce_ok"] and r["reason"] > 0
    assert len(set(r["its_all"])) == 1 and r["its"] == its, (r, its)
    assert r["linf_err"] < 0.02


def test_halo_stress_two_gpus(cuda):
    """Thousands of unsynchronised MatMults with a different x each, skewed ranks, every result
    checked (tests/mpiaij_stress_worker.py); the 8-GPU run of the same worker is logged under profiles/."""
    import json
    import os
    import subprocess
    import sys
    torch = cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200), os.path.join(root, "tests", "mpiaij_stress_worker.py"), "40", "3000"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    r = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert r["first_result_equals_oracle"] and r["mismatches"] == 0, r
