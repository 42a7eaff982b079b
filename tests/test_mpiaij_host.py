"""Host-side integer work of the MatMult_MPIAIJ path (A/B split, garray, scatter lists) against
the oracle, in-process and across two gloo ranks.  Bit-exact bar."""
import os
import sys

import numpy as np
import pytest

import oracle


def build_rank(pk, N, size, rank):
    g = pk.gen_poisson7(N, size, rank)
    return pk.MpiAij(size, rank, g["base"], g["ai"], g["aj"], g["aa"]), g


def oracle_rank(N, size, rank):
    p = oracle.poisson7(N, size=size, rank=rank)
    (Ai, Aj, Aa), (Bi, Bj, Ba) = oracle.mpiaij_split(p["ai"], p["aj"], p["aa"], p["rstart"], p["rend"])
    Bjc, garray = oracle.mpiaij_setup_multiply(Bj)
    return dict(A=(Ai, Aj, Aa), B=(Bi, Bjc, Ba), garray=garray, p=p)


@pytest.mark.parametrize("N,size", [(6, 1), (10, 2), (12, 4), (12, 8), (13, 8), (10, 3), (9, 6)])
def test_split_garray_and_scatter_lists(pk, N, size):
    base = oracle.dmda_bases(N, N, N, size)
    ranks = [build_rank(pk, N, size, r)[0] for r in range(size)]
    refs = [oracle_rank(N, size, r) for r in range(size)]
    for r, (M, o) in enumerate(zip(ranks, refs)):
        for which, key in ((0, "A"), (1, "B")):
            ai, aj, aa = M.block(which)
            assert np.array_equal(ai, o[key][0]) and np.array_equal(aj, o[key][1]) and np.array_equal(aa, o[key][2])
        assert np.array_equal(M.garray(), o["garray"])
        assert np.array_equal(M.recv_offsets(), oracle.scatter_recv_offsets(base, o["garray"]))
        assert M.brows == int((np.diff(o["B"][0]) > 0).sum())
    # scatter lists: what rank r sends to q is q's garray run inside r's rows, in garray order
    for r, M in enumerate(ranks):
        for q in range(size):
            M.set_peer_garray(q, refs[q]["garray"])
        for q in range(size):
            idx, off = M.send_list(q)
            if q == r:
                assert len(idx) == 0
                continue
            roff = oracle.scatter_recv_offsets(base, refs[q]["garray"])
            seg = refs[q]["garray"][roff[r]:roff[r + 1]]
            assert np.array_equal(idx, seg - base[r])
            if len(seg):
                assert off == roff[r]
    for M in ranks:
        M.destroy()


def test_survey_halo_sizes_scaled(pk):
    """2x2x2 grid -> three face neighbours; per-rank halo = 3 faces."""
    N, size = 20, 8
    M, _ = build_rank(pk, N, size, 0)
    h = N // 2
    assert (M.nloc, M.nghost, M.bnnz, M.nsrc) == (h ** 3, 3 * h * h, 3 * h * h, 3)
    assert M.annz == 7 * h ** 3 - 6 * h ** 2
    assert M.brows == 3 * h * h - 3 * h + 1
    M.destroy()
    M2, _ = build_rank(pk, N, 2, 0)
    assert (M2.nghost, M2.nsrc) == (N * N, 1)
    M2.destroy()


def _gloo_worker(rank, world, N, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import petsc_openacc_b200 as pk
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        M, g = build_rank(pk, N, world, rank)
        garrays = [None] * world
        dist.all_gather_object(garrays, M.garray())
        for peer in range(world):
            M.set_peer_garray(peer, garrays[peer])
        out = {p: M.send_list(p) for p in range(world)}
        q.put((rank, M.garray(), out, M.recv_offsets()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_exchange_of_scatter_lists(pk):
    import torch.multiprocessing as mp
    N, world = 10, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, N, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, garray, sends, roff = q.get(timeout=180)
        res[r] = (garray, sends, roff)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    base = oracle.dmda_bases(N, N, N, world)
    refs = [oracle_rank(N, world, r) for r in range(world)]
    for r in range(world):
        garray, sends, roff = res[r]
        assert np.array_equal(garray, refs[r]["garray"])
        peer = 1 - r
        idx, off = sends[peer]
        proff = oracle.scatter_recv_offsets(base, refs[peer]["garray"])
        assert np.array_equal(idx, refs[peer]["garray"][proff[r]:proff[r + 1]] - base[r])
        assert off == proff[r] and len(idx) == N * N


def test_unsymmetric_communication_pattern_is_reported(pk):
    """The push exchange alternates two receive buffers and needs every destination to be a source
    too.  An upper-triangular coupling (rank 0 reads rank 1's rows, rank 1 reads nothing of rank 0's)
    is reported by b200_mpiaij_pattern_symmetric -- and refused by the MatMult entries -- while the
    reference problem at every rank count is symmetric."""
    base = np.array([0, 3, 6], np.int32)
    # rank 0: rows 0..2 with a column owned by rank 1; rank 1: rows 3..5, own columns only
    r0 = pk.MpiAij(2, 0, base, [0, 2, 3, 5], [0, 4, 1, 2, 5], [1.0, 2.0, 3.0, 4.0, 5.0])
    r1 = pk.MpiAij(2, 1, base, [0, 1, 2, 3], [3, 4, 5], [1.0, 1.0, 1.0])
    g0, g1 = r0.garray(), r1.garray()
    assert list(g0) == [4, 5] and len(g1) == 0
    r0.set_peer_garray(1, g1)
    r1.set_peer_garray(0, g0)
    assert r0.pattern_symmetric() == (True, -1)          # rank 0 only receives
    ok, peer = r1.pattern_symmetric()                    # rank 1 sends to 0 and receives nothing from it
    assert not ok and peer == 0
    assert "symmetric communication pattern" in pk.lib.b200_last_error().decode()
    r0.destroy(); r1.destroy()
    for size in (2, 4, 8):
        ms = []
        for rank in range(size):
            p = oracle.poisson7(8, size=size, rank=rank)
            ms.append(pk.MpiAij(size, rank, oracle.dmda_bases(8, 8, 8, size), p["ai"], p["aj"], p["aa"]))
        for a in ms:
            for b in ms:
                if a is not b:
                    a.set_peer_garray(b.rank, b.garray())
        assert all(m.pattern_symmetric()[0] for m in ms)
        [m.destroy() for m in ms]


def test_fused_launch_tile_schedule(pk):
    """Tile -> CTA of the fused MatMult_MPIAIJ launch: every tile exactly once, every CTA at least one,
    plain round-robin when nothing is charged, fewer tiles for the CTAs that close a ghost-heavy tile or
    carry a push block, and the light tiles still dealt in index order (lockstep sweep)."""
    grid = 740
    none = np.zeros(13184, np.int32)
    assert np.array_equal(pk.mpiaij_tile_schedule(none, grid), np.arange(13184) % grid)
    # rank 0 of the 8-rank 300^3 decomposition: the last 88 tiles are a contiguous face (256 ghost rows each),
    # a y-face tile every 88 tiles, x-face rows everywhere
    g = np.full(13184, 2, np.int32)
    g[::88] = 150
    g[-88:] = 256
    cta = pk.mpiaij_tile_schedule(g, grid, npush=32, push_charge=6.0)
    assert cta.min() == 0 and cta.max() == grid - 1
    count = np.bincount(cta, minlength=grid)
    assert count.min() >= 1 and count.sum() == len(g)
    load = np.bincount(cta, weights=1.0 + g / 128.0, minlength=grid)
    load[:32] += 6.0
    assert load.max() - load.min() <= 2.2                       # balanced to within one tile + its charge
    assert count[:32].max() <= count[32:].max() - 5                # the push CTAs stream ~6 tiles less
    heavy_owner = cta[-88:]
    assert len(set(heavy_owner)) == 88                           # one ghost face tile per CTA at most
    assert count[heavy_owner].max() <= count.max() - 2           # and those CTAs sit out ~3 rounds
    # lockstep: the light tiles of a CTA are spread over the whole matrix, like round-robin's
    light = np.nonzero(g < 64)[0]
    for b in (40, 400, 739):
        mine = light[cta[light] == b]
        assert np.all(np.diff(mine) > 0) and np.diff(mine).max() < 3 * grid
    # tiny matrices: the charges must not starve a CTA
    small = np.array([300, 0, 0, 0, 0, 0, 0], np.int32)
    assert sorted(pk.mpiaij_tile_schedule(small, 7, npush=3, push_charge=6.0)) == list(range(7))
