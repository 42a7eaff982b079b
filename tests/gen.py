"""Synthetic inputs for tests and sweeps (numpy, vectorised).  SURVEY 8(d) definitions.

These are INPUT generators, independent of both the oracle and the product; the 7-point matrix
here is cross-checked against the oracle's restatement of src/helper.cpp in tests/test_oracle.py.
"""
import numpy as np


def splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform_pm1(n, seed=0xB200, offset=0):
    """x[i] = uniform[-1,1) from splitmix64(seed ^ i)"""
    with np.errstate(over="ignore"):
        i = np.arange(offset, offset + n, dtype=np.uint64)
        z = splitmix64(np.uint64(seed) ^ i)
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0


def poisson7_natural(N, refpoint=True):
    """7-point Neumann Laplacian of src/helper.cpp on one rank (natural ordering), vectorised."""
    n = N ** 3
    idx = np.arange(n, dtype=np.int64)
    i, j, k = idx % N, (idx // N) % N, idx // (N * N)
    dx = 1.0 / N
    v = 1.0 / (dx * dx)
    # neighbour order of ascending column: -N^2, -N, -1, 0, +1, +N, +N^2
    present = [k > 0, j > 0, i > 0, np.ones(n, bool), i < N - 1, j < N - 1, k < N - 1]
    offs = [-N * N, -N, -1, 0, 1, N, N * N]
    cnt = sum(p.astype(np.int32) for p in present)
    ai = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=ai[1:])
    nz = int(ai[-1])
    aj = np.empty(nz, dtype=np.int32)
    aa = np.empty(nz, dtype=np.float64)
    # diagonal: 0 - v - v ... in the order idx 1..6 = (i-1, i+1, j-1, j+1, k-1, k+1)
    diag = np.zeros(n)
    for p in (present[2], present[4], present[1], present[5], present[0], present[6]):
        diag = np.where(p, diag - v, diag)
    pos = ai[:-1].copy()
    for s, (p, o) in enumerate(zip(present, offs)):
        w = pos[p]
        aj[w] = (idx[p] + o).astype(np.int32)
        aa[w] = diag[p] if o == 0 else v
        pos[p] += 1
    ai = ai.astype(np.int32)
    if refpoint:
        scale = float(np.add.reduce(diag.tolist())) if n <= 200000 else None
        if scale is None:
            s = 0.0
            for d in diag:  # sequential like VecSum; slow path only for big N
                s += d
            scale = s
        scale /= float(n)
        # row 0 -> zeros, diag = scale; column 0 entries of other rows -> 0
        aa[ai[0]:ai[1]] = 0.0
        aa[ai[0]] = scale  # column 0 is the first entry of row 0
        col0 = np.nonzero(aj == 0)[0]
        col0 = col0[col0 >= ai[1]]
        aa[col0] = 0.0
    return ai, aj, aa


def stencil27(N, seed=None):
    """27-point box stencil, non-periodic: off-diagonal -1, diagonal = number of neighbours
    (or seeded uniform values when seed is given).  nnz = (3N-2)^3."""
    n = N ** 3
    idx = np.arange(n, dtype=np.int64)
    i, j, k = idx % N, (idx // N) % N, idx // (N * N)
    masks, offs = [], []
    for dk in (-1, 0, 1):
        for dj in (-1, 0, 1):
            for di in (-1, 0, 1):
                m = ((i + di >= 0) & (i + di < N) & (j + dj >= 0) & (j + dj < N) &
                     (k + dk >= 0) & (k + dk < N))
                masks.append(m)
                offs.append(di + dj * N + dk * N * N)
    cnt = sum(m.astype(np.int32) for m in masks)
    ai = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=ai[1:])
    nz = int(ai[-1])
    aj = np.empty(nz, dtype=np.int32)
    aa = np.empty(nz, dtype=np.float64)
    pos = ai[:-1].copy()
    for m, o in zip(masks, offs):
        w = pos[m]
        aj[w] = (idx[m] + o).astype(np.int32)
        aa[w] = (cnt[m] - 1).astype(np.float64) if o == 0 else -1.0
        pos[m] += 1
    if seed is not None:
        aa = uniform_pm1(nz, seed)
    return ai.astype(np.int32), aj, aa


def powerlaw(m, n=None, alpha=2.0, lmax=10000, seed=0x5EED):
    """Irregular CSR: row length L_i = min(lmax, max(1, floor(u_i^(-1/(alpha-1))))), columns =
    sorted unique splitmix64 draws mod n, values uniform [-1,1)."""
    n = m if n is None else n
    with np.errstate(over="ignore"):
        u = (splitmix64(np.uint64(seed) ^ np.arange(m, dtype=np.uint64)) >> np.uint64(11)).astype(
            np.float64) / 9007199254740992.0
    u = np.maximum(u, 1e-300)
    L = np.minimum(lmax, np.maximum(1, np.floor(u ** (-1.0 / (alpha - 1.0))))).astype(np.int64)
    L = np.minimum(L, n)
    start = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(L, out=start[1:])
    tot = int(start[-1])
    row = np.repeat(np.arange(m, dtype=np.int64), L)
    kk = np.arange(tot, dtype=np.int64) - start[row]
    with np.errstate(over="ignore"):
        h = splitmix64((np.uint64(seed) * np.uint64(0x100000001B3)) ^
                       (row.astype(np.uint64) << np.uint64(20)) ^ kk.astype(np.uint64))
    col = (h % np.uint64(n)).astype(np.int64)
    # sort by (row, col) and drop duplicates inside a row
    key = row * n + col
    key = np.unique(key)
    row2 = key // n
    aj = (key % n).astype(np.int32)
    cnt = np.bincount(row2, minlength=m)
    ai = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(cnt, out=ai[1:])
    aa = uniform_pm1(len(aj), seed ^ 0xABCDEF)
    return ai.astype(np.int32), aj, aa


def random_csr(m, n, density_rows, rng, empty_frac=0.0):
    """Small random CSR with sorted unique columns; some rows empty."""
    lens = rng.integers(0, density_rows + 1, size=m)
    if empty_frac > 0:
        lens[rng.random(m) < empty_frac] = 0
    lens = np.minimum(lens, n)
    ai = np.zeros(m + 1, dtype=np.int32)
    np.cumsum(lens, out=ai[1:])
    aj = np.empty(int(ai[-1]), dtype=np.int32)
    for r in range(m):
        aj[ai[r]:ai[r + 1]] = np.sort(rng.choice(n, size=lens[r], replace=False))
    aa = rng.uniform(-1, 1, size=len(aj))
    return ai, aj, aa
