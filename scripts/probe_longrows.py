"""Regular matrices with LONG rows (the shape of the coarse multigrid operators: 32 and ~130 entries
per row, int32 indices): the stream kernel (thread per row out of shared memory) against the
warp-granular exact-order kernel (k_wmerge).  Usage: python scripts/probe_longrows.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import petsc_openacc_b200 as pk


def box_stencil(N, h):
    """(2h+1)^3-point box stencil on an N^3 grid, non-periodic, ascending columns, values uniform."""
    n = N ** 3
    idx = np.arange(n, dtype=np.int64)
    i, j, k = idx % N, (idx // N) % N, idx // (N * N)
    masks, offs = [], []
    for dk in range(-h, h + 1):
        for dj in range(-h, h + 1):
            for di in range(-h, h + 1):
                masks.append((i + di >= 0) & (i + di < N) & (j + dj >= 0) & (j + dj < N) & (k + dk >= 0) & (k + dk < N))
                offs.append(di + dj * N + dk * N * N)
    cnt = sum(m.astype(np.int32) for m in masks)
    ai = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=ai[1:])
    aj = np.empty(int(ai[-1]), dtype=np.int32)
    pos = ai[:-1].copy()
    for m, o in zip(masks, offs):
        aj[pos[m]] = (idx[m] + o).astype(np.int32)
        pos[m] += 1
    return ai.astype(np.int32), aj, pk.gen_vector(len(aj), 5)


pk.init(0)
os.environ["B200_INDEX8"] = os.environ.get("B200_INDEX8", "0")
for name, N, h in (("27-point 150^3 (level-1 like: 27/row)", 150, 1), ("125-point 70^3 (level-2 like: ~115/row)", 70, 2)):
    ai, aj, aa = box_stencil(N, h)
    m, nz = len(ai) - 1, len(aj)
    A = pk.Csr(ai, aj, aa)
    info = A.info()
    x = torch.from_numpy(pk.gen_vector(m, 1)).cuda()
    y = torch.zeros(m, dtype=torch.float64, device="cuda")
    b = torch.from_numpy(pk.gen_vector(m, 2)).cuda()
    print(f"{name}: rows {m} nnz {nz} mean {nz/m:.1f}; plan exact={pk.KERNEL_NAMES[info.kernel_exact]} idx8={info.index8_diagonals}", flush=True)
    ref = None
    for tag, kern in (("stream", pk.KERNEL_STREAM), ("wmerge", pk.KERNEL_MERGE)):
        A.set_kernel(kern)
        for op, fn in (("MatMult", lambda: A.mult(x, y, pk.MODE_EXACT)), ("residual", lambda: A.residual(x, b, y, pk.MODE_EXACT))):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(30):
                fn()
            e1.record(); e1.synchronize()
            ms = e0.elapsed_time(e1) / 30
            extra = ""
            if op == "MatMult":
                if ref is None:
                    ref = y.clone()
                else:
                    extra = f" same bits: {bool(torch.equal(ref, y))}"
            print(f"   {tag:8s} {op:9s} {ms*1e3:8.1f} us  {(nz*12+m*20)/ms/1e6:7.0f} GB/s algorithmic{extra}", flush=True)
    A.destroy()
