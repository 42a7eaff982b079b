"""Power-law MatMult variants on one GPU: warp-granular exact-order kernel (k_wmerge) against the
split-row merge (round 1's block-granular pair k_mergex + k_longrow, removed since, ran 1.01 ms), and for the
transpose the stream kernel against k_wmerge.  Usage: python scripts/probe_powerlaw2.py [rows]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import petsc_openacc_b200 as pk

pk.init(0)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ai, aj, aa = pk.gen_powerlaw(m)
nz = len(aj)
A = pk.Csr(ai, aj, aa)
x = torch.from_numpy(pk.gen_vector(m, 1)).cuda()
y = torch.zeros(m, dtype=torch.float64, device="cuda")
yref = None


def timed(tag, fn):
    global yref
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 20
    same = ""
    if yref is None:
        yref = y.clone()
    else:
        same = f" same bits as first: {bool(torch.equal(y, yref))}"
    print(f"{tag:52s} {ms:8.4f} ms {nz/ms/1e6:7.1f} Gnnz/s  alg {(nz*12+m*20)/ms/1e6:6.0f} GB/s{same}", flush=True)


timed("exact: k_wmerge (default)", lambda: A.mult(x, y, pk.MODE_EXACT))
os.environ["B200_MERGE_SPLIT"] = "1"
timed("fast: split-row k_merge (B200_MERGE_SPLIT=1)", lambda: A.mult(x, y, pk.MODE_FAST))
os.environ.pop("B200_MERGE_SPLIT")
timed("exact_fma: k_wmerge", lambda: A.mult(x, y, pk.MODE_EXACT_FMA))
yref = None
A.build_transpose()
timed("transpose exact: plan default (stream)", lambda: A.mult_transpose(x, y, pk.MODE_EXACT))
# reference: the plain random gather out[k] = x[aj[k]] of the same index stream (a library op: the
# rate the memory system gives to 98 M independent 8-byte gathers from an 80 MB vector)
idx = torch.from_numpy(aj).cuda()
out = torch.empty(nz, dtype=torch.float64, device="cuda")
for _ in range(3):
    torch.index_select(x, 0, idx, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    torch.index_select(x, 0, idx, out=out)
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{'torch.index_select(x, aj) (gather only, writes 8 B/nnz)':52s} {ms:8.4f} ms {nz/ms/1e6:7.1f} Gnnz/s", flush=True)
