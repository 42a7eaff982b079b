"""Times the other BASELINE configs on one GPU: 27-point 200^3 (configs[3]) and the 10M-row
power-law matrix incl. MatMultTranspose (configs[4]).  Usage: python scripts/sweep_configs.py [27|pl]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gen
import petsc_openacc_b200 as pk


def timeit(fn, n=30, w=5):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def run(name, ai, aj, aa, n, transpose=False):
    m, nz = len(ai) - 1, len(aj)
    nbytes = nz * 12 + m * 20
    A = pk.Csr(ai, aj, aa, n=n)
    info = A.info()
    x = torch.from_numpy(gen.uniform_pm1(n)).cuda()
    y = torch.zeros(m, dtype=torch.float64, device="cuda")
    print(f"{name}: m={m} nz={nz} mean={nz/m:.2f} rmax={info.rmax} plan fast={pk.KERNEL_NAMES[info.kernel_fast]} exact={pk.KERNEL_NAMES[info.kernel_exact]} "
          f"tiles={info.stream_tiles} mtiles={info.merge_tiles} algorithmic bytes={nbytes}", flush=True)
    ref = None
    for kname, k, mode in (("auto-fast", pk.KERNEL_AUTO, pk.MODE_FAST), ("auto-exact", pk.KERNEL_AUTO, pk.MODE_EXACT),
                           ("auto-exfma", pk.KERNEL_AUTO, pk.MODE_EXACT_FMA),
                           ("row", pk.KERNEL_ROW, pk.MODE_EXACT_FMA), ("vector", pk.KERNEL_VECTOR, pk.MODE_FAST),
                           ("stream", pk.KERNEL_STREAM, pk.MODE_EXACT_FMA), ("merge", pk.KERNEL_MERGE, pk.MODE_FAST)):
        try:
            A.set_kernel(k)
        except pk.B200Error as e:
            print(f"  {kname:10s} n/a"); continue
        t = timeit(lambda: A.mult(x, y, mode))
        if ref is None:
            ref = y.clone()
        print(f"  {kname:10s} {t:8.4f} ms  {nbytes/t/1e6:8.1f} GB/s  {2*nz/t/1e6:7.1f} GFLOP/s  maxdiff {float((y-ref).abs().max()):.2e}", flush=True)
    A.set_kernel(pk.KERNEL_AUTO)
    if transpose:
        xt = torch.from_numpy(gen.uniform_pm1(m, 7)).cuda()
        yt = torch.zeros(n, dtype=torch.float64, device="cuda")
        t0 = time.time(); A.build_transpose(); print(f"  explicit transpose built in {time.time()-t0:.1f}s")
        for mode, nm in ((pk.MODE_FAST, "transpose fast"), (pk.MODE_EXACT, "transpose exact")):
            t = timeit(lambda: A.mult_transpose(xt, yt, mode))
            print(f"  {nm:16s} {t:8.4f} ms  {(nz*12+n*20)/t/1e6:8.1f} GB/s", flush=True)
    A.destroy()


which = sys.argv[1] if len(sys.argv) > 1 else "27"
pk.init(0)
if which == "27":
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    ai, aj, aa = gen.stencil27(N)
    run(f"27-point {N}^3", ai, aj, aa, N ** 3)
else:
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
    t0 = time.time(); ai, aj, aa = pk.gen_powerlaw(M); print(f"generated in {time.time()-t0:.0f}s")
    run(f"power-law {M}", ai, aj, aa, M, transpose=True)
