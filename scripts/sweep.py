"""Kernel sweep on one GPU: time every kernel / stream configuration on a stencil matrix.
Usage: python scripts/sweep.py [N] [stencil=7|27] [quick]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import gen
import petsc_openacc_b200 as pk


def time_kernel(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = np.array(ts)
    return float(np.median(ts)), float(ts.min())


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    st = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    pk.init(0)
    t = time.time()
    if st == 7:
        ai, aj, aa = gen.poisson7_natural(N, refpoint=False)
    elif st == 27:
        ai, aj, aa = gen.stencil27(N)
    else:
        ai, aj, aa = gen.powerlaw(N)
    m, nz = len(ai) - 1, len(aj)
    print(f"generated m={m} nz={nz} in {time.time()-t:.1f}s", flush=True)
    bytes_alg = nz * 12 + m * 20
    x = torch.from_numpy(gen.uniform_pm1(m)).cuda()
    y = torch.zeros(m, dtype=torch.float64, device="cuda")
    yref = None

    def report(tag, A, mode, kernel):
        nonlocal yref
        try:
            A.set_kernel(kernel)
        except pk.B200Error as e:
            print(f"{tag:48s} n/a ({e})")
            return
        med, best = time_kernel(lambda: A.mult(x, y, mode))
        chk = ""
        if yref is None:
            yref = y.clone()
        else:
            chk = f" maxdiff={float((y - yref).abs().max()):.3e}"
        print(f"{tag:48s} med {med:8.4f} ms  best {best:8.4f} ms  {bytes_alg/med/1e6:8.1f} GB/s (best {bytes_alg/best/1e6:8.1f}){chk}", flush=True)
        A.set_kernel(pk.KERNEL_AUTO)

    A = pk.Csr(ai, aj, aa)
    info = A.info()
    print("plan: fast=%s exact=%s lanes=%d tiles=%d rmax=%d" % (pk.KERNEL_NAMES[info.kernel_fast], pk.KERNEL_NAMES[info.kernel_exact], info.vector_lanes, info.stream_tiles, info.rmax))
    report("row exact", A, pk.MODE_EXACT, pk.KERNEL_ROW)
    report("row fma", A, pk.MODE_EXACT_FMA, pk.KERNEL_ROW)
    report("vector (auto lanes=%d)" % info.vector_lanes, A, pk.MODE_FAST, pk.KERNEL_VECTOR)
    report("stream default exact", A, pk.MODE_EXACT, pk.KERNEL_STREAM)
    report("stream default fma", A, pk.MODE_EXACT_FMA, pk.KERNEL_STREAM)
    # the optional SELL-32-sigma copy (k_sell): sigma = 1 keeps the natural row order
    for sigma in (1, 4096):
        try:
            A.build_sell(sigma)
        except pk.B200Error as e:
            print(f"sell sigma={sigma}: n/a ({e})")
            continue
        pad = A.info().sell_padded_nnz / max(nz, 1)
        report(f"sell-32-{sigma} exact (padding x{pad:.3f})", A, pk.MODE_EXACT, pk.KERNEL_SELL)
        report(f"sell-32-{sigma} fma", A, pk.MODE_EXACT_FMA, pk.KERNEL_SELL)
    A.destroy()
    mean = nz / m
    if len(sys.argv) > 3 and sys.argv[3] == "quick":   # kernels only, no stream-configuration sweep
        return
    for threads in (256, 128):
        for capmul in (1.0,):
            for stages in (1, 2, 3):
                for ctas in (2, 4, 6, 8, 12):
                    cap = int(threads * mean * capmul * 1.05) + 32
                    os.environ.update(B200_STREAM_THREADS=str(threads), B200_STREAM_CAP=str(cap),
                                      B200_STREAM_STAGES=str(stages), B200_STREAM_CTAS_PER_SM=str(ctas))
                    try:
                        A = pk.Csr(ai, aj, aa)
                    except pk.B200Error as e:
                        print("create failed", threads, cap, stages, ctas, e)
                        continue
                    if A.info().stream_tiles:
                        report(f"stream T={threads} cap={cap} S={stages} ctas/SM<={ctas}", A, pk.MODE_EXACT_FMA, pk.KERNEL_STREAM)
                    A.destroy()
    # copy roofline on this box for reference
    a = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
    b = torch.empty_like(a)
    med, best = time_kernel(lambda: b.copy_(a))
    print(f"torch copy 2x2GiB: best {2*a.numel()*8/best/1e6:.1f} GB/s")


if __name__ == "__main__":
    main()
