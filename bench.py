#!/usr/bin/env python
"""bench.py -- throughput of the SeqAIJ hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload matmult|stencil27|powerlaw|cg] [--grid G]

Default workload (`matmult`, the one BASELINE.json's metric is quoted on): a "step" is one MatMult
(y = A x) over the 300^3 7-point Poisson matrix: 188,460,000 non-zeros, 27,000,000 rows, fp64 values,
int32 indices -- BASELINE.json configs[1].  At N > 1 (torchrun, one rank per GPU) the same matrix is
row-partitioned like MatMult_MPIAIJ (configs[2]): strong scaling.

One JSON line on rank 0:
  value     whole-job algorithmic GB/s, (nnz*12 + rows*20) bytes per MatMult / device time,
            x and y resident in HBM, K back-to-back launches between two CUDA events;
  e2e       the same metric through the host-vector entry point (what MatMult_SeqAIJ(Mat,Vec,Vec)
            sees with PETSc 3.7.6 host Vecs): pinned-host x uploaded and y downloaded every step;
  roofline  achieved GB/s of the dominant kernel against MEASURED_PEAKS.json's copy bandwidth;
  cpu_baseline  the reference's CPU row loop on the box's host cores: oracle/_ref (its own loop text
            compiled from its patch files, kind "reference") when that was built, else the oracle's
            restatement (kind "port").
--impl reference times that CPU kernel as the arm to compare against (nothing of the product is loaded).

Other workloads (same line format, their own metric string; N = 1):
  stencil27  MatMult on the 27-point 200^3 matrix                   (BASELINE configs[3])
  powerlaw   MatMult and MatMultTranspose on the 10 M-row power-law matrix (BASELINE configs[4])
  cg         one iteration of KSPCG + PCJACOBI on the 300^3 problem: a "step" is one iteration
             (at N > 1 the row-partitioned solve)
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MatMult GB/s (% of HBM roofline) & GFLOP/s, 300^3 Poisson fp64, 1/2/4/8 B200"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
CPU_FLAGS = "gcc -O2 -ffp-contract=off, no -march (built in the CPU container, runs on the GPU box's host)"


def algorithmic_bytes(nnz, rows):
    """nnz*(8+4) + rows*(4+8+8), independent of the storage format."""
    return nnz * 12 + rows * 20


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(key="dram_bytes_per_launch"):
    """dram bytes per launch of the dominant kernel from the committed ncu summary of this round
    (a capture of this same command; profiles/ncu_summary.json says which), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    try:
        with open(p) as f:
            return json.load(f).get(key)
    except Exception:
        return None


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index=0, period=0.002):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.period, self.index = period, index
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {}
        for k in dir(nv):
            if k.startswith("nvmlClocksThrottleReason") or k.startswith("nvmlClocksEventReason"):
                v = getattr(nv, k)
                if isinstance(v, int) and v:
                    names.setdefault(v, k.replace("nvmlClocksThrottleReason", "").replace("nvmlClocksEventReason", ""))
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit and nm not in ("None", "All"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        # GpuIdle / ApplicationsClocksSetting are not throttles
        benign = {"GpuIdle", "ApplicationsClocksSetting"}
        rs = sorted(x for x in self.reasons if x not in benign)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": rs, "samples": len(self.samples)}


def gen_poisson(pk, n, size=1, rank=0):
    out = np.zeros(12, np.int32)
    pk.check(pk.lib.b200_gen_poisson7_info(n, n, n, size, rank, out.ctypes.data_as(C.c_void_p)))
    nloc, rstart, nnz = int(out[9]), int(out[10]), int(out[11])
    ai = np.zeros(nloc + 1, np.int32)
    aj = np.zeros(nnz, np.int32)
    aa = np.zeros(nnz)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    pk.check(pk.lib.b200_gen_poisson7(n, n, n, size, rank, 1, p(ai), p(aj), p(aa), None, None))
    return ai, aj, aa, rstart


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_kernel(oracle):
    """(threaded MatMult, kind, description): the reference's own loop text when oracle/_ref was
    built (ref_harness.c around the loops cut from src/openacc-step1/MatMult_SeqAIJ.patch), else
    the oracle's restatement of it.  Same bits either way (tests/test_oracle.py)."""
    if oracle.ref_lib() is not None:
        return oracle.ref_matmult_mt, "reference", "oracle/_ref/libref_matmult.so ref_matmult_mt: the reference's MatMult_SeqAIJ row loop"
    return oracle.matmult_mt, "port", "oracle/seqaij_oracle.c orc_matmult_mt"


def workload_name(args):
    n = args.grid
    return {
        "matmult": (f"3D Poisson 7-point {n}^3 fp64 MatMult_SeqAIJ (BASELINE configs[1])" if args.gpus == 1 else
                    f"3D Poisson 7-point {n}^3 fp64 MatMult_MPIAIJ row-partitioned over {args.gpus} ranks (BASELINE configs[2])"),
        "stencil27": f"3D 27-point stencil {n}^3 fp64 MatMult_SeqAIJ (BASELINE configs[3])",
        "powerlaw": f"power-law CSR, {n} rows, row lengths 1-10000, fp64 MatMult_SeqAIJ + MatMultTranspose (BASELINE configs[4])",
        "cg": (f"3D Poisson 7-point {n}^3 fp64, one KSPCG + PCJACOBI iteration" +
               ("" if args.gpus == 1 else f", row-partitioned over {args.gpus} ranks")),
    }[args.workload]


def workload_metric(args):
    return {"matmult": METRIC,
            "stencil27": "MatMult GB/s (% of HBM roofline), 27-point 200^3 fp64, 1 B200",
            "powerlaw": "MatMult GB/s, power-law 10M rows fp64 (merge-path + MatMultTranspose), 1 B200",
            "cg": "CG iteration (KSPCG + PCJACOBI) GB/s of the bytes one iteration must move, 300^3 Poisson fp64"}[args.workload]


def workload_config(args, rows, nnz):
    """`config` of the JSON line: the same dictionary on both arms (the driver compares them)."""
    return {"workload": workload_name(args), "rows": int(rows), "nnz": int(nnz),
            "algorithmic_bytes": int(algorithmic_bytes(nnz, rows)),
            "l2": "inputs (0.3-2.8 GB per GPU) larger than the 126 MB L2; no flush"}


def build_matrix(args, product):
    """(ai, aj, aa) of the workload.  product: the package's threaded C++ generators; otherwise the
    oracle's / numpy generators, so that the reference arm loads nothing of the product."""
    w, n = args.workload, args.grid
    if w in ("matmult", "cg"):
        if product is not None:
            return gen_poisson(product, n)[:3]
        import oracle
        p = oracle.poisson7(n)
        return p["ai"], p["aj"], p["aa"]
    if w == "stencil27":
        if product is not None:
            return product.gen_stencil27(n)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import gen
        return gen.stencil27(n)
    if product is not None:
        return product.gen_powerlaw(n)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gen
    return gen.powerlaw(n)


def cg_bytes(nnz, rows):
    """Bytes one Jacobi-CG iteration must move with the three-pass schedule of b200_vec.cu: the
    MatMult's algorithmic bytes + 10 vector reads/writes (k_cg_p: x p r dinv -> x p; k_cg_r: r w dinv -> r)."""
    return algorithmic_bytes(nnz, rows) + 10 * rows * 8


def run_reference(args):
    """The reference's CPU implementation of the path (its own MatMult_SeqAIJ loop text from
    oracle/_ref, or the oracle's restatement where that is not built; PETSc as a whole cannot be
    built offline), one contiguous row block per host thread ~ one MPI rank per core.  Nothing of the
    product is imported on this arm: matrix and vectors come from the oracle's generators."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    ai, aj, aa = build_matrix(args, None)
    m, nnz = len(ai) - 1, len(aj)
    ncols = m
    cores = host_threads()
    matmult_mt, kind, what = cpu_kernel(oracle)
    if args.workload == "cg":
        # the oracle's KSPCG + PCJACOBI (one core: its vector loops are scalar C), `steps` iterations
        p = oracle.poisson7(args.grid)
        t0 = time.perf_counter()
        _, its, _ = oracle.cg_jacobi(ai, aj, aa, p["rhs"], rtol=1e-30, atol=1e-300, max_it=args.steps)
        dt = (time.perf_counter() - t0) / max(abs(its), 1)
        nbytes, cores, what = cg_bytes(nnz, m), 1, "oracle/seqaij_oracle.c orc_cg_jacobi"
        kind = "port"
        sample = f"{abs(its)} iterations of the oracle's KSPCG + PCJACOBI on the {args.grid}^3 problem, one core"
    else:
        x = oracle.gen_vector(ncols, 0xB200)
        y = np.empty(m)
        for _ in range(max(args.warmup, 1)):
            matmult_mt(ai, aj, aa, x, cores, y=y)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            matmult_mt(ai, aj, aa, x, cores, y=y)
        dt = (time.perf_counter() - t0) / args.steps
        nbytes = algorithmic_bytes(nnz, m)
        sample = f"{args.steps} full MatMults ({what}), one nnz-balanced row block per thread"
    gbs = nbytes / dt / 1e9
    line = {
        "impl": "reference", "metric": workload_metric(args), "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "gflops": 2.0 * nnz / dt / 1e9,
        "config": workload_config(args, m, nnz),
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": kind, "flags": CPU_FLAGS, "sample": sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline(ai, aj, aa, x, budget_s=2.0, one_core=True):
    import oracle
    cores = host_threads()
    m, nnz = len(ai) - 1, len(aj)
    y = np.empty(m)
    matmult_mt, kind, what = cpu_kernel(oracle)
    matmult_mt(ai, aj, aa, x, cores, y=y)
    reps, t0 = 0, time.perf_counter()
    while reps < 10 or (time.perf_counter() - t0 < budget_s and reps < 200):
        matmult_mt(ai, aj, aa, x, cores, y=y)
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    out = {"value": algorithmic_bytes(nnz, m) / dt / 1e9, "unit": "GB/s", "cores": cores,
           "kind": kind, "ms_per_matmult": dt * 1e3, "flags": CPU_FLAGS,
           "sample": f"{reps} full MatMults of this matrix ({what}), one row block per thread"}
    if one_core:
        # the reference's 1-core "original" protocol (runs/single-node-scaling.pbs:56), two passes
        t1 = time.perf_counter()
        for _ in range(2):
            matmult_mt(ai, aj, aa, x, 1, y=y)
        dt1 = (time.perf_counter() - t1) / 2
        out.update(value_1core=algorithmic_bytes(nnz, m) / dt1 / 1e9, ms_per_matmult_1core=dt1 * 1e3)
    return out, y


def time_launches(torch, fn, steps, warmup, stream):
    """(ms per step over `steps` back-to-back launches between two events, per-launch times of a second
    pass with an event after every launch)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record(stream)
    for k in range(steps):
        fn()
        ev[k + 1].record(stream)
    torch.cuda.synchronize()
    per = np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(steps)])
    return ms, per


def run_ours(args):
    import torch

    import petsc_openacc_b200 as pk
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torchrun (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU fallback")
    torch.cuda.set_device(local)
    pk.init(local)
    if world > 1:
        if args.workload not in ("matmult", "cg"):
            raise SystemExit("the stencil27 / powerlaw workloads are single-GPU lines")
        import bench_mpiaij
        return bench_mpiaij.run(args, pk, sys.modules[__name__])
    if args.workload == "cg":
        return run_cg(args, pk, torch, local)

    ai, aj, aa = build_matrix(args, pk)
    m, nnz = len(ai) - 1, len(aj)
    nbytes = algorithmic_bytes(nnz, m)
    A = pk.Csr(ai, aj, aa)
    info = A.info()
    mode = {"fast": pk.MODE_FAST, "exact": pk.MODE_EXACT, "exact_fma": pk.MODE_EXACT_FMA}[args.mode]
    hx = pk.PinnedArray(m)
    hy = pk.PinnedArray(m)
    hx.array[:] = pk.gen_vector(m, 0xB200)
    x = torch.from_numpy(hx.array).cuda()
    y = torch.zeros(m, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream()
    warm = max(args.warmup, 3)

    # ---- device-resident throughput --------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    l0 = pk.launch_count()
    A.mult(x, y, mode, stream)
    launches = (pk.launch_count() - l0) * args.steps      # kernels of this library inside the timed region
    ms, per = time_launches(torch, lambda: A.mult(x, y, mode, stream), args.steps, warm, stream)
    value = nbytes / ms / 1e6

    # ---- end to end through the host-vector entry (H2D x + kernel + D2H y every step) -------
    e2e_steps = max(5, min(args.steps, 30))
    for _ in range(2):
        A.mult_host(hx.array, hy.array, mode)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        A.mult_host(hx.array, hy.array, mode)
    e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
    clocks = sampler.stop()
    y_host = hy.array.copy()
    host_matches_device = bool(np.array_equal(y_host, y.cpu().numpy()))

    extra = {}
    # the same MatMult with plain int32 column indices (no diagonal-code compression), so that the
    # number against the 12-bytes-per-non-zero model is on record next to the default plan's
    plain = None
    if info.index8_diagonals:
        os.environ["B200_INDEX8"] = "0"
        A32 = pk.Csr(ai, aj, aa)
        os.environ.pop("B200_INDEX8")
        y32 = torch.zeros(m, dtype=torch.float64, device="cuda")
        ms32, _ = time_launches(torch, lambda: A32.mult(x, y32, mode, stream), max(20, args.steps // 4), 5, stream)
        plain = {"ms_per_step": ms32, "value": nbytes / ms32 / 1e6, "unit": "GB/s",
                 "same_bits_as_default_plan": bool(torch.equal(y32, y)),
                 "dram_bytes_model": nbytes, "note": "B200_INDEX8=0: 4-byte column indices streamed"}
        A32.destroy()
        del y32
    if args.workload == "powerlaw":
        # MatMultTranspose through the cached explicit transpose (built once, on the device)
        A.build_transpose()
        yt = torch.zeros(m, dtype=torch.float64, device="cuda")
        mst, _ = time_launches(torch, lambda: A.mult_transpose(x, yt, mode, stream), max(10, args.steps // 2), 3, stream)
        import oracle
        want = oracle.matmulttranspose(ai, aj, aa, hx.array, m, fma=(mode != pk.MODE_EXACT))
        extra["transpose"] = {"ms_per_step": mst, "value": nbytes / mst / 1e6, "unit": "GB/s",
                              "parity_vs_oracle": "bit-exact" if np.array_equal(want, yt.cpu().numpy()) else "MISMATCH"}
        del yt
    peak, peak_src = measured_peak()
    kname = pk.KERNEL_NAMES[info.kernel_fast if mode == pk.MODE_FAST else info.kernel_exact]
    if kname == "merge":
        kname = "wmerge"
    stream_bytes = nnz * (9 if info.index8_diagonals else 12) + m * 16 + int(m * (1.125 if info.rowlen8 else 4))
    roof = {"bound": "hbm", "achieved": nbytes / ms / 1e6, "peak": peak, "unit": "GB/s",
            "frac": nbytes / ms / 1e6 / peak,
            "traffic": ncu_traffic() if args.workload == "matmult" else ncu_traffic(f"dram_bytes_per_launch_{args.workload}"),
            "kernel": f"k_{kname}", "kernel_ms": ms, "kernel_ms_median_with_event_per_launch": float(np.median(per)),
            "kernel_ms_best": float(per.min()), "peak_source": peak_src,
            "frac_of_nominal_8000": nbytes / ms / 1e6 / 8000.0,
            "dram_bytes_streamed_model": stream_bytes,
            "dram_gbs_streamed_model": stream_bytes / ms / 1e6,
            "note": ("achieved = ALGORITHMIC bytes (nnz*12 + rows*20) / average launch duration over the timed region. The "
                     "default plan streams 1-byte diagonal codes instead of 4-byte column indices and 1-byte row lengths instead "
                     "of 4-byte row pointers (lossless, bit-exact), so real DRAM traffic is nnz*9 + rows*17.1 and frac can exceed "
                     "1; 'int32_index' is the same MatMult without that compression.") if info.index8_diagonals else
                    "achieved = algorithmic bytes (nnz*12 + rows*20) / average launch duration over the timed region"}
    if plain:
        roof["int32_index"] = dict(plain, frac=plain["value"] / peak)
    cpu, y_cpu = cpu_baseline(ai, aj, aa, hx.array)
    # parity of the timed result against the oracle, reported (the tests are the gate)
    if mode == pk.MODE_EXACT:
        parity = "bit-exact" if np.array_equal(y_cpu, y_host) else "MISMATCH"
    else:
        import oracle
        bound = 1e-13 * oracle.row_abs_sum(ai, aj, aa, hx.array)
        parity = "within 1e-13 row bound" if np.all(np.abs(y_cpu - y_host) <= bound) else "MISMATCH"
    if not host_matches_device:
        parity += " (host-vector path and device path DISAGREE)"
    line = {
        "metric": workload_metric(args), "value": value, "unit": "GB/s", "n_gpus": 1, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "gflops": 2.0 * nnz / ms / 1e6,
        "config": workload_config(args, m, nnz),
        "plan": {"mode": args.mode, "kernel": f"k_{kname}", "index8_diagonals": int(info.index8_diagonals),
                 "rowlen8": int(info.rowlen8), "parity_vs_oracle": parity, "pdl": os.environ.get("B200_PDL", "1") != "0"},
        "roofline": roof, "cpu_baseline": cpu,
        "e2e": {"value": nbytes / e2e_ms / 1e6, "unit": "GB/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": m * 8, "d2h_bytes_per_step": m * 8, "steps": e2e_steps,
                "pcie_gbs_per_direction": m * 8 / e2e_ms / 1e6,
                "api": "b200_spmv_host (MatMult_SeqAIJ with host Vecs, pinned)"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    line.update(extra)
    print(json.dumps(line), flush=True)
    A.destroy()
    return 0


def run_cg(args, pk, torch, local):
    """One B200: `steps` iterations of the fused KSPCG + PCJACOBI (b200_cg_jacobi) on the reference
    problem; the solve is stopped by max_it (rtol 0), so every iteration is a full one."""
    import oracle
    n = args.grid
    ai, aj, aa = build_matrix(args, pk)
    m, nnz = len(ai) - 1, len(aj)
    rhs = oracle.poisson7(n)["rhs"] if n <= 100 else pk.gen_poisson7(n, vectors=True)["rhs"]
    A = pk.Csr(ai, aj, aa)
    info = A.info()
    mode = {"fast": pk.MODE_FAST, "exact": pk.MODE_EXACT, "exact_fma": pk.MODE_EXACT_FMA}[args.mode]
    hb, hxs = pk.PinnedArray(m), pk.PinnedArray(m)
    hb.array[:] = rhs
    b = torch.from_numpy(hb.array).cuda()
    x = torch.zeros(m, dtype=torch.float64, device="cuda")
    warm = max(args.warmup, 3)
    A.cg_jacobi(b, x, rtol=1e-30, atol=1e-300, max_it=warm, mode=mode)
    sampler = ClockSampler(local)
    sampler.start()
    res = A.cg_jacobi(b, x, rtol=1e-30, atol=1e-300, max_it=args.steps, mode=mode)
    assert res.its == args.steps, (res.its, res.reason)
    ms = res.solve_ms / res.its
    # end to end: right-hand side up from pinned host memory, solution down, every solve
    t0 = time.perf_counter()
    b.copy_(torch.from_numpy(hb.array), non_blocking=True)
    res2 = A.cg_jacobi(b, x, rtol=1e-30, atol=1e-300, max_it=args.steps, mode=mode)
    torch.from_numpy(hxs.array).copy_(x)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / res2.its
    clocks = sampler.stop()
    nbytes = cg_bytes(nnz, m)
    peak, peak_src = measured_peak()
    moved = nnz * (9 if info.index8_diagonals else 12) + m * 20 + 10 * m * 8
    # parity: a converging solve of a small grid against the oracle's CG in the same run
    small = oracle.poisson7(24)
    As = pk.Csr(small["ai"], small["aj"], small["aa"])
    xs = torch.zeros(As.m, dtype=torch.float64, device="cuda")
    rs = As.cg_jacobi(torch.from_numpy(small["rhs"]).cuda(), xs, rtol=1e-10, atol=1e-50, max_it=5000, mode=mode)
    _, its_o, _ = oracle.cg_jacobi(small["ai"], small["aj"], small["aa"], small["rhs"], rtol=1e-10, atol=1e-50, max_it=5000)
    As.destroy()
    # CPU beside it: the oracle's CG, a bounded number of iterations on one core
    t1 = time.perf_counter()
    k_cpu = 4 if n >= 200 else 20
    oracle.cg_jacobi(ai, aj, aa, rhs, rtol=1e-30, atol=1e-300, max_it=k_cpu)
    cpu_ms = (time.perf_counter() - t1) * 1e3 / k_cpu
    line = {
        "metric": workload_metric(args), "value": nbytes / ms / 1e6, "unit": "GB/s", "n_gpus": 1, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args, m, nnz), bytes_per_iteration=int(nbytes)),
        "plan": {"mode": args.mode, "launches_per_iteration": res.launches / res.its,
                 "parity_vs_oracle": f"24^3 solve: {rs.its} iterations, oracle {its_o}"},
        "roofline": {"bound": "hbm", "achieved": nbytes / ms / 1e6, "peak": peak, "unit": "GB/s", "frac": nbytes / ms / 1e6 / peak,
                     "traffic": ncu_traffic("dram_bytes_per_iteration_cg"), "peak_source": peak_src,
                     "kernel": "k_cg_p + k_stream<EPI_DOT> + k_cg_r", "dram_bytes_moved_model": moved,
                     "dram_gbs_moved_model": moved / ms / 1e6,
                     "note": "bytes = MatMult algorithmic bytes (nnz*12 + rows*20) + 10 vector passes of rows*8 (DESIGN.md)"},
        "cpu_baseline": {"value": nbytes / cpu_ms / 1e6, "unit": "GB/s", "cores": 1, "kind": "port", "flags": CPU_FLAGS,
                         "ms_per_iteration": cpu_ms, "sample": f"{k_cpu} iterations of oracle orc_cg_jacobi on the same problem"},
        "e2e": {"value": nbytes / e2e_ms / 1e6, "unit": "GB/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": m * 8 / res2.its,
                "d2h_bytes_per_step": m * 8 / res2.its, "api": "b200_cg_jacobi: rhs up and solution down once per solve"},
        "gpu_launches": int(res.launches), "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    A.destroy()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="matmult", choices=["matmult", "stencil27", "powerlaw", "cg"])
    ap.add_argument("--grid", type=int, default=None, help="grid size (rows for powerlaw); default per workload")
    ap.add_argument("--mode", default="exact", choices=["fast", "exact", "exact_fma"])
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"])
    args = ap.parse_args()
    if args.grid is None:
        args.grid = {"matmult": 300, "cg": 300, "stencil27": 200, "powerlaw": 10_000_000}[args.workload]
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
