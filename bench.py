#!/usr/bin/env python
"""bench.py -- MatMult throughput of the SeqAIJ hot path on the 300^3 7-point Poisson matrix.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--grid 300]

A "step" is one MatMult (y = A x) over the whole matrix: 188,460,000 non-zeros, 27,000,000 rows,
fp64 values, int32 indices -- BASELINE.json configs[1].  At N > 1 (torchrun, one rank per GPU) the
same matrix is row-partitioned like MatMult_MPIAIJ (configs[2]): strong scaling.

One JSON line on rank 0:
  value     whole-job algorithmic GB/s, (nnz*12 + rows*20) bytes per MatMult / device time,
            x and y resident in HBM;
  e2e       the same metric through the host-vector entry point (what MatMult_SeqAIJ(Mat,Vec,Vec)
            sees with PETSc 3.7.6 host Vecs): pinned-host x uploaded and y downloaded every step;
  roofline  achieved GB/s of the dominant kernel against MEASURED_PEAKS.json's copy bandwidth;
  cpu_baseline  the reference's CPU row loop on the box's host cores: oracle/_ref (its own loop text
            compiled from its patch files, kind "reference") when that was built, else the oracle's
            restatement (kind "port").
--impl reference times that CPU kernel as the arm to compare against.
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


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu summary, or None."""
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    try:
        with open(p) as f:
            return json.load(f).get("dram_bytes_per_launch")
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


def workload_config(n, gpus, rows, nnz):
    """`config` of the JSON line: the same dictionary on both arms (the driver compares them)."""
    if gpus == 1:
        what = f"3D Poisson 7-point {n}^3 fp64 MatMult_SeqAIJ (BASELINE configs[1])"
    else:
        what = f"3D Poisson 7-point {n}^3 fp64 MatMult_MPIAIJ row-partitioned over {gpus} ranks (BASELINE configs[2])"
    return {"workload": what, "rows": int(rows), "nnz": int(nnz),
            "algorithmic_bytes": int(algorithmic_bytes(nnz, rows)),
            "l2": "inputs (2.2-2.8 GB; 280-350 MB per rank at 8) larger than the 126 MB L2; no flush"}


def run_reference(args):
    """The reference's CPU MatMult_SeqAIJ loop (its own text from oracle/_ref, or the oracle's
    restatement where that is not built; PETSc as a whole cannot be built offline), one contiguous
    row block per host thread ~ one MPI rank per core.  Nothing of the product is imported on this
    arm: the matrix and x come from the oracle's generator."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    n = args.grid
    p = oracle.poisson7(n)
    ai, aj, aa = p["ai"], p["aj"], p["aa"]
    m, nnz = len(ai) - 1, len(aj)
    x = oracle.gen_vector(m, 0xB200)
    y = np.empty(m)
    cores = host_threads()
    matmult_mt, kind, what = cpu_kernel(oracle)
    for _ in range(max(args.warmup, 1)):
        matmult_mt(ai, aj, aa, x, cores, y=y)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        matmult_mt(ai, aj, aa, x, cores, y=y)
    dt = (time.perf_counter() - t0) / args.steps
    gbs = algorithmic_bytes(nnz, m) / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "gflops": 2.0 * nnz / dt / 1e9,
        "config": workload_config(n, args.gpus, m, nnz),
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": kind,
                         "flags": "gcc -O2 -ffp-contract=off, no -march (built in the CPU container, runs on the GPU box's host)",
                         "sample": f"{args.steps} full {n}^3 MatMults ({what}), one nnz-balanced row block per thread"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline(ai, aj, aa, x, n):
    import oracle
    cores = host_threads()
    m, nnz = len(ai) - 1, len(aj)
    y = np.empty(m)
    matmult_mt, kind, what = cpu_kernel(oracle)
    matmult_mt(ai, aj, aa, x, cores, y=y)
    reps, t0 = 0, time.perf_counter()
    while reps < 10 or (time.perf_counter() - t0 < 2.0 and reps < 200):
        matmult_mt(ai, aj, aa, x, cores, y=y)
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    # the reference's 1-core "original" protocol (runs/single-node-scaling.pbs:56), two passes
    t1 = time.perf_counter()
    for _ in range(2):
        matmult_mt(ai, aj, aa, x, 1, y=y)
    dt1 = (time.perf_counter() - t1) / 2
    return {"value": algorithmic_bytes(nnz, m) / dt / 1e9, "unit": "GB/s", "cores": cores,
            "kind": kind, "ms_per_matmult": dt * 1e3,
            "flags": "gcc -O2 -ffp-contract=off, no -march (built in the CPU container, runs on the GPU box's host)",
            "value_1core": algorithmic_bytes(nnz, m) / dt1 / 1e9, "ms_per_matmult_1core": dt1 * 1e3,
            "sample": f"{reps} full {n}^3 MatMults ({what}), one row block per thread"}, y


def run_ours(args):
    import torch
    import torch.distributed as dist

    import petsc_openacc_b200 as pk
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torchrun (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU fallback")
    torch.cuda.set_device(local)
    pk.init(local)
    if world > 1:
        import bench_mpiaij
        return bench_mpiaij.run(args, pk, METRIC, algorithmic_bytes, measured_peak, ClockSampler, workload_config)

    n = args.grid
    ai, aj, aa, _ = gen_poisson(pk, n)
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

    # ---- device-resident throughput --------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        A.mult(x, y, mode)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = pk.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    torch.cuda.synchronize()
    ev[0].record(stream)
    for k in range(args.steps):
        A.mult(x, y, mode, stream)
        ev[k + 1].record(stream)
    torch.cuda.synchronize()
    launches = pk.launch_count() - l0
    total_ms = ev[0].elapsed_time(ev[-1])
    per = np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)])
    ms = total_ms / args.steps
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

    # the same MatMult with plain int32 column indices (no diagonal-code compression), so that the
    # number against the 12-bytes-per-non-zero model is on record next to the default plan's
    plain = None
    if info.index8_diagonals:
        os.environ["B200_INDEX8"] = "0"
        A32 = pk.Csr(ai, aj, aa)
        os.environ.pop("B200_INDEX8")
        y32 = torch.zeros(m, dtype=torch.float64, device="cuda")
        for _ in range(5):
            A32.mult(x, y32, mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n32 = max(20, args.steps // 4)
        e0.record(stream)
        for _ in range(n32):
            A32.mult(x, y32, mode, stream)
        e1.record(stream)
        torch.cuda.synchronize()
        ms32 = e0.elapsed_time(e1) / n32
        plain = {"ms_per_step": ms32, "value": nbytes / ms32 / 1e6, "unit": "GB/s",
                 "same_bits_as_default_plan": bool(torch.equal(y32, y)),
                 "dram_bytes_model": nbytes, "note": "B200_INDEX8=0: 4-byte column indices streamed"}
        A32.destroy()
        del y32
    peak, peak_src = measured_peak()
    kname = pk.KERNEL_NAMES[info.kernel_fast if mode == pk.MODE_FAST else info.kernel_exact]
    kernel_ms = float(np.median(per))
    stream_bytes = nnz * (9 if info.index8_diagonals else 12) + m * 20
    roof = {"bound": "hbm", "achieved": nbytes / kernel_ms / 1e6, "peak": peak, "unit": "GB/s",
            "frac": nbytes / kernel_ms / 1e6 / peak, "traffic": ncu_traffic(),
            "kernel": f"k_{kname}", "kernel_ms_median": kernel_ms, "kernel_ms_best": float(per.min()),
            "peak_source": peak_src,
            "frac_of_nominal_8000": nbytes / kernel_ms / 1e6 / 8000.0,
            "dram_bytes_streamed_model": stream_bytes,
            "dram_gbs_streamed_model": stream_bytes / kernel_ms / 1e6,
            "note": ("achieved = ALGORITHMIC bytes (nnz*12 + rows*20) / time. The default plan streams 1-byte diagonal codes "
                     "instead of 4-byte column indices (lossless, bit-exact), so real DRAM traffic is nnz*9 + rows*20 and frac "
                     "can exceed 1; 'int32_index' below is the same MatMult without that compression.") if info.index8_diagonals else
                    "achieved = algorithmic bytes (nnz*12 + rows*20) / time"}
    if plain:
        roof["int32_index"] = dict(plain, frac=plain["value"] / peak)
    cpu, y_cpu = cpu_baseline(ai, aj, aa, hx.array, n)
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
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "gflops": 2.0 * nnz / ms / 1e6,
        "config": workload_config(n, 1, m, nnz),
        "plan": {"mode": args.mode, "kernel": f"k_{kname}", "index8_diagonals": int(info.index8_diagonals),
                 "parity_vs_oracle": parity},
        "roofline": roof, "cpu_baseline": cpu,
        "e2e": {"value": nbytes / e2e_ms / 1e6, "unit": "GB/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": m * 8, "d2h_bytes_per_step": m * 8, "steps": e2e_steps,
                "api": "b200_spmv_host (MatMult_SeqAIJ with host Vecs, pinned)"},
        "gpu_launches": int(launches), "clocks": clocks,
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
    ap.add_argument("--grid", type=int, default=300)
    ap.add_argument("--mode", default="exact", choices=["fast", "exact", "exact_fma"])
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
