"""bench_mpiaij.py -- the N > 1 leg of bench.py: MatMult_MPIAIJ on the row-partitioned 300^3
Poisson matrix (BASELINE configs[2]), one rank per GPU, strong scaling.

torch.distributed (NCCL) is plumbing only: rendezvous, exchanging the garray lists and the 64-byte
CUDA-IPC handles, barriers and the max-over-ranks of the device time.  The halo itself is pushed
by this library's kernels over NVLink peer mappings; `--halo nccl` times the NCCL send/recv
transport with the same pack / off-diagonal kernels as the baseline it replaces.
"""
import json
import os
import time

import numpy as np


def run(args, pk, B):
    """B = the bench module (metric strings, byte formulas, clock sampler, config)."""
    import torch
    import torch.distributed as dist
    algorithmic_bytes, measured_peak, ClockSampler = B.algorithmic_bytes, B.measured_peak, B.ClockSampler
    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = args.grid
    mode = {"fast": pk.MODE_FAST, "exact": pk.MODE_EXACT, "exact_fma": pk.MODE_EXACT_FMA}[args.mode]

    g = pk.gen_poisson7(n, world, rank)
    base = g["base"]
    M = pk.MpiAij(world, rank, base, g["ai"], g["aj"], g["aa"])
    garrays = [None] * world
    dist.all_gather_object(garrays, M.garray())
    for q in range(world):
        M.set_peer_garray(q, garrays[q])
    M.upload()
    handles = [None] * world
    dist.all_gather_object(handles, M.ipc_handle())
    halo = "p2p-push"
    try:
        for q in range(world):
            if q != rank and len(M.send_list(q)[0]):
                M.open_peer_window(q, handles[q])
    except pk.B200Error as e:
        halo = f"nccl (cuda ipc unavailable: {e})"
    flag = torch.tensor([0 if halo == "p2p-push" else 1], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    use_p2p = int(flag.item()) == 0 and args.halo != "nccl"
    if not use_p2p and halo == "p2p-push":
        halo = "nccl"

    nloc = M.nloc
    rows_global, nnz_global = int(base[-1]), 7 * n ** 3 - 6 * n ** 2
    nbytes = algorithmic_bytes(nnz_global, rows_global)
    if args.workload == "cg":
        return run_cg(args, pk, B, M, g, dev, rows_global, nnz_global)
    xg = pk.gen_vector(rows_global, 0xB200)
    hx, hy = pk.PinnedArray(nloc), pk.PinnedArray(nloc)
    hx.array[:] = xg[base[rank]:base[rank + 1]]
    x = torch.from_numpy(hx.array).to(dev)
    y = torch.zeros(nloc, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()

    # NCCL transport pieces (baseline)
    sends = {q: M.send_list(q) for q in range(world) if q != rank}
    sends = {q: v for q, v in sends.items() if len(v[0])}
    roff = M.recv_offsets()
    lvec = torch.zeros(max(M.nghost, 1), dtype=torch.float64, device=dev)
    sbufs = {q: torch.empty(len(v[0]), dtype=torch.float64, device=dev) for q, v in sends.items()}

    def mult_nccl():
        ops = []
        for q, buf in sbufs.items():
            M.pack(q, x, buf)
            ops.append(dist.P2POp(dist.isend, buf, q))
        for q in range(world):
            if q != rank and roff[q + 1] > roff[q]:
                ops.append(dist.P2POp(dist.irecv, lvec[roff[q]:roff[q + 1]], q))
        reqs = dist.batch_isend_irecv(ops) if ops else []
        M.mult_local(x, y, mode)
        for r in reqs:
            r.wait()
        M.mult_add_ghost(lvec, y, mode)

    def mult_p2p():
        M.mult(x, y, mode, stream)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    warm = max(args.warmup, 3)
    sampler = ClockSampler(local)
    l0 = pk.launch_count()
    if use_p2p:
        if rank == 0:
            sampler.start()
        ms = timed(mult_p2p, args.steps, warm)
        M.check()
        launches = (pk.launch_count() - l0) * args.steps // (args.steps + warm)     # of the timed region, this rank
        ms_nccl = timed(mult_nccl, max(5, args.steps // 4), 3)
    else:
        if rank == 0:
            sampler.start()
        ms = timed(mult_nccl, args.steps, warm)
        launches = (pk.launch_count() - l0) * args.steps // (args.steps + warm)
        ms_nccl = ms
    y_dev = y.cpu().numpy()

    # end to end with host vectors: every rank uploads its x rows and downloads its y rows
    e2e_steps = max(5, min(args.steps, 30))
    if use_p2p:
        for _ in range(2):
            M.mult_host(hx.array, hy.array, mode)
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            M.mult_host(hx.array, hy.array, mode)
        dist.barrier()
        e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
        M.check()
    else:
        # NCCL transport: the same steps spelled out (x rows up, MatMult, y rows down)
        hx_t, hy_t = torch.from_numpy(hx.array), torch.from_numpy(hy.array)

        def host_step():
            x.copy_(hx_t)
            mult_nccl()
            hy_t.copy_(y)
            torch.cuda.synchronize()
        for _ in range(2):
            host_step()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_step()
        dist.barrier()
        e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # parity of the timed result against the oracle (each rank checks its own rows)
    import oracle
    Ai, Aj, Aa = M.block(0)
    Bi, Bj, Ba = M.block(1)
    fma = mode != pk.MODE_EXACT
    ref = oracle.matmult(Ai, Aj, Aa, hx.array, fma=fma)
    if M.nghost:
        ref = oracle.matmultadd(Bi, Bj, Ba, xg[garrays[rank]], ref, fma=fma)
    ok = torch.tensor([1 if np.array_equal(ref, y_dev) else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    parity = "bit-exact" if int(ok.item()) == 1 else "MISMATCH"

    if rank == 0:
        peak, peak_src = measured_peak()
        value = nbytes / ms / 1e6
        line = {
            "metric": B.workload_metric(args), "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "gflops": 2.0 * nnz_global / ms / 1e6,
            "config": B.workload_config(args, rows_global, nnz_global),
            "plan": {"mode": args.mode, "halo": halo, "process_grid": [int(v) for v in g["info"][:3]],
                     "rows_per_rank": nloc, "nghost_rank0": M.nghost, "parity_vs_oracle": parity},
            "roofline": {"bound": "hbm", "achieved": value / world, "peak": peak, "unit": "GB/s",
                         "frac": value / world / peak, "traffic": B.ncu_traffic(f"dram_bytes_per_launch_halo_{world}"),
                         "kernel": "k_stream<HALO> (one fused launch per rank: NVLink push + A x + ghost rows)",
                         "peak_source": peak_src, "note": "per-GPU share of the whole-job rate, halo and launches included"},
            "nccl_halo": {"ms_per_step": ms_nccl, "value": nbytes / ms_nccl / 1e6, "unit": "GB/s",
                          "note": "same pack/off-diagonal kernels, torch.distributed batch_isend_irecv transport"},
            "e2e": {"value": nbytes / e2e_ms / 1e6, "unit": "GB/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": rows_global * 8, "d2h_bytes_per_step": rows_global * 8,
                    "steps": e2e_steps, "api": "b200_mpiaij_mult_host (MatMult_MPIAIJ with host Vecs, pinned)" if use_p2p else "host rows up + NCCL-transport MatMult + rows down"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    M.destroy()
    dist.destroy_process_group()
    return 0


def run_cg(args, pk, B, M, g, dev, rows_global, nnz_global):
    """`steps` iterations of KSPCG + PCJACOBI on the row-partitioned matrix (b200_mpiaij_cg_jacobi):
    fused MatMult_MPIAIJ with (p, A p) folded in, all-reduces over the peer windows, no host sync."""
    import torch
    import torch.distributed as dist
    world, rank = M.size, M.rank
    handles = [None] * world
    dist.all_gather_object(handles, M.ipc_handle())
    for q in range(world):
        M.set_rank_window(q, handle=handles[q] if q != rank else None)
    gv = pk.gen_poisson7(args.grid, world, rank, vectors=True)
    b = torch.from_numpy(gv["rhs"]).to(dev)
    x = torch.zeros(M.nloc, dtype=torch.float64, device=dev)
    mode = {"fast": pk.MODE_FAST, "exact": pk.MODE_EXACT, "exact_fma": pk.MODE_EXACT_FMA}[args.mode]
    warm = max(args.warmup, 3)
    M.cg_jacobi(b, x, rtol=1e-30, atol=1e-300, max_it=warm, mode=mode)
    torch.cuda.synchronize()
    dist.barrier()
    sampler = B.ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    res = M.cg_jacobi(b, x, rtol=1e-30, atol=1e-300, max_it=args.steps, mode=mode)
    M.check()
    t = torch.tensor([res.solve_ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    its = torch.tensor([res.its], device=dev)
    dist.all_reduce(its, op=dist.ReduceOp.MIN)
    ms = float(t.item()) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        nbytes = B.cg_bytes(nnz_global, rows_global)
        peak, peak_src = B.measured_peak()
        line = {
            "metric": B.workload_metric(args), "value": nbytes / ms / 1e6, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict(B.workload_config(args, rows_global, nnz_global), bytes_per_iteration=int(nbytes)),
            "plan": {"mode": args.mode, "launches_per_iteration": res.launches / max(res.its, 1),
                     "iterations_done_min_over_ranks": int(its.item()), "process_grid": [int(v) for v in g["info"][:3]]},
            "roofline": {"bound": "hbm", "achieved": nbytes / ms / 1e6 / world, "peak": peak, "unit": "GB/s",
                         "frac": nbytes / ms / 1e6 / world / peak, "traffic": None, "peak_source": peak_src,
                         "kernel": "k_cg_p + k_stream<EPI_DOT,HALO> + k_allreduce + k_cg_r + k_allreduce",
                         "note": "per-GPU share; bytes = MatMult algorithmic bytes + 10 vector passes"},
            "cpu_baseline": None,
            "e2e": {"value": nbytes / ms / 1e6, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "device-resident solve; see the matmult workload for the host-vector path"},
            "gpu_launches": int(res.launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    M.destroy()
    dist.destroy_process_group()
    return 0
