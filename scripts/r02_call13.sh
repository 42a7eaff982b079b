#!/bin/bash
# round 2: the 8-GPU call -- scaling bench, attribution, distributed CG, halo stress, distributed multigrid
mkdir -p gpurun_out
O=gpurun_out
P=29520
trun() { lim=$1; n=$2; shift 2; P=$((P+1)); timeout $lim python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P "$@"; }
nvidia-smi topo -m > $O/r02m_topology.log 2>&1; nproc >> $O/r02m_topology.log
trun 600 8 bench.py --gpus 8 --steps 200 --warmup 10 > $O/r02m_bench_n8.json 2> $O/r02m_bench_n8.err; echo "bench n8 rc=$?"
B200_PDL=0 B200_MPIAIJ_SCHED=0 trun 600 8 bench.py --gpus 8 --steps 200 --warmup 10 > $O/r02m_bench_n8_r01like.json 2> $O/r02m_bench_n8_r01like.err; echo "bench n8 (no pdl, no sched) rc=$?"
trun 600 4 bench.py --gpus 4 --steps 200 --warmup 10 > $O/r02m_bench_n4.json 2> $O/r02m_bench_n4.err; echo "bench n4 rc=$?"
trun 600 8 bench.py --gpus 8 --workload cg --steps 300 --warmup 10 > $O/r02m_bench_cg_n8.json 2> $O/r02m_bench_cg_n8.err; echo "cg n8 rc=$?"
trun 300 8 tests/mpiaij_stress_worker.py 120 10000 > $O/r02m_stress_n8.log 2>&1; echo "stress rc=$?"
trun 300 8 tests/mpiaij_cg_worker.py 300 1e-14 > $O/r02m_cg_solve_300_n8.log 2>&1; echo "cg solve rc=$?"
trun 900 8 scripts/dgamg_worker.py 300 > $O/r02m_dgamg_300_n8.log 2>&1; echo "dgamg rc=$?"
for f in bench_n8 bench_n8_r01like bench_n4 bench_cg_n8; do echo "== $f"; python - <<PY
import json
try:
    d=json.loads(open("$O/r02m_$f.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"].get("frac"), d["e2e"].get("ms_per_step"), d.get("plan"), d.get("nccl_halo"))
except Exception as e:
    print("no line", e); print(open("$O/r02m_$f.err").read()[-1500:])
PY
done
tail -2 $O/r02m_stress_n8.log; tail -2 $O/r02m_cg_solve_300_n8.log; tail -2 $O/r02m_dgamg_300_n8.log
