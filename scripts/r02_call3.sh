#!/bin/bash
# round 2, third GPU call (2 GPUs): the multi-GPU tests, the 2-rank benches, the distributed multigrid, host topology
mkdir -p gpurun_out
O=gpurun_out
{ nvidia-smi topo -m; echo; lscpu | head -25; echo; cat /sys/devices/system/node/online; grep -i "allowed" /proc/self/status; nproc; free -g | head -2;
  for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done; } > $O/r02c_topology.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02c_pytest.log
tail -6 $O/r02c_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 5 > $O/r02c_bench_n2.json 2> $O/r02c_bench_n2.err; echo "bench n2 rc=$?"
B200_PDL=0 timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 5 > $O/r02c_bench_n2_nopdl.json 2> $O/r02c_bench_n2_nopdl.err; echo "bench n2 nopdl rc=$?"
timeout 600 $TR bench.py --gpus 2 --workload cg --steps 200 --warmup 5 > $O/r02c_bench_cg_n2.json 2> $O/r02c_bench_cg_n2.err; echo "cg n2 rc=$?"
timeout 900 $TR scripts/dgamg_worker.py 200 > $O/r02c_dgamg_200_n2.log 2>&1; echo "dgamg rc=$?"
timeout 300 $TR tests/mpiaij_stress_worker.py 100 10000 > $O/r02c_stress_n2.log 2>&1; echo "stress rc=$?"
for f in bench_n2 bench_n2_nopdl bench_cg_n2; do echo "== $f"; python - <<PY
import json
try:
    d=json.loads(open("$O/r02c_$f.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"].get("frac"), d["e2e"].get("ms_per_step"), d.get("plan"), d.get("nccl_halo"))
except Exception as e:
    print("no line", e); print(open("$O/r02c_$f.err").read()[-1500:])
PY
done
tail -3 $O/r02c_dgamg_200_n2.log; tail -2 $O/r02c_stress_n2.log; cat $O/r02c_topology.log | head -40
