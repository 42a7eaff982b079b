#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r02l_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02l_pytest.log
tail -5 $O/r02l_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 5 > $O/r02l_bench_n2.json 2> $O/r02l_bench_n2.err; echo "bench n2 rc=$?"
B200_MPIAIJ_HOST_PIPELINE=0 timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 5 > $O/r02l_bench_n2_nopipe.json 2> $O/r02l_bench_n2_nopipe.err; echo "bench n2 nopipe rc=$?"
timeout 300 $TR tests/mpiaij_stress_worker.py 100 10000 > $O/r02l_stress_n2.log 2>&1; echo "stress rc=$?"
for f in bench_n2 bench_n2_nopipe; do echo "== $f"; python - <<PY
import json
try:
    d=json.loads(open("$O/r02l_$f.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"].get("frac"), d["e2e"], d.get("plan"))
except Exception as e:
    print("no line", e); print(open("$O/r02l_$f.err").read()[-1500:])
PY
done
tail -2 $O/r02l_stress_n2.log
