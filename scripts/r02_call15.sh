#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
P=29540
trun() { lim=$1; n=$2; shift 2; P=$((P+1)); timeout $lim python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P "$@"; }
trun 600 4 bench.py --gpus 4 --steps 200 --warmup 10 > $O/r02o_bench_n4.json 2> $O/r02o_bench_n4.err; echo "bench n4 rc=$?"
trun 600 2 bench.py --gpus 2 --steps 200 --warmup 10 > $O/r02o_bench_n2.json 2> $O/r02o_bench_n2.err; echo "bench n2 rc=$?"
trun 300 4 tests/mpiaij_stress_worker.py 100 10000 > $O/r02o_stress_n4.log 2>&1; echo "stress rc=$?"
for f in bench_n4 bench_n2; do echo "== $f"; python - <<PY
import json
try:
    d=json.loads(open("$O/r02o_$f.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"].get("frac"), d["e2e"].get("ms_per_step"), d["plan"]["parity_vs_oracle"])
except Exception as e:
    print("no line", e); print(open("$O/r02o_$f.err").read()[-1500:])
PY
done
tail -1 $O/r02o_stress_n4.log
