#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
P=29580
trun() { lim=$1; n=$2; shift 2; P=$((P+1)); timeout $lim python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P "$@"; }
timeout 600 python -m pytest tests/test_gpu_mpiaij.py -m gpu -q -x > $O/r02q_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02q_pytest.log; tail -3 $O/r02q_pytest.log
trun 600 4 bench.py --gpus 4 --steps 400 --warmup 10 > $O/r02q_bench_n4.json 2> $O/r02q_bench_n4.err; echo "bench n4 rc=$?"
B200_MPIAIJ_PUSH_CTAS=4 trun 600 4 bench.py --gpus 4 --steps 400 --warmup 10 > $O/r02q_bench_n4_p4.json 2> $O/r02q_bench_n4_p4.err; echo "bench n4 p4 rc=$?"
B200_MPIAIJ_PUSH_CTAS=148 B200_MPIAIJ_PUSH_CHARGE_TENTHS=5 trun 600 4 bench.py --gpus 4 --steps 400 --warmup 10 > $O/r02q_bench_n4_p148.json 2> $O/r02q_bench_n4_p148.err; echo "bench n4 p148 rc=$?"
B200_MPIAIJ_PROBE_NOPUSH=1 B200_MPIAIJ_PROBE_NOWAIT=1 trun 600 4 bench.py --gpus 4 --steps 400 --warmup 10 > $O/r02q_bench_n4_nopush.json 2> $O/r02q_bench_n4_nopush.err; echo "bench n4 nopush rc=$?"
trun 300 4 tests/mpiaij_stress_worker.py 100 10000 > $O/r02q_stress_n4.log 2>&1; echo "stress rc=$?"
for f in bench_n4 bench_n4_p4 bench_n4_p148 bench_n4_nopush; do echo "== $f"; python - <<PY
import json
try:
    d=json.loads(open("$O/r02q_$f.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step")}, d["plan"]["parity_vs_oracle"])
except Exception as e:
    print("no line", e); print(open("$O/r02q_$f.err").read()[-1500:])
PY
done
tail -1 $O/r02q_stress_n4.log
