#!/bin/bash
# 8 GPUs: where do the microseconds between the single-GPU proxy of the per-rank kernel and the real 8-rank step go?
mkdir -p gpurun_out
O=gpurun_out
P=29560
trun() { lim=$1; n=$2; shift 2; P=$((P+1)); timeout $lim python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P "$@"; }
trun 600 8 bench.py --gpus 8 --steps 400 --warmup 10 > $O/r02p_bench_n8.json 2> $O/r02p_bench_n8.err; echo "default rc=$?"
B200_MPIAIJ_PROBE_NOPUSH=1 B200_MPIAIJ_PROBE_NOWAIT=1 trun 600 8 bench.py --gpus 8 --steps 400 --warmup 10 > $O/r02p_bench_n8_nopush_nowait.json 2> $O/r02p_bench_n8_nopush_nowait.err; echo "nopush nowait rc=$?"
B200_MPIAIJ_PROBE_NOWAIT=1 trun 600 8 bench.py --gpus 8 --steps 400 --warmup 10 > $O/r02p_bench_n8_nowait.json 2> $O/r02p_bench_n8_nowait.err; echo "nowait rc=$?"
B200_MPIAIJ_PROBE_NOPUSH=1 B200_MPIAIJ_PROBE_NOWAIT=1 B200_MPIAIJ_PROBE_NOGHOST=1 trun 600 8 bench.py --gpus 8 --steps 400 --warmup 10 > $O/r02p_bench_n8_nothing.json 2> $O/r02p_bench_n8_nothing.err; echo "nothing rc=$?"
for f in bench_n8 bench_n8_nopush_nowait bench_n8_nowait bench_n8_nothing; do echo "== $f"; python - <<PY
import json
try:
    d=json.loads(open("$O/r02p_$f.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step")}, d["roofline"].get("frac"), d["e2e"].get("ms_per_step"), d["plan"]["parity_vs_oracle"])
except Exception as e:
    print("no line", e); print(open("$O/r02p_$f.err").read()[-1500:])
PY
done
