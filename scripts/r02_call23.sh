#!/bin/bash
# round 2, last multi-GPU call: the scaling lines with the final kernel
mkdir -p gpurun_out
O=gpurun_out
P=29620
trun() { lim=$1; n=$2; shift 2; P=$((P+1)); timeout $lim python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P "$@"; }
trun 300 8 bench.py --gpus 8 --steps 400 --warmup 10 > $O/r02t_bench_n8.json 2> $O/r02t_bench_n8.err; echo "n8 rc=$?"
trun 300 4 bench.py --gpus 4 --steps 400 --warmup 10 > $O/r02t_bench_n4.json 2> $O/r02t_bench_n4.err; echo "n4 rc=$?"
for f in bench_n8 bench_n4; do python - "$f" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r02t_{tag}.json").read().strip().splitlines()[-1])
    print(tag, {k: d[k] for k in ("value", "ms_per_step")}, d["e2e"].get("ms_per_step"), d["plan"]["parity_vs_oracle"])
except Exception as e:
    print(tag, "no line", e); print(open(f"gpurun_out/r02t_{tag}.err").read()[-800:])
PY
done
