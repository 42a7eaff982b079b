#!/bin/bash
# 8 GPUs: push concentration x schedule charge, then the stress test with the final protocol
mkdir -p gpurun_out
O=gpurun_out
P=29600
trun() { lim=$1; n=$2; shift 2; P=$((P+1)); timeout $lim python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P "$@"; }
show() { python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r02r_{tag}.json").read().strip().splitlines()[-1])
    print(tag, {k: d[k] for k in ("value", "ms_per_step")}, d["e2e"].get("ms_per_step"), d["plan"]["parity_vs_oracle"])
except Exception as e:
    print(tag, "no line", e); print(open(f"gpurun_out/r02r_{tag}.err").read()[-800:])
PY
}
trun 600 8 bench.py --gpus 8 --steps 400 --warmup 10 > $O/r02r_bench_n8.json 2> $O/r02r_bench_n8.err; show bench_n8
B200_MPIAIJ_PUSH_CTAS=64 B200_MPIAIJ_PUSH_CHARGE_TENTHS=25 trun 600 8 bench.py --gpus 8 --steps 400 --warmup 10 > $O/r02r_p64_c25.json 2> $O/r02r_p64_c25.err; show p64_c25
B200_MPIAIJ_PUSH_CTAS=32 B200_MPIAIJ_PUSH_CHARGE_TENTHS=60 trun 600 8 bench.py --gpus 8 --steps 400 --warmup 10 > $O/r02r_p32_c60.json 2> $O/r02r_p32_c60.err; show p32_c60
B200_MPIAIJ_PUSH_CTAS=16 B200_MPIAIJ_PUSH_CHARGE_TENTHS=80 trun 600 8 bench.py --gpus 8 --steps 400 --warmup 10 > $O/r02r_p16_c80.json 2> $O/r02r_p16_c80.err; show p16_c80
B200_MPIAIJ_PUSH_CTAS=148 B200_MPIAIJ_PUSH_CHARGE_TENTHS=10 trun 600 8 bench.py --gpus 8 --steps 400 --warmup 10 > $O/r02r_p148_c10.json 2> $O/r02r_p148_c10.err; show p148_c10
trun 300 8 tests/mpiaij_stress_worker.py 120 10000 > $O/r02r_stress_n8.log 2>&1; echo "stress rc=$?"; tail -1 $O/r02r_stress_n8.log
