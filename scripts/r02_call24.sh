#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_shape_sweep.py tests/test_gpu_fullsize.py tests/test_gpu_cg.py tests/test_pcgamg.py -m gpu -q -x > $O/r02u_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02u_pytest.log; tail -4 $O/r02u_pytest.log
timeout 600 python bench.py --steps 50 --warmup 5 > $O/r02u_bench.json 2> $O/r02u_bench.err; echo "bench rc=$?"
B200_ROWLEN8=0 timeout 600 python bench.py --steps 50 --warmup 5 > $O/r02u_bench_norl8.json 2> $O/r02u_bench_norl8.err; echo "bench norl8 rc=$?"
timeout 600 python bench.py --workload cg --steps 200 --warmup 5 > $O/r02u_bench_cg.json 2> $O/r02u_bench_cg.err; echo "cg rc=$?"
timeout 600 python bench.py --workload stencil27 --steps 50 --warmup 5 > $O/r02u_bench_s27.json 2> $O/r02u_bench_s27.err; echo "s27 rc=$?"
for f in bench bench_norl8 bench_cg bench_s27; do python - "$f" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r02u_{tag}.json").read().strip().splitlines()[-1])
    print(tag, {k: d[k] for k in ("value", "ms_per_step")}, d["e2e"].get("ms_per_step"), d["plan"].get("parity_vs_oracle"))
except Exception as e:
    print(tag, "no line", e); print(open(f"gpurun_out/r02u_{tag}.err").read()[-800:])
PY
done
