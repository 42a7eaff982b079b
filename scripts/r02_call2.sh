#!/bin/bash
# round 2, second GPU call: PDL stream kernel, fused CG, bench workloads
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02b_pytest.log
tail -15 $O/r02b_pytest.log
timeout 600 python bench.py --steps 50 --warmup 5 > $O/r02b_bench.json 2> $O/r02b_bench.err; echo "bench rc=$?"
B200_PDL=0 timeout 600 python bench.py --steps 50 --warmup 5 > $O/r02b_bench_nopdl.json 2> $O/r02b_bench_nopdl.err; echo "bench nopdl rc=$?"
timeout 600 python bench.py --workload cg --steps 200 --warmup 5 > $O/r02b_bench_cg.json 2> $O/r02b_bench_cg.err; echo "cg rc=$?"
timeout 600 python bench.py --workload stencil27 --steps 50 --warmup 5 > $O/r02b_bench_s27.json 2> $O/r02b_bench_s27.err; echo "s27 rc=$?"
timeout 900 python bench.py --workload powerlaw --steps 20 --warmup 5 > $O/r02b_bench_pl.json 2> $O/r02b_bench_pl.err; echo "pl rc=$?"
for f in bench bench_nopdl bench_cg bench_s27 bench_pl; do echo "== $f"; python - <<PY
import json
try:
    d=json.loads(open("$O/r02b_$f.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"].get("frac"), d["e2e"].get("ms_per_step"), d.get("plan"), d.get("transpose"))
except Exception as e:
    print("no line", e); print(open("$O/r02b_$f.err").read()[-1500:])
PY
done
