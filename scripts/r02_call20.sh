#!/bin/bash
# round 2: the state at the end of the kernel work -- full GPU tests, the bench lines, ncu evidence (1 GPU)
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r02s_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02s_pytest.log; tail -4 $O/r02s_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02s_bench.json 2> $O/r02s_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02s_bench_ref.json 2> $O/r02s_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --workload cg --steps 200 --warmup 5 > $O/r02s_bench_cg.json 2> $O/r02s_bench_cg.err; echo "cg rc=$?"
timeout 600 python bench.py --workload stencil27 --steps 50 --warmup 5 > $O/r02s_bench_s27.json 2> $O/r02s_bench_s27.err; echo "s27 rc=$?"
timeout 900 python bench.py --workload powerlaw --steps 20 --warmup 5 > $O/r02s_bench_pl.json 2> $O/r02s_bench_pl.err; echo "pl rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02s_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02s_smoke.log
# ncu: launch list of the bench command, then full captures of the stream kernel and of one CG iteration
python bench.py --steps 2 --warmup 1 > $O/r02s_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02s_launches.csv python bench.py --steps 2 --warmup 1 > $O/r02s_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python scripts/profile_case.py poisson300 6 > $O/r02s_plain_p300.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_stream -s 3 -c 1 -f -o $O/r02s_k_stream python scripts/profile_case.py poisson300 6 > $O/r02s_ncu_p300.log 2>&1
echo "ncu k_stream rc=$?"
python scripts/profile_case.py cg300 8 > $O/r02s_plain_cg300.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_cg_p|k_cg_r|k_stream" -s 12 -c 3 -f -o $O/r02s_cg python scripts/profile_case.py cg300 8 > $O/r02s_ncu_cg300.log 2>&1
echo "ncu cg rc=$?"
python scripts/profile_case.py stencil27 6 > $O/r02s_plain_s27.log 2>&1 && \
ncu --set full --clock-control none -k regex:k_stream -s 3 -c 1 -f -o $O/r02s_k_stream_s27 python scripts/profile_case.py stencil27 6 > $O/r02s_ncu_s27.log 2>&1
echo "ncu s27 rc=$?"
for f in bench bench_cg bench_s27 bench_pl; do python - <<PY
import json
try:
    d=json.loads(open("$O/r02s_$f.json").read().strip().splitlines()[-1])
    print("$f", {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"].get("frac"), d["e2e"].get("ms_per_step"), d.get("transpose"))
except Exception as e:
    print("$f no line", e); print(open("$O/r02s_$f.err").read()[-800:])
PY
done
