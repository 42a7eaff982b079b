#!/bin/bash
# round 2, closing 1-GPU call: everything green with the final library, the contract line, ncu of the final stream kernel
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r02v_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02v_pytest.log; tail -4 $O/r02v_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02v_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02v_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02v_bench.json 2> $O/r02v_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02v_bench_ref.json 2> $O/r02v_bench_ref.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 1 > $O/r02v_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02v_launches.csv python bench.py --steps 2 --warmup 1 > $O/r02v_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python scripts/profile_case.py poisson300 6 > $O/r02v_plain_p300.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_stream -s 3 -c 1 -f -o $O/r02v_k_stream python scripts/profile_case.py poisson300 6 > $O/r02v_ncu_p300.log 2>&1
echo "ncu k_stream rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02v_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["roofline"]["frac"], d["roofline"]["traffic"], d["e2e"]["ms_per_step"], d["plan"])
PY
