#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mpiaij.py tests/test_gpu_cg.py -m gpu -q -x > $O/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02n_pytest.log; tail -3 $O/r02n_pytest.log
timeout 600 python scripts/probe_fused2.py 8 300 400 > $O/r02n_probe_fused_8.log 2>&1; grep -E "A only|fused" $O/r02n_probe_fused_8.log
timeout 600 python scripts/probe_fused2.py 2 300 200 > $O/r02n_probe_fused_2.log 2>&1; grep -E "A only|fused" $O/r02n_probe_fused_2.log
