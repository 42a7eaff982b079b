#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mpiaij.py tests/test_gpu_parity.py tests/test_dgamg.py -m gpu -q -x > $O/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02g_pytest.log; tail -4 $O/r02g_pytest.log
timeout 600 python scripts/probe_fused2.py 8 300 400 > $O/r02g_probe_fused_8.log 2>&1; echo "rc=$?"; cat $O/r02g_probe_fused_8.log
B200_MPIAIJ_SCHED=0 timeout 600 python scripts/probe_fused2.py 8 300 400 > $O/r02g_probe_fused_8_nosched.log 2>&1; echo "rc=$?"; cat $O/r02g_probe_fused_8_nosched.log
timeout 600 python scripts/probe_fused2.py 2 300 200 > $O/r02g_probe_fused_2.log 2>&1; echo "rc=$?"; cat $O/r02g_probe_fused_2.log
timeout 600 python scripts/probe_powerlaw2.py > $O/r02g_probe_powerlaw2.log 2>&1; echo "rc=$?"; tail -2 $O/r02g_probe_powerlaw2.log
