#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mpiaij.py -m gpu -q -x > $O/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02h_pytest.log; tail -3 $O/r02h_pytest.log
for rep in 1 2; do
timeout 600 python scripts/probe_fused2.py 8 300 400 > $O/r02h_probe_fused_8_$rep.log 2>&1; echo "rc=$?"; cat $O/r02h_probe_fused_8_$rep.log
B200_MPIAIJ_SCHED=0 timeout 600 python scripts/probe_fused2.py 8 300 400 > $O/r02h_probe_fused_8_nosched_$rep.log 2>&1; echo "rc=$?"; cat $O/r02h_probe_fused_8_nosched_$rep.log
done
timeout 600 python scripts/probe_fused2.py 2 300 200 > $O/r02h_probe_fused_2.log 2>&1; echo "rc=$?"; cat $O/r02h_probe_fused_2.log
B200_MPIAIJ_SCHED=0 timeout 600 python scripts/probe_fused2.py 2 300 200 > $O/r02h_probe_fused_2_nosched.log 2>&1; echo "rc=$?"; cat $O/r02h_probe_fused_2_nosched.log
