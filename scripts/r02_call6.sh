#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_shape_sweep.py tests/test_gpu_fullsize.py tests/test_gpu_mpiaij.py tests/test_gpu_cg.py -m gpu -q -x > $O/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02f_pytest.log; tail -4 $O/r02f_pytest.log
timeout 600 python scripts/probe_powerlaw2.py > $O/r02f_probe_powerlaw2.log 2>&1; echo "rc=$?"; cat $O/r02f_probe_powerlaw2.log
timeout 600 python scripts/probe_fused2.py 8 300 400 > $O/r02f_probe_fused_8.log 2>&1; echo "rc=$?"; cat $O/r02f_probe_fused_8.log
