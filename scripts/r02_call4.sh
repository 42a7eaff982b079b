#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python scripts/probe_fused2.py 8 300 400 > $O/r02d_probe_fused_8.log 2>&1; echo "rc=$?"
B200_PDL=0 timeout 600 python scripts/probe_fused2.py 8 300 400 > $O/r02d_probe_fused_8_nopdl.log 2>&1; echo "rc=$?"
timeout 900 python scripts/probe_powerlaw_l2.py > $O/r02d_probe_powerlaw_l2.log 2>&1; echo "rc=$?"
cat $O/r02d_probe_fused_8.log $O/r02d_probe_fused_8_nopdl.log $O/r02d_probe_powerlaw_l2.log
