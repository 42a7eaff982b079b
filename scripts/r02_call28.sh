#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python scripts/profile_fused.py > $O/r02y_plain_fused.log 2>&1 && \
ncu --set full --clock-control none -k regex:k_stream -s 2 -c 1 -f -o $O/r02y_k_stream_halo python scripts/profile_fused.py > $O/r02y_ncu_fused.log 2>&1
echo "ncu halo rc=$?"
timeout 600 python scripts/probe_powerlaw2.py 2>&1 | head -3 > $O/r02y_pl_default.log; cat $O/r02y_pl_default.log
B200_L2_PERSIST=1 B200_L2_PERSIST_VERBOSE=1 timeout 600 python scripts/probe_powerlaw2.py 2>&1 | grep -v "^\[b200\] L2 window" | head -3 > $O/r02y_pl_persist.log; cat $O/r02y_pl_persist.log
B200_L2_PERSIST=1 B200_L2_PERSIST_VERBOSE=1 timeout 600 python scripts/profile_case.py powerlaw 2 2>&1 | grep "L2 window" | head -2
