#!/bin/bash
# ncu of the power-law kernels (full set, source counters), after a plain run of the same command
mkdir -p gpurun_out
O=gpurun_out
CMD="python scripts/profile_case.py powerlaw 6"
$CMD > $O/r02e_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_mergex|k_longrow" -s 6 -c 2 -o $O/r02e_powerlaw -f $CMD > $O/r02e_ncu.log 2>&1
echo "ncu rc=$?"; tail -5 $O/r02e_plain.log; tail -5 $O/r02e_ncu.log
