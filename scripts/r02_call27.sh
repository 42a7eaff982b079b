#!/bin/bash
# ncu of the two kernels of this round that had no capture yet: k_wmerge (power law) and k_stream<HALO> (rank 0 of 8, one GPU)
mkdir -p gpurun_out
O=gpurun_out
python scripts/profile_case.py powerlaw 6 > $O/r02x_plain_pl.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_wmerge -s 3 -c 1 -f -o $O/r02x_k_wmerge python scripts/profile_case.py powerlaw 6 > $O/r02x_ncu_pl.log 2>&1
echo "ncu wmerge rc=$?"
python scripts/profile_fused.py > $O/r02x_plain_fused.log 2>&1 && \
ncu --set full --clock-control none -k regex:k_stream -s 10 -c 1 -f -o $O/r02x_k_stream_halo python scripts/profile_fused.py > $O/r02x_ncu_fused.log 2>&1
echo "ncu halo rc=$?"; tail -2 $O/r02x_ncu_fused.log
