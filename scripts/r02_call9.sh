#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python scripts/probe_fused2.py 2 300 200 > $O/r02i_$tag.log 2>&1; echo "== $tag: $@"; grep -E "A only|fused" $O/r02i_$tag.log | head -2; }
run base X=1
run noghost B200_MPIAIJ_PROBE_NOGHOST=1
run nopdl B200_PDL=0
run nopdl_noghost B200_PDL=0 B200_MPIAIJ_PROBE_NOGHOST=1
run nosched_noghost B200_MPIAIJ_SCHED=0 B200_MPIAIJ_PROBE_NOGHOST=1
