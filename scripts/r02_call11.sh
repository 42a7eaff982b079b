#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python scripts/probe_longrows.py > $O/r02k_probe_longrows.log 2>&1; echo "rc=$?"; cat $O/r02k_probe_longrows.log
CMD="petsc-openacc_b200/bin/ksp_poisson -config petsc-openacc_b200/host/configs/solver_cg_gamg.info -da_grid_x 200 -da_grid_y 200 -da_grid_z 200 -b200_json 1 -b200_solve_repeat 4"
$CMD > $O/r02k_gamg200_pdl1.log 2>&1; tail -1 $O/r02k_gamg200_pdl1.log
B200_PDL=0 $CMD > $O/r02k_gamg200_pdl0.log 2>&1; tail -1 $O/r02k_gamg200_pdl0.log
$CMD > $O/r02k_gamg200_pdl1b.log 2>&1; tail -1 $O/r02k_gamg200_pdl1b.log
