#!/bin/bash
# launch list (device time per launch) of the CG+GAMG solve at 300^3: which level costs what
mkdir -p gpurun_out
O=gpurun_out
CMD="petsc-openacc_b200/bin/ksp_poisson -config petsc-openacc_b200/host/configs/solver_cg_gamg.info -da_grid_x 300 -da_grid_y 300 -da_grid_z 300 -b200_json 1 -b200_solve_repeat 2"
$CMD > $O/r02j_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size --clock-control none --csv --log-file $O/r02j_gamg_launches.csv $CMD > $O/r02j_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $O/r02j_plain.log; wc -l $O/r02j_gamg_launches.csv
