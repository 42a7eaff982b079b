#!/bin/bash
# round 2, first GPU call: everything that was written without a GPU, then the baselines of this round
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02a_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > $O/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02a_pytest.log
tail -5 $O/r02a_pytest.log
timeout 600 python scripts/sweep.py 300 7 quick > $O/r02a_sweep_300.log 2>&1; echo "rc=$?" >> $O/r02a_sweep_300.log
timeout 600 python scripts/sweep.py 200 27 quick > $O/r02a_sweep_27_200.log 2>&1; echo "rc=$?" >> $O/r02a_sweep_27_200.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02a_bench.json 2> $O/r02a_bench.err; echo "bench rc=$?"
timeout 900 petsc-openacc_b200/bin/ksp_poisson -config petsc-openacc_b200/host/configs/solver_cg_gamg.info \
   -da_grid_x 300 -da_grid_y 300 -da_grid_z 300 -pc_gamg_b200_view 1 -b200_json 1 -b200_solve_repeat 3 > $O/r02a_gamg_300.log 2>&1; echo "gamg rc=$?"
tail -3 $O/r02a_sweep_300.log $O/r02a_sweep_27_200.log; tail -c 600 $O/r02a_bench.json; tail -12 $O/r02a_gamg_300.log
