// ksp_poisson.cpp -- driver with the reference's command line and log format
// (src/main_ksp.cpp:70-142): -config <options file>, -da_grid_x/y/z, three timed phases, and the
// five-line report that scripts/generate_plots.py:87-90 parses.
#include <petscksp.h>
#include <petsctime.h>

#include "poisson_system.h"

int main(int argc, char **argv)
{
  PetscErrorCode     ierr;
  DM                 da;
  DMDALocalInfo      info;
  Vec                lhs, rhs, exact;
  Mat                A;
  KSP                ksp;
  KSPConvergedReason reason;
  PetscInt           its;
  PetscReal          res, linf;
  PetscLogDouble     t0, t1, t2, t3;
  char               config[PETSC_MAX_PATH_LEN];
  const PetscInt     nx = -100, ny = -100, nz = -100;  // defaults, overridden by -da_grid_*

  ierr = MPI_Init(&argc, &argv);CHKERRQ(ierr);
  ierr = PetscInitialize(&argc, &argv, nullptr, nullptr);CHKERRQ(ierr);
  ierr = PetscOptionsGetString(nullptr, nullptr, "-config", config, PETSC_MAX_PATH_LEN, nullptr);CHKERRQ(ierr);
  ierr = PetscOptionsInsertFile(PETSC_COMM_WORLD, nullptr, config, PETSC_FALSE);CHKERRQ(ierr);

  ierr = PetscTime(&t0);CHKERRQ(ierr);
  ierr = createSystem(nx, ny, nz, da, A, lhs, rhs, exact);CHKERRQ(ierr);
  ierr = DMDAGetLocalInfo(da, &info);CHKERRQ(ierr);
  ierr = PetscTime(&t1);CHKERRQ(ierr);

  ierr = KSPCreate(PETSC_COMM_WORLD, &ksp);CHKERRQ(ierr);
  ierr = KSPSetOperators(ksp, A, A);CHKERRQ(ierr);
  ierr = KSPSetType(ksp, KSPCG);CHKERRQ(ierr);
  ierr = KSPSetReusePreconditioner(ksp, PETSC_TRUE);CHKERRQ(ierr);
  ierr = KSPSetFromOptions(ksp);CHKERRQ(ierr);
  ierr = KSPSetUp(ksp);CHKERRQ(ierr);
  ierr = PetscTime(&t2);CHKERRQ(ierr);

  ierr = KSPSolve(ksp, rhs, lhs);CHKERRQ(ierr);
  ierr = PetscTime(&t3);CHKERRQ(ierr);

  ierr = KSPGetConvergedReason(ksp, &reason);CHKERRQ(ierr);
  if (reason < 0) SETERRQ1(PETSC_COMM_WORLD, PETSC_ERR_CONV_FAILED, "Diverger reason: %d\n", reason);
  ierr = KSPGetIterationNumber(ksp, &its);CHKERRQ(ierr);
  ierr = KSPGetResidualNorm(ksp, &res);CHKERRQ(ierr);
  ierr = VecAXPY(lhs, -1.0, exact);CHKERRQ(ierr);
  ierr = VecNorm(lhs, NORM_INFINITY, &linf);CHKERRQ(ierr);

  ierr = PetscPrintf(PETSC_COMM_WORLD,
                     "[Nx, Ny, Nz]: [%d, %d, %d]\nNumber of iterations: %d\nL2 norm of final residual: %f\n"
                     "Maximum norm of error: %f\nTime [init, create solver, solve]: [%f, %f, %f]\n",
                     info.mx, info.my, info.mz, its, res, linf, t1 - t0, t2 - t1, t3 - t2);CHKERRQ(ierr);

  ierr = KSPDestroy(&ksp);CHKERRQ(ierr);
  ierr = destroySystem(da, A, lhs, rhs, exact);CHKERRQ(ierr);
  ierr = PetscFinalize();CHKERRQ(ierr);
  ierr = MPI_Finalize();CHKERRQ(ierr);
  return ierr;
}
