// poisson_system.h -- same interface as the reference's src/helper.h:32-96.
#pragma once
#include <petscsys.h>
#include <petscdmda.h>
#include <petscvec.h>
#include <petscmat.h>

extern "C" PetscErrorCode createSystem(const PetscInt &Nx, const PetscInt &Ny, const PetscInt &Nz, DM &da, Mat &A,
                                       Vec &lhs, Vec &rhs, Vec &exact);
extern "C" PetscErrorCode destroySystem(DM &da, Mat &A, Vec &lhs, Vec &rhs, Vec &exact);
extern "C" PetscErrorCode generateRHS(const DM &grid, Vec &rhs);
extern "C" PetscErrorCode generateExt(const DM &grid, Vec &exact);
extern "C" PetscErrorCode generateA(const DM &grid, Mat &A);
extern "C" PetscErrorCode setRefPoint(Mat &A, Vec &rhs, const Vec &exact);
extern "C" PetscErrorCode b200_create_poisson_system(PetscInt N, DM *da, Mat *A, Vec *lhs, Vec *rhs, Vec *exact);
extern "C" PetscErrorCode b200_destroy_poisson_system(DM *da, Mat *A, Vec *lhs, Vec *rhs, Vec *exact);
