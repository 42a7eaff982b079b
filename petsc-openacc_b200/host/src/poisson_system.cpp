// poisson_system.cpp -- the reference's problem builder on the PETSc-shaped API:
// createSystem / destroySystem / generateRHS / generateExt / generateA / setRefPoint with the
// signatures of src/helper.h:32-96.  Same arithmetic as src/helper.cpp (cell-centred cosines,
// 1/(dx*dx) off-diagonals, Neumann diagonal accumulated over the present neighbours in the order
// x-,x+,y-,y+,z-,z+, reference point fixed through MatZeroRowsColumns), organised around a table
// of stencil offsets instead of per-entry assignments.
#include "poisson_system.h"

#include <cmath>

namespace {
// unparenthesised on purpose: src/helper.cpp:17-18 expands to exactly this product chain
constexpr double kTwoPi = 2.0 * 1.0 * M_PI;
inline double rhs_scale() { return -3.0 * 2.0 * 1.0 * M_PI * 2.0 * 1.0 * M_PI; }

struct Spacing { PetscScalar dx, dy, dz; };
Spacing spacing(const DMDALocalInfo &info) { return {1.0 / info.mx, 1.0 / info.my, 1.0 / info.mz}; }

// fills v[k][j][i] = scale * cos * cos * cos at cell centres (scale = 1: exact solution)
PetscErrorCode fill_cosines(const DM &grid, Vec &v, bool is_rhs)
{
  PetscErrorCode ierr;
  DMDALocalInfo  info;
  PetscScalar ***arr;
  ierr = DMDAGetLocalInfo(grid, &info);CHKERRQ(ierr);
  const Spacing h = spacing(info);
  ierr = DMDAVecGetArray(grid, v, &arr);CHKERRQ(ierr);
  for (int k = info.zs; k < info.zs + info.zm; ++k)
    for (int j = info.ys; j < info.ys + info.ym; ++j)
      for (int i = info.xs; i < info.xs + info.xm; ++i) {
        const double cx = std::cos(kTwoPi * (i + 0.5) * h.dx);
        const double cy = std::cos(kTwoPi * (j + 0.5) * h.dy);
        const double cz = std::cos(kTwoPi * (k + 0.5) * h.dz);
        arr[k][j][i] = is_rhs ? rhs_scale() * cx * cy * cz : cx * cy * cz;
      }
  ierr = DMDAVecRestoreArray(grid, v, &arr);CHKERRQ(ierr);
  return 0;
}
}  // namespace

extern "C" PetscErrorCode generateRHS(const DM &grid, Vec &rhs) { return fill_cosines(grid, rhs, true); }
extern "C" PetscErrorCode generateExt(const DM &grid, Vec &exact) { return fill_cosines(grid, exact, false); }

extern "C" PetscErrorCode generateA(const DM &grid, Mat &A)
{
  PetscErrorCode         ierr;
  DMDALocalInfo          info;
  ISLocalToGlobalMapping ltog;
  ierr = DMDAGetLocalInfo(grid, &info);CHKERRQ(ierr);
  ierr = DMGetLocalToGlobalMapping(grid, &ltog);CHKERRQ(ierr);
  const Spacing h = spacing(info);
  // entry 0 is the diagonal; 1..6 = x-, x+, y-, y+, z-, z+
  static const int off[7][3] = {{0, 0, 0}, {-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
  PetscReal        coef[7];
  coef[1] = coef[2] = 1.0 / (h.dx * h.dx);
  coef[3] = coef[4] = 1.0 / (h.dy * h.dy);
  coef[5] = coef[6] = 1.0 / (h.dz * h.dz);
  for (int k = info.zs; k < info.zs + info.zm; ++k)
    for (int j = info.ys; j < info.ys + info.ym; ++j)
      for (int i = info.xs; i < info.xs + info.xm; ++i) {
        PetscInt cols[7];
        for (int e = 0; e < 7; ++e) {
          MatStencil s;
          s.i = i + off[e][0]; s.j = j + off[e][1]; s.k = k + off[e][2]; s.c = 0;
          ierr = DMDAConvertToCell(grid, s, &cols[e]);CHKERRQ(ierr);
        }
        ierr = ISLocalToGlobalMappingApply(ltog, 7, cols, cols);CHKERRQ(ierr);
        coef[0] = 0.0;
        for (int e = 1; e < 7; ++e)
          if (cols[e] > -1) coef[0] -= coef[e];  // all-Neumann: only neighbours inside the domain
        ierr = MatSetValues(A, 1, &cols[0], 7, cols, coef, INSERT_VALUES);CHKERRQ(ierr);
      }
  ierr = MatAssemblyBegin(A, MAT_FINAL_ASSEMBLY);CHKERRQ(ierr);
  ierr = MatAssemblyEnd(A, MAT_FINAL_ASSEMBLY);CHKERRQ(ierr);
  return 0;
}

extern "C" PetscErrorCode setRefPoint(Mat &A, Vec &rhs, const Vec &exact)
{
  PetscErrorCode ierr;
  Vec            diag;
  PetscInt       n, row0[1] = {0};
  PetscReal      mean;
  ierr = MatCreateVecs(A, nullptr, &diag);CHKERRQ(ierr);
  ierr = MatGetDiagonal(A, diag);CHKERRQ(ierr);
  ierr = VecGetSize(diag, &n);CHKERRQ(ierr);
  ierr = VecSum(diag, &mean);CHKERRQ(ierr);
  mean /= double(n);
  ierr = MatZeroRowsColumns(A, 1, row0, mean, exact, rhs);CHKERRQ(ierr);
  ierr = VecDestroy(&diag);CHKERRQ(ierr);
  return 0;
}

extern "C" PetscErrorCode createSystem(const PetscInt &Nx, const PetscInt &Ny, const PetscInt &Nz, DM &da, Mat &A,
                                       Vec &lhs, Vec &rhs, Vec &exact)
{
  PetscErrorCode ierr;
  ierr = DMDACreate3d(PETSC_COMM_WORLD, DM_BOUNDARY_GHOSTED, DM_BOUNDARY_GHOSTED, DM_BOUNDARY_GHOSTED,
                      DMDA_STENCIL_STAR, Nx, Ny, Nz, PETSC_DECIDE, PETSC_DECIDE, PETSC_DECIDE, 1, 1, nullptr,
                      nullptr, nullptr, &da);CHKERRQ(ierr);
  ierr = DMSetMatType(da, MATAIJ);CHKERRQ(ierr);
  ierr = DMCreateGlobalVector(da, &lhs);CHKERRQ(ierr);
  ierr = DMCreateGlobalVector(da, &rhs);CHKERRQ(ierr);
  ierr = DMCreateGlobalVector(da, &exact);CHKERRQ(ierr);
  ierr = DMCreateMatrix(da, &A);CHKERRQ(ierr);
  ierr = VecSet(lhs, 0.0);CHKERRQ(ierr);
  ierr = generateRHS(da, rhs);CHKERRQ(ierr);
  ierr = generateExt(da, exact);CHKERRQ(ierr);
  ierr = generateA(da, A);CHKERRQ(ierr);
  ierr = setRefPoint(A, rhs, exact);CHKERRQ(ierr);
  return 0;
}

extern "C" PetscErrorCode destroySystem(DM &da, Mat &A, Vec &lhs, Vec &rhs, Vec &exact)
{
  PetscErrorCode ierr;
  ierr = VecDestroy(&exact);CHKERRQ(ierr);
  ierr = VecDestroy(&rhs);CHKERRQ(ierr);
  ierr = VecDestroy(&lhs);CHKERRQ(ierr);
  ierr = MatDestroy(&A);CHKERRQ(ierr);
  ierr = DMDestroy(&da);CHKERRQ(ierr);
  return 0;
}

// C entry for ctypes-driven tests: build the system for an N^3 grid and hand back the objects
extern "C" PetscErrorCode b200_create_poisson_system(PetscInt N, DM *da, Mat *A, Vec *lhs, Vec *rhs, Vec *exact)
{
  const PetscInt n = N;
  return createSystem(n, n, n, *da, *A, *lhs, *rhs, *exact);
}
extern "C" PetscErrorCode b200_destroy_poisson_system(DM *da, Mat *A, Vec *lhs, Vec *rhs, Vec *exact)
{
  return destroySystem(*da, *A, *lhs, *rhs, *exact);
}
