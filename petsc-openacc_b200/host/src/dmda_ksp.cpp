// dmda_ksp.cpp -- the DMDA and KSP calls of the reference application (src/helper.cpp,
// src/main_ksp.cpp) on one rank.
//
// DMDA: 3-D structured grid, dof 1, stencil width 1, ghosted boundaries -- what
// DMDACreate3d(..., DM_BOUNDARY_GHOSTED x3, DMDA_STENCIL_STAR, Nx,Ny,Nz, PETSC_DECIDE x3, 1, 1, ...)
// at src/helper.cpp:31-36 builds.  Local (ghosted) numbering, the local-to-global map with -1 on
// cells outside the domain, and DMDAVecGetArray's [k][j][i] view follow PETSc [P376].
// KSP: KSPCG (KSPSolve_CG [P376]: left preconditioning, preconditioned residual norm,
// KSPConvergedDefault) with PCJACOBI, PCNONE or PCGAMG (src/pcgamg.cpp); the Lanczos coefficients
// of the CG recurrence can be recorded for the eigenvalue estimate PCGAMG needs.
#include <cmath>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "b200_aij.h"
#include "pc_impl.h"
// (after b200_aij.h: the symbols header only forward-declares Mat/Vec when PETSc types are absent)
#include "../../../include/b200_petsc_symbols.h"
#include "../../../include/b200_seqaij.h"

struct _p_ISLocalToGlobalMapping {
  std::vector<PetscInt> map;  // local ghosted index -> global index or -1
};

struct _p_DM {
  PetscInt M, N, P;                    // global cells
  PetscInt xs, ys, zs, xm, ym, zm;     // owned box (one rank: the whole grid)
  PetscInt gxs, gys, gzs, gxm, gym, gzm;
  PetscInt dof, sw;
  DMBoundaryType bx, by, bz;
  DMDAStencilType st;
  std::string mattype;
  _p_ISLocalToGlobalMapping ltog;
  std::map<Vec, std::vector<void *>> views;  // DMDAVecGetArray allocations
};

extern "C" PetscErrorCode DMDACreate3d(MPI_Comm comm, DMBoundaryType bx, DMBoundaryType by, DMBoundaryType bz,
                                       DMDAStencilType st, PetscInt M, PetscInt N, PetscInt P, PetscInt m, PetscInt n,
                                       PetscInt p, PetscInt dof, PetscInt s, const PetscInt lx[], const PetscInt ly[],
                                       const PetscInt lz[], DM *da)
{
  PetscErrorCode ierr;
  (void)lx; (void)ly; (void)lz;
  if (dof != 1 || s != 1) SETERRQ(comm, PETSC_ERR_SUP, "DMDA: only dof = 1, stencil width = 1");
  if ((m != PETSC_DECIDE && m != 1) || (n != PETSC_DECIDE && n != 1) || (p != PETSC_DECIDE && p != 1))
    SETERRQ(comm, PETSC_ERR_SUP, "DMDA: one process per DM here (multi-GPU goes through b200_mpiaij.h)");
  // negative sizes are defaults that -da_grid_x/y/z may override [P376] (src/main_ksp.cpp:33-35)
  if (M < 0) { M = -M; ierr = PetscOptionsGetInt(NULL, NULL, "-da_grid_x", &M, NULL);CHKERRQ(ierr); }
  if (N < 0) { N = -N; ierr = PetscOptionsGetInt(NULL, NULL, "-da_grid_y", &N, NULL);CHKERRQ(ierr); }
  if (P < 0) { P = -P; ierr = PetscOptionsGetInt(NULL, NULL, "-da_grid_z", &P, NULL);CHKERRQ(ierr); }
  if (M < 1 || N < 1 || P < 1) SETERRQ(comm, PETSC_ERR_ARG_OUTOFRANGE, "DMDA: grid sizes must be positive");
  if ((long long)M * N * P > 2147483647LL / 8) SETERRQ(comm, PETSC_ERR_ARG_OUTOFRANGE, "DMDA: grid too large for 32-bit indices");
  DM d = new _p_DM;
  d->M = M; d->N = N; d->P = P;
  d->xs = d->ys = d->zs = 0;
  d->xm = M; d->ym = N; d->zm = P;
  d->dof = dof; d->sw = s; d->bx = bx; d->by = by; d->bz = bz; d->st = st;
  auto ghost = [&](DMBoundaryType b, PetscInt len, PetscInt &gs, PetscInt &gm) {
    if (b == DM_BOUNDARY_NONE) { gs = 0; gm = len; }
    else { gs = -s; gm = len + 2 * s; }
  };
  ghost(bx, M, d->gxs, d->gxm); ghost(by, N, d->gys, d->gym); ghost(bz, P, d->gzs, d->gzm);
  // local-to-global: natural numbering on one rank; ghosts outside the domain map to -1
  // (DM_BOUNDARY_GHOSTED), periodic images wrap
  d->ltog.map.assign((size_t)d->gxm * d->gym * d->gzm, -1);
  auto wrap = [](DMBoundaryType b, PetscInt c, PetscInt len) -> PetscInt {
    if (c >= 0 && c < len) return c;
    if (b == DM_BOUNDARY_PERIODIC) return (c % len + len) % len;
    return -1;
  };
  for (PetscInt k = 0; k < d->gzm; ++k)
    for (PetscInt j = 0; j < d->gym; ++j)
      for (PetscInt i = 0; i < d->gxm; ++i) {
        PetscInt gi = wrap(bx, i + d->gxs, M), gj = wrap(by, j + d->gys, N), gk = wrap(bz, k + d->gzs, P);
        if (gi < 0 || gj < 0 || gk < 0) continue;
        d->ltog.map[(size_t)i + (size_t)j * d->gxm + (size_t)k * d->gxm * d->gym] = gi + gj * M + gk * M * N;
      }
  d->mattype = MATAIJ;
  *da = d;
  return 0;
}
extern "C" PetscErrorCode DMSetMatType(DM da, MatType t)
{
  if (strcmp(t, MATAIJ) && strcmp(t, MATSEQAIJ)) SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_SUP, "DM: matrix type %s not supported", t);
  da->mattype = t;
  return 0;
}
extern "C" PetscErrorCode DMCreateGlobalVector(DM da, Vec *v) { return VecCreateSeq(PETSC_COMM_SELF, da->M * da->N * da->P, v); }
// DMCreateMatrix_DA_3d [P376]: exact preallocation for the clipped star (or box) stencil.  PETSc
// also inserts explicit zeros for the pattern; the application overwrites every one of them
// (src/helper.cpp:236), so only the preallocation is done here.
extern "C" PetscErrorCode DMCreateMatrix(DM da, Mat *A)
{
  const PetscInt M = da->M, N = da->N, P = da->P, n = M * N * P;
  std::vector<PetscInt> nnz((size_t)n);
  for (PetscInt k = 0; k < P; ++k)
    for (PetscInt j = 0; j < N; ++j)
      for (PetscInt i = 0; i < M; ++i) {
        PetscInt c;
        auto ext = [](DMBoundaryType b, PetscInt x, PetscInt len) { return (b == DM_BOUNDARY_PERIODIC) ? 3 : 1 + (x > 0) + (x < len - 1); };
        if (da->st == DMDA_STENCIL_STAR) c = 1 + (ext(da->bx, i, M) - 1) + (ext(da->by, j, N) - 1) + (ext(da->bz, k, P) - 1);
        else c = ext(da->bx, i, M) * ext(da->by, j, N) * ext(da->bz, k, P);
        nnz[(size_t)i + (size_t)j * M + (size_t)k * M * N] = c;
      }
  return MatCreateSeqAIJ(PETSC_COMM_SELF, n, n, 0, nnz.data(), A);
}
extern "C" PetscErrorCode DMDAGetLocalInfo(DM da, DMDALocalInfo *info)
{
  info->dim = 3; info->dof = da->dof; info->sw = da->sw;
  info->mx = da->M; info->my = da->N; info->mz = da->P;
  info->xs = da->xs; info->ys = da->ys; info->zs = da->zs;
  info->xm = da->xm; info->ym = da->ym; info->zm = da->zm;
  info->gxs = da->gxs; info->gys = da->gys; info->gzs = da->gzs;
  info->gxm = da->gxm; info->gym = da->gym; info->gzm = da->gzm;
  info->bx = da->bx; info->by = da->by; info->bz = da->bz; info->st = da->st; info->da = da;
  return 0;
}
extern "C" PetscErrorCode DMGetLocalToGlobalMapping(DM da, ISLocalToGlobalMapping *l) { *l = &da->ltog; return 0; }
// DMDAConvertToCell [P376]: (i,j,k) of a cell in the ghosted local box -> its local number
extern "C" PetscErrorCode DMDAConvertToCell(DM da, MatStencil s, PetscInt *cell)
{
  const PetscInt i = s.i - da->gxs, j = s.j - da->gys, k = s.k - da->gzs;
  *cell = -1;
  if (i < 0 || i >= da->gxm) SETERRQ3(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "Stencil i %d should be in [%d, %d)", s.i, da->gxs, da->gxs + da->gxm);
  if (j < 0 || j >= da->gym) SETERRQ3(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "Stencil j %d should be in [%d, %d)", s.j, da->gys, da->gys + da->gym);
  if (k < 0 || k >= da->gzm) SETERRQ3(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "Stencil k %d should be in [%d, %d)", s.k, da->gzs, da->gzs + da->gzm);
  *cell = i + j * da->gxm + k * da->gxm * da->gym;
  return 0;
}
extern "C" PetscErrorCode ISLocalToGlobalMappingApply(ISLocalToGlobalMapping l, PetscInt N, const PetscInt in[], PetscInt out[])
{
  const PetscInt n = (PetscInt)l->map.size();
  for (PetscInt t = 0; t < N; ++t) {
    if (in[t] < 0) { out[t] = in[t]; continue; }
    if (in[t] >= n) SETERRQ2(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "Local index %d too large %d (max)", in[t], n - 1);
    out[t] = l->map[in[t]];
  }
  return 0;
}
// DMDAVecGetArray: a[k][j][i] over the owned box of a global vector
extern "C" PetscErrorCode DMDAVecGetArray(DM da, Vec v, void *array)
{
  PetscScalar *base;
  PetscErrorCode ierr = VecGetArray(v, &base);CHKERRQ(ierr);
  const PetscInt xm = da->xm, ym = da->ym, zm = da->zm;
  PetscScalar ***k3 = (PetscScalar ***)malloc(sizeof(PetscScalar **) * (size_t)zm);
  PetscScalar  **j2 = (PetscScalar **)malloc(sizeof(PetscScalar *) * (size_t)zm * ym);
  for (PetscInt k = 0; k < zm; ++k) {
    for (PetscInt j = 0; j < ym; ++j) j2[(size_t)k * ym + j] = base + ((size_t)k * ym + j) * xm - da->xs;
    k3[k] = j2 + (size_t)k * ym - da->ys;
  }
  da->views[v] = {(void *)k3, (void *)j2};
  *(PetscScalar ****)array = k3 - da->zs;
  return 0;
}
extern "C" PetscErrorCode DMDAVecRestoreArray(DM da, Vec v, void *array)
{
  auto it = da->views.find(v);
  if (it != da->views.end()) { for (void *p : it->second) free(p); da->views.erase(it); }
  *(PetscScalar ****)array = NULL;
  PetscScalar *dummy = NULL;
  return VecRestoreArray(v, &dummy);
}
extern "C" PetscErrorCode DMDestroy(DM *da)
{
  if (da && *da) { delete *da; *da = NULL; }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// KSP
// ---------------------------------------------------------------------------------------------
struct _p_KSP {
  Mat A = NULL, P = NULL;
  PetscReal rtol = 1e-5, abstol = 1e-50, dtol = 1e5;
  PetscInt  max_it = 10000, its = 0;
  PetscReal rnorm = 0.0;
  KSPConvergedReason reason = KSP_CONVERGED_ITERATING;
  std::string pc = "jacobi";  // PETSc's default for one rank is ilu; the reference always sets -pc_type
  bool fused = false, setup = false;
  bool norm_none = false;     // KSP_NORM_NONE: no convergence test, exactly max_it iterations
  bool monitor = false, print_reason = false;  // -ksp_monitor, -ksp_converged_reason
  Vec dinv = NULL;
  B200PCGamg *mg = NULL;
  std::vector<PetscReal> *lanczos_d = NULL, *lanczos_e = NULL;  // KSPSetComputeSingularValues
};

extern "C" PetscErrorCode KSPCreate(MPI_Comm, KSP *k) { *k = new _p_KSP; return 0; }
extern "C" PetscErrorCode KSPSetOperators(KSP k, Mat A, Mat P) { k->A = A; k->P = P; k->setup = false; return 0; }
extern "C" PetscErrorCode KSPSetType(KSP, KSPType t)
{
  if (strcmp(t, KSPCG)) SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_SUP, "KSP type %s not supported (only cg)", t);
  return 0;
}
extern "C" PetscErrorCode KSPSetReusePreconditioner(KSP, PetscBool) { return 0; }
extern "C" PetscErrorCode KSPSetTolerances(KSP k, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt maxits)
{
  if (rtol != PETSC_DEFAULT) k->rtol = rtol;
  if (abstol != PETSC_DEFAULT) k->abstol = abstol;
  if (dtol != PETSC_DEFAULT) k->dtol = dtol;
  if (maxits != PETSC_DEFAULT) k->max_it = maxits;
  return 0;
}
extern "C" PetscErrorCode KSPSetFromOptions(KSP k)
{
  PetscErrorCode ierr;
  char      buf[256];
  PetscBool set;
  ierr = PetscOptionsGetString(NULL, NULL, "-ksp_type", buf, sizeof buf, &set);CHKERRQ(ierr);
  if (set) { ierr = KSPSetType(k, buf);CHKERRQ(ierr); }
  ierr = PetscOptionsGetReal(NULL, NULL, "-ksp_rtol", &k->rtol, NULL);CHKERRQ(ierr);
  ierr = PetscOptionsGetReal(NULL, NULL, "-ksp_atol", &k->abstol, NULL);CHKERRQ(ierr);
  ierr = PetscOptionsGetInt(NULL, NULL, "-ksp_max_it", &k->max_it, NULL);CHKERRQ(ierr);
  ierr = PetscOptionsGetString(NULL, NULL, "-pc_type", buf, sizeof buf, &set);CHKERRQ(ierr);
  if (set) {
    if (!strcmp(buf, "jacobi") || !strcmp(buf, "none") || !strcmp(buf, "gamg")) k->pc = buf;
    else SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_SUP, "PC type %s is not supported (jacobi, none, gamg)", buf);
  }
  ierr = PetscOptionsGetString(NULL, NULL, "-ksp_b200_fused", NULL, 0, &set);CHKERRQ(ierr);
  k->fused = set;
  ierr = PetscOptionsGetString(NULL, NULL, "-ksp_monitor", NULL, 0, &set);CHKERRQ(ierr);
  k->monitor = set;
  ierr = PetscOptionsGetString(NULL, NULL, "-ksp_converged_reason", NULL, 0, &set);CHKERRQ(ierr);
  k->print_reason = set;
  k->setup = false;
  return 0;
}
extern "C" PetscErrorCode KSPSetUp(KSP k)
{
  PetscErrorCode ierr;
  if (!k->A) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_WRONGSTATE, "KSPSetOperators first");
  if (k->setup) return 0;
  if (k->mg) { ierr = b200_pcgamg_destroy(&k->mg);CHKERRQ(ierr); }
  if (k->pc == "jacobi") {
    // PCSetUp_Jacobi [P376]: diagonal, reciprocal, zero entries -> 1
    PetscInt rows, have = -1;
    ierr = MatGetLocalSize(k->P ? k->P : k->A, &rows, NULL);CHKERRQ(ierr);
    if (k->dinv) { ierr = VecGetLocalSize(k->dinv, &have);CHKERRQ(ierr); }
    if (k->dinv && have != rows) { ierr = VecDestroy(&k->dinv);CHKERRQ(ierr); }   // operators of another size were set
    if (!k->dinv) { ierr = MatCreateVecs(k->P ? k->P : k->A, NULL, &k->dinv);CHKERRQ(ierr); }
    ierr = MatGetDiagonal(k->P ? k->P : k->A, k->dinv);CHKERRQ(ierr);
    PetscScalar *d;
    PetscInt     n;
    ierr = VecGetLocalSize(k->dinv, &n);CHKERRQ(ierr);
    ierr = VecGetArray(k->dinv, &d);CHKERRQ(ierr);
    for (PetscInt i = 0; i < n; ++i) d[i] = (d[i] != 0.0) ? 1.0 / d[i] : 1.0;
    ierr = VecRestoreArray(k->dinv, &d);CHKERRQ(ierr);
  } else if (k->pc == "gamg") {
    ierr = b200_pcgamg_setup(k->P ? k->P : k->A, &k->mg);CHKERRQ(ierr);
  }
  k->setup = true;
  return 0;
}

static PetscErrorCode pc_apply(KSP k, Vec r, Vec z)
{
  if (k->pc == "jacobi") return VecPointwiseMult(z, r, k->dinv);
  if (k->pc == "gamg") return b200_pcgamg_apply(k->mg, r, z);
  return VecCopy(r, z);
}

static bool converged(KSP k, PetscReal rnorm, PetscReal rnorm0, PetscInt it)
{
  // KSPConvergedDefault [P376]
  if (k->norm_none) return false;
  const PetscReal ttol = PetscMax(k->rtol * rnorm0, k->abstol);
  if (rnorm != rnorm) { k->reason = KSP_DIVERGED_NANORINF; return true; }
  if (rnorm < ttol) { k->reason = (rnorm < k->abstol) ? KSP_CONVERGED_ATOL : KSP_CONVERGED_RTOL; return true; }
  if (it > 0 && rnorm >= k->dtol * rnorm0) { k->reason = KSP_DIVERGED_DTOL; return true; }
  return false;
}

// KSPSolve_CG [P376], zero initial guess (src/main_ksp.cpp:103 passes lhs = 0)
extern "C" PetscErrorCode KSPSolve(KSP k, Vec b, Vec x)
{
  PetscErrorCode ierr;
  ierr = KSPSetUp(k);CHKERRQ(ierr);
  k->reason = KSP_CONVERGED_ITERATING;
  k->its = 0;
  if (k->fused && k->pc == "jacobi" && !k->norm_none && !k->monitor) {
    // one-library-call variant: everything (scalars included) stays on the device
    Mat_SeqAIJ *a = (Mat_SeqAIJ *)k->A->data;
    int rc = b200_petsc_ensure_resident(&k->A->spptr, k->A->rmap->n, k->A->cmap->n, a->i, a->j, a->a, (int64_t)k->A->state);
    if (rc) return rc;
    b200_csr_t h = (b200_csr_t)b200_petsc_handle(&k->A->spptr);
    const PetscScalar *db;
    PetscScalar *dx;
    ierr = VecB200GetDeviceArrayRead(b, &db);CHKERRQ(ierr);
    ierr = VecB200GetDeviceArrayWrite(x, &dx);CHKERRQ(ierr);
    b200_cg_result_t out;
    rc = b200_cg_jacobi(h, db, dx, k->rtol, k->abstol, k->max_it, b200_petsc_mode(), &out, NULL);
    if (rc) return PetscError(PETSC_COMM_SELF, __LINE__, __func__, __FILE__, rc, "b200_cg_jacobi");
    k->its = out.its; k->rnorm = out.rnorm; k->reason = (KSPConvergedReason)out.reason;
    return 0;
  }
  Vec r, z, p, w;
  ierr = VecDuplicate(b, &r);CHKERRQ(ierr);
  ierr = VecDuplicate(b, &z);CHKERRQ(ierr);
  ierr = VecDuplicate(b, &p);CHKERRQ(ierr);
  ierr = VecDuplicate(b, &w);CHKERRQ(ierr);
  ierr = VecSet(x, 0.0);CHKERRQ(ierr);
  ierr = VecCopy(b, r);CHKERRQ(ierr);
  ierr = pc_apply(k, r, z);CHKERRQ(ierr);
  PetscReal dp = 0.0, rnorm0;
  PetscScalar beta = 0.0, betaold = 1.0, dpi, a = 1.0;
  if (!k->norm_none) { ierr = VecNorm(z, NORM_2, &dp);CHKERRQ(ierr); }
  rnorm0 = dp;
  k->rnorm = dp;
  PetscInt it = 0;
  // KSPMonitorDefault's line format [P376]
  if (k->monitor) { ierr = PetscPrintf(PETSC_COMM_WORLD, "%3d KSP Residual norm %14.12e \n", 0, dp);CHKERRQ(ierr); }
  if (!converged(k, dp, rnorm0, 0)) {
    ierr = VecDot(z, r, &beta);CHKERRQ(ierr);
    PetscScalar dpiold = 0.0;
    while (it < k->max_it) {
      PetscScalar bq = 0.0, eoff = 0.0;
      // KSPSolve_CG's exits for a vanished or sign-changing (z,r) [P376]
      if (beta == 0.0) { k->reason = KSP_CONVERGED_ATOL; break; }
      if (it > 0 && beta * betaold < 0.0) { k->reason = KSP_DIVERGED_INDEFINITE_PC; break; }
      if (it == 0) { ierr = VecCopy(z, p);CHKERRQ(ierr); }
      else {
        bq = beta / betaold;
        eoff = std::sqrt(std::fabs(bq)) / a;   // KSPCG's e[i] with the previous step length
        ierr = VecAYPX(p, bq, z);CHKERRQ(ierr);
      }
      betaold = beta;
      ierr = MatMult(k->A, p, w);CHKERRQ(ierr);          // -> A->ops->mult = MatMult_SeqAIJ
      ierr = VecDot(p, w, &dpi);CHKERRQ(ierr);
      if (!k->norm_none && (dpi == 0.0 || (it > 0 && dpi * dpiold <= 0.0))) { k->reason = KSP_DIVERGED_INDEFINITE_MAT; break; }
      dpiold = dpi;
      a = beta / dpi;
      if (k->lanczos_d) {                                // KSPCG's d[i], e[i] [P376]
        k->lanczos_e->push_back(eoff);
        k->lanczos_d->push_back(std::sqrt(std::fabs(bq)) * eoff + 1.0 / a);
      }
      ierr = VecAXPY(x, a, p);CHKERRQ(ierr);
      ierr = VecAXPY(r, -a, w);CHKERRQ(ierr);
      ierr = pc_apply(k, r, z);CHKERRQ(ierr);
      if (!k->norm_none) { ierr = VecNorm(z, NORM_2, &dp);CHKERRQ(ierr); }
      ++it;
      k->rnorm = dp;
      if (k->monitor) { ierr = PetscPrintf(PETSC_COMM_WORLD, "%3d KSP Residual norm %14.12e \n", it, dp);CHKERRQ(ierr); }
      if (converged(k, dp, rnorm0, it)) break;
      ierr = VecDot(z, r, &beta);CHKERRQ(ierr);
    }
    if (k->reason == KSP_CONVERGED_ITERATING) k->reason = k->norm_none ? KSP_CONVERGED_ITS : KSP_DIVERGED_ITS;
  }
  k->its = it;
  if (k->print_reason) {
    ierr = PetscPrintf(PETSC_COMM_WORLD, "Linear solve %s due to reason %d iterations %d\n", k->reason > 0 ? "converged" : "did not converge",
                       (int)k->reason, it);CHKERRQ(ierr);
  }
  ierr = VecDestroy(&r);CHKERRQ(ierr);
  ierr = VecDestroy(&z);CHKERRQ(ierr);
  ierr = VecDestroy(&p);CHKERRQ(ierr);
  ierr = VecDestroy(&w);CHKERRQ(ierr);
  return 0;
}
extern "C" PetscErrorCode KSPGetConvergedReason(KSP k, KSPConvergedReason *r) { *r = k->reason; return 0; }
extern "C" PetscErrorCode KSPGetIterationNumber(KSP k, PetscInt *n) { *n = k->its; return 0; }
extern "C" PetscErrorCode KSPGetResidualNorm(KSP k, PetscReal *r) { *r = k->rnorm; return 0; }
extern "C" PetscErrorCode KSPDestroy(KSP *k)
{
  if (k && *k) {
    if ((*k)->dinv) VecDestroy(&(*k)->dinv);
    if ((*k)->mg) b200_pcgamg_destroy(&(*k)->mg);
    delete *k;
    *k = NULL;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// largest eigenvalue of the symmetric tridiagonal (d, e[1..]) by Sturm-sequence bisection
// (PETSc hands the same matrix to LAPACK's sterf in KSPComputeExtremeSingularValues_CG)
// ---------------------------------------------------------------------------------------------
extern "C" PetscErrorCode b200_tridiag_emax(PetscInt n, const PetscReal d[], const PetscReal e[], PetscReal *emax)
{
  if (n < 1) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "empty tridiagonal");
  PetscReal lo = d[0], hi = d[0];
  for (PetscInt i = 0; i < n; ++i) {
    const PetscReal rad = (i > 0 ? std::fabs(e[i]) : 0.0) + (i + 1 < n ? std::fabs(e[i + 1]) : 0.0);
    lo = PetscMin(lo, d[i] - rad);
    hi = PetscMax(hi, d[i] + rad);
  }
  // count(x) = number of eigenvalues below x = negative pivots of T - x I
  auto below = [&](PetscReal x) {
    PetscInt  c = 0;
    PetscReal q = 1.0;
    for (PetscInt i = 0; i < n; ++i) {
      const PetscReal off = (i > 0) ? e[i] * e[i] : 0.0;
      q = d[i] - x - (i > 0 ? off / q : 0.0);
      if (q == 0.0) q = 1e-300;
      if (q < 0.0) ++c;
    }
    return c;
  };
  for (int it = 0; it < 200 && hi - lo > 4e-16 * PetscMax(std::fabs(lo), std::fabs(hi)); ++it) {
    const PetscReal mid = 0.5 * (lo + hi);
    if (below(mid) >= n) hi = mid; else lo = mid;
  }
  *emax = 0.5 * (lo + hi);
  return 0;
}

PetscErrorCode b200_ksp_estimate_emax(Mat A, PetscInt its, PetscReal *emax)
{
  PetscErrorCode         ierr;
  KSP                    e;
  Vec                    bb, xx;
  PetscScalar           *arr;
  PetscInt               n;
  std::vector<PetscReal> d, off;
  ierr = MatCreateVecs(A, &xx, &bb);CHKERRQ(ierr);
  ierr = VecGetLocalSize(bb, &n);CHKERRQ(ierr);
  ierr = VecGetArray(bb, &arr);CHKERRQ(ierr);
  if (b200_gen_vector(arr, n, 0x6A36ULL)) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_LIB, "b200_gen_vector");
  ierr = VecRestoreArray(bb, &arr);CHKERRQ(ierr);
  ierr = KSPCreate(PETSC_COMM_SELF, &e);CHKERRQ(ierr);
  ierr = KSPSetOperators(e, A, A);CHKERRQ(ierr);
  e->pc = "jacobi";
  e->norm_none = true;
  e->max_it = PetscMax(its, 1);
  e->lanczos_d = &d;
  e->lanczos_e = &off;
  ierr = KSPSolve(e, bb, xx);CHKERRQ(ierr);
  ierr = KSPDestroy(&e);CHKERRQ(ierr);
  ierr = VecDestroy(&bb);CHKERRQ(ierr);
  ierr = VecDestroy(&xx);CHKERRQ(ierr);
  // a breakdown (dpi = 0 on a tiny operator) leaves non-finite entries: cut the recurrence there
  PetscInt m = 0;
  while (m < (PetscInt)d.size() && std::isfinite(d[m]) && std::isfinite(off[m])) ++m;
  if (m < 1) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_CONV_FAILED, "eigenvalue estimate broke down in the first iteration");
  return b200_tridiag_emax(m, d.data(), off.data(), emax);
}

// ---------------------------------------------------------------------------------------------
// PC access for the tests: the hierarchy and the bare preconditioner application
// ---------------------------------------------------------------------------------------------
extern "C" PetscErrorCode PCGAMGGetNumLevelsB200(KSP k, PetscInt *n)
{
  if (!k->mg) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_WRONGSTATE, "KSPSetUp with -pc_type gamg first");
  *n = b200_pcgamg_num_levels(k->mg);
  return 0;
}
extern "C" PetscErrorCode PCGAMGGetLevelB200(KSP k, PetscInt level, Mat *A, Mat *P, Vec *dinv, const PetscInt **agg, PetscInt *nagg,
                                             PetscReal *emax)
{
  if (!k->mg) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_WRONGSTATE, "KSPSetUp with -pc_type gamg first");
  return b200_pcgamg_level(k->mg, level, A, P, dinv, agg, nagg, emax);
}
extern "C" PetscErrorCode KSPApplyPCB200(KSP k, Vec r, Vec z)
{
  PetscErrorCode ierr = KSPSetUp(k);CHKERRQ(ierr);
  return pc_apply(k, r, z);
}
