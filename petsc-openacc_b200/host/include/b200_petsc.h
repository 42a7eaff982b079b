/*
 * b200_petsc.h -- the slice of PETSc 3.7.6's public C API that the reference application
 * (src/main_ksp.cpp, src/helper.cpp) and the SeqAIJ hot path use, provided by libb200petsc.so.
 *
 * Why it exists: PETSc 3.7.6 is downloaded by the reference's build (scripts/petsc.sh:38-40) and
 * cannot be had offline, so the host side above the C ABI mirrors the reference's operator
 * interface here -- same names, same argument meaning, same error convention (0 / non-zero
 * PetscErrorCode through CHKERRQ).  When real PETSc headers are available the five replaced
 * symbols in src/seqaij_symbols.cpp compile against them instead (-DB200_WITH_PETSC, see
 * INTEGRATION.md); nothing else in this directory is needed then.
 *
 * Deliberate differences from PETSc: single process (MPI_* are stubs, PETSC_COMM_WORLD has one
 * rank; the multi-GPU path is include/b200_mpiaij.h driven by torch.distributed); Vec keeps a
 * device mirror with lazy host<->device sync; KSPCG only, with PCJACOBI, PCNONE or a PCGAMG that
 * covers the combination the reference's options file selects (src/pcgamg.cpp).
 */
#ifndef B200_PETSC_H
#define B200_PETSC_H

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#ifdef __cplusplus
#include <cmath>
extern "C" {
#endif

/* ---- basic types (scripts/petsc-release.sh:6,62; no 64-bit indices) ---------------------- */
typedef int    PetscErrorCode;
typedef int    PetscInt;
typedef int    PetscMPIInt;
typedef double PetscScalar;
typedef double PetscReal;
typedef double MatScalar;
typedef double PetscLogDouble;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
typedef int    MPI_Comm;
typedef struct _p_PetscOptions *PetscOptions;

#define PETSC_COMM_WORLD 1
#define PETSC_COMM_SELF 2
#define MPI_COMM_WORLD 1
#define PETSC_DECIDE (-1)
#define PETSC_DEFAULT (-2)
#define PETSC_MAX_PATH_LEN 4096
#ifndef PETSC_NULL
#define PETSC_NULL NULL
#endif

/* error codes used on the path (petscerror.h values) */
#define PETSC_ERR_MEM 55
#define PETSC_ERR_SUP 56
#define PETSC_ERR_ARG_SIZ 60
#define PETSC_ERR_ARG_WRONG 62
#define PETSC_ERR_ARG_OUTOFRANGE 63
#define PETSC_ERR_ARG_WRONGSTATE 73
#define PETSC_ERR_LIB 76
#define PETSC_ERR_CONV_FAILED 82
#define PETSC_ERR_FILE_OPEN 65
#define PETSC_ERR_USER 83

PetscErrorCode PetscError(MPI_Comm, int line, const char *func, const char *file, PetscErrorCode code, const char *fmt, ...);
#define CHKERRQ(ierr) do { if ((ierr)) return PetscError(PETSC_COMM_SELF, __LINE__, __func__, __FILE__, (ierr), " "); } while (0)
#define SETERRQ(comm, code, msg) return PetscError(comm, __LINE__, __func__, __FILE__, code, msg)
#define SETERRQ1(comm, code, msg, a) return PetscError(comm, __LINE__, __func__, __FILE__, code, msg, a)
#define SETERRQ2(comm, code, msg, a, b) return PetscError(comm, __LINE__, __func__, __FILE__, code, msg, a, b)
#define SETERRQ3(comm, code, msg, a, b, c) return PetscError(comm, __LINE__, __func__, __FILE__, code, msg, a, b, c)
#define PetscFunctionBegin do {} while (0)
#define PetscFunctionBeginUser do {} while (0)
#define PetscFunctionReturn(a) return (a)
#define PetscMax(a, b) (((a) < (b)) ? (b) : (a))
#define PetscMin(a, b) (((a) < (b)) ? (a) : (b))

/* ---- MPI stubs (one rank) ------------------------------------------------------------------ */
int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm, int *rank);
int MPI_Comm_size(MPI_Comm, int *size);

/* ---- sys ----------------------------------------------------------------------------------- */
PetscErrorCode PetscInitialize(int *argc, char ***args, const char file[], const char help[]);
PetscErrorCode PetscFinalize(void);
PetscErrorCode PetscOptionsGetString(PetscOptions, const char pre[], const char name[], char str[], size_t len, PetscBool *set);
PetscErrorCode PetscOptionsGetInt(PetscOptions, const char pre[], const char name[], PetscInt *v, PetscBool *set);
PetscErrorCode PetscOptionsGetReal(PetscOptions, const char pre[], const char name[], PetscReal *v, PetscBool *set);
PetscErrorCode PetscOptionsInsertFile(MPI_Comm, PetscOptions, const char file[], PetscBool require);
PetscErrorCode PetscOptionsSetValue(PetscOptions, const char name[], const char value[]);
PetscErrorCode PetscOptionsClear(PetscOptions);
PetscErrorCode PetscTime(PetscLogDouble *t);
PetscErrorCode PetscPrintf(MPI_Comm, const char fmt[], ...);
PetscErrorCode PetscLogFlops(PetscLogDouble f);
PetscErrorCode PetscGetFlops(PetscLogDouble *f);

/* ---- Vec ----------------------------------------------------------------------------------- */
typedef struct _p_Vec *Vec;
typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_FROBENIUS = 2, NORM_INFINITY = 3, NORM_1_AND_2 = 4 } NormType;
typedef enum { NOT_SET_VALUES, INSERT_VALUES, ADD_VALUES, MAX_VALUES, INSERT_ALL_VALUES, ADD_ALL_VALUES } InsertMode;

PetscErrorCode VecCreateSeq(MPI_Comm, PetscInt n, Vec *v);
PetscErrorCode VecDuplicate(Vec v, Vec *newv);
PetscErrorCode VecDestroy(Vec *v);
PetscErrorCode VecSet(Vec x, PetscScalar alpha);
PetscErrorCode VecCopy(Vec x, Vec y);
PetscErrorCode VecGetSize(Vec x, PetscInt *size);
PetscErrorCode VecGetLocalSize(Vec x, PetscInt *size);
PetscErrorCode VecGetArray(Vec x, PetscScalar **a);
PetscErrorCode VecRestoreArray(Vec x, PetscScalar **a);
PetscErrorCode VecGetArrayRead(Vec x, const PetscScalar **a);
PetscErrorCode VecRestoreArrayRead(Vec x, const PetscScalar **a);
PetscErrorCode VecSum(Vec x, PetscScalar *sum);
PetscErrorCode VecDot(Vec x, Vec y, PetscScalar *val);
PetscErrorCode VecNorm(Vec x, NormType type, PetscReal *val);
PetscErrorCode VecAXPY(Vec y, PetscScalar alpha, Vec x);   /* y = y + alpha x */
PetscErrorCode VecAYPX(Vec y, PetscScalar alpha, Vec x);   /* y = x + alpha y */
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y);
PetscErrorCode VecReciprocal(Vec v);
/* device mirror (extension): device pointer valid until the next host write access */
PetscErrorCode VecB200GetDeviceArrayRead(Vec x, const PetscScalar **d);
PetscErrorCode VecB200GetDeviceArrayWrite(Vec x, PetscScalar **d);
PetscErrorCode VecB200GetDeviceArray(Vec x, PetscScalar **d); /* read-write */
PetscErrorCode VecB200HasDevice(Vec x, PetscBool *flg);        /* device copy is current */

/* ---- Mat ----------------------------------------------------------------------------------- */
typedef struct _p_Mat *Mat;
typedef const char    *MatType;
#define MATAIJ "aij"
#define MATSEQAIJ "seqaij"
#define MATMPIAIJ "mpiaij"
typedef enum { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 } MatAssemblyType;
typedef struct { PetscInt k, j, i, c; } MatStencil;

PetscErrorCode MatCreateSeqAIJ(MPI_Comm, PetscInt m, PetscInt n, PetscInt nz, const PetscInt nnz[], Mat *A);
PetscErrorCode MatSetValues(Mat A, PetscInt m, const PetscInt idxm[], PetscInt n, const PetscInt idxn[], const PetscScalar v[], InsertMode addv);
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType type);
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType type);
PetscErrorCode MatMult(Mat A, Vec x, Vec y);
PetscErrorCode MatMultAdd(Mat A, Vec v1, Vec v2, Vec v3);            /* v3 = v2 + A v1 */
PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y);
PetscErrorCode MatMultTransposeAdd(Mat A, Vec v1, Vec v2, Vec v3);   /* v3 = v2 + A' v1 */
PetscErrorCode MatGetDiagonal(Mat A, Vec v);
PetscErrorCode MatGetSize(Mat A, PetscInt *m, PetscInt *n);
PetscErrorCode MatGetLocalSize(Mat A, PetscInt *m, PetscInt *n);
PetscErrorCode MatCreateVecs(Mat A, Vec *right, Vec *left);
PetscErrorCode MatZeroRowsColumns(Mat A, PetscInt n, const PetscInt rows[], PetscScalar diag, Vec x, Vec b);
PetscErrorCode MatScale(Mat A, PetscScalar a);
PetscErrorCode MatDestroy(Mat *A);
/* assembled matrix from a finished CSR with ascending columns (arrays are copied) */
PetscErrorCode MatCreateSeqAIJFromCSRB200(PetscInt m, PetscInt n, const PetscInt i[], const PetscInt j[], const PetscScalar a[], Mat *A);
/* read access to the assembled CSR (PETSc: MatSeqAIJGetArray / MatGetRowIJ) for tests */
PetscErrorCode MatSeqAIJGetCSRB200(Mat A, PetscInt *m, PetscInt *n, PetscInt *nz, const PetscInt **i, const PetscInt **j, const PetscScalar **a);
PetscErrorCode MatSeqAIJGetInfoB200(Mat A, PetscInt *nonzerorowcnt, PetscInt *rmax, PetscBool *compressedrow, PetscInt *cprow_nrows, PetscInt *fshift);

/* ---- input formats: PETSc binary (MatLoad / VecLoad) and MatrixMarket ----------------------- */
typedef struct _p_PetscViewer *PetscViewer;
typedef enum { FILE_MODE_READ, FILE_MODE_WRITE, FILE_MODE_APPEND, FILE_MODE_UPDATE, FILE_MODE_APPEND_UPDATE } PetscFileMode;
#define PETSC_ERR_FILE_UNEXPECTED 79
PetscErrorCode PetscViewerBinaryOpen(MPI_Comm, const char name[], PetscFileMode mode, PetscViewer *viewer);
PetscErrorCode PetscViewerDestroy(PetscViewer *viewer);
/* PETSc 3.7.6 signature is MatLoad(Mat, PetscViewer) on a created Mat; here the Mat is created by the call */
PetscErrorCode MatLoad(Mat *newmat, PetscViewer viewer);
PetscErrorCode VecLoad(Vec *newvec, PetscViewer viewer);
PetscErrorCode MatLoadMatrixMarketB200(const char path[], Mat *newmat);

/* ---- DMDA (what src/helper.cpp uses) ------------------------------------------------------- */
typedef struct _p_DM *DM;
typedef struct _p_ISLocalToGlobalMapping *ISLocalToGlobalMapping;
typedef enum { DM_BOUNDARY_NONE, DM_BOUNDARY_GHOSTED, DM_BOUNDARY_MIRROR, DM_BOUNDARY_PERIODIC, DM_BOUNDARY_TWIST } DMBoundaryType;
typedef enum { DMDA_STENCIL_STAR, DMDA_STENCIL_BOX } DMDAStencilType;
typedef struct {
  PetscInt        dim, dof, sw;
  PetscInt        mx, my, mz;
  PetscInt        xs, ys, zs;
  PetscInt        xm, ym, zm;
  PetscInt        gxs, gys, gzs;
  PetscInt        gxm, gym, gzm;
  DMBoundaryType  bx, by, bz;
  DMDAStencilType st;
  DM              da;
} DMDALocalInfo;

PetscErrorCode DMDACreate3d(MPI_Comm, DMBoundaryType bx, DMBoundaryType by, DMBoundaryType bz, DMDAStencilType st,
                            PetscInt M, PetscInt N, PetscInt P, PetscInt m, PetscInt n, PetscInt p, PetscInt dof,
                            PetscInt s, const PetscInt lx[], const PetscInt ly[], const PetscInt lz[], DM *da);
PetscErrorCode DMSetMatType(DM, MatType);
PetscErrorCode DMCreateGlobalVector(DM, Vec *);
PetscErrorCode DMCreateMatrix(DM, Mat *);
PetscErrorCode DMDAGetLocalInfo(DM, DMDALocalInfo *);
PetscErrorCode DMGetLocalToGlobalMapping(DM, ISLocalToGlobalMapping *);
PetscErrorCode DMDAConvertToCell(DM, MatStencil s, PetscInt *cell);
PetscErrorCode ISLocalToGlobalMappingApply(ISLocalToGlobalMapping, PetscInt N, const PetscInt in[], PetscInt out[]);
PetscErrorCode DMDAVecGetArray(DM, Vec, void *array);
PetscErrorCode DMDAVecRestoreArray(DM, Vec, void *array);
PetscErrorCode DMDestroy(DM *);

/* ---- KSP (KSPCG + PCJACOBI / PCNONE / PCGAMG) ---------------------------------------------- */
typedef struct _p_KSP *KSP;
typedef const char    *KSPType;
#define KSPCG "cg"
typedef enum {
  KSP_CONVERGED_RTOL_NORMAL = 1, KSP_CONVERGED_ATOL_NORMAL = 9, KSP_CONVERGED_RTOL = 2, KSP_CONVERGED_ATOL = 3,
  KSP_CONVERGED_ITS = 4, KSP_DIVERGED_NULL = -2, KSP_DIVERGED_ITS = -3, KSP_DIVERGED_DTOL = -4,
  KSP_DIVERGED_BREAKDOWN = -5, KSP_DIVERGED_INDEFINITE_PC = -8, KSP_DIVERGED_NANORINF = -9, KSP_DIVERGED_INDEFINITE_MAT = -10,
  KSP_CONVERGED_ITERATING = 0
} KSPConvergedReason;

PetscErrorCode KSPCreate(MPI_Comm, KSP *);
PetscErrorCode KSPSetOperators(KSP, Mat Amat, Mat Pmat);
PetscErrorCode KSPSetType(KSP, KSPType);
PetscErrorCode KSPSetReusePreconditioner(KSP, PetscBool);
PetscErrorCode KSPSetFromOptions(KSP);
PetscErrorCode KSPSetTolerances(KSP, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt maxits);
PetscErrorCode KSPSetUp(KSP);
PetscErrorCode KSPSolve(KSP, Vec b, Vec x);
PetscErrorCode KSPGetConvergedReason(KSP, KSPConvergedReason *);
PetscErrorCode KSPGetIterationNumber(KSP, PetscInt *);
PetscErrorCode KSPGetResidualNorm(KSP, PetscReal *);
PetscErrorCode KSPDestroy(KSP *);
/* -pc_type gamg: read access to the hierarchy (level 0 = finest; P, agg are NULL on the coarsest)
 * and the bare preconditioner application z = M^-1 r, for tests and tools */
PetscErrorCode PCGAMGGetNumLevelsB200(KSP, PetscInt *nlevels);
PetscErrorCode PCGAMGGetLevelB200(KSP, PetscInt level, Mat *A, Mat *P, Vec *dinv, const PetscInt **agg, PetscInt *nagg, PetscReal *emax);
PetscErrorCode KSPApplyPCB200(KSP, Vec r, Vec z);
/* largest eigenvalue of the symmetric tridiagonal with diagonal d[0..n-1], off-diagonal e[1..n-1] */
PetscErrorCode b200_tridiag_emax(PetscInt n, const PetscReal d[], const PetscReal e[], PetscReal *emax);

#ifdef __cplusplus
}
#endif
#endif /* B200_PETSC_H */
