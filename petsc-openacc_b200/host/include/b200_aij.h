/*
 * b200_aij.h -- private object layouts, the counterpart of PETSc 3.7.6's
 * src/mat/impls/aij/seq/aij.h + petsc/private/matimpl.h + vecimpl.h for the fields the hot path
 * touches.  Field names follow PETSc so that src/seqaij_symbols.cpp reads the same against either.
 * Visible in the reference's patches: a->i, a->j, a->a, a->nz (src/openacc-step2/
 * MatDestroy_SeqAIJ.patch:22-29), a->nonzerorowcnt (src/openacc-step2/MatMult_SeqAIJ.patch:47),
 * a->ilen / a->imax / a->compressedrow (src/openacc-step2/MatAssemblyEnd_SeqAIJ.patch:31-35,77),
 * A->rmap->n (:21).
 */
#ifndef B200_AIJ_H
#define B200_AIJ_H
#include "b200_petsc.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct _n_PetscLayout { PetscInt n, N, rstart, rend; } *PetscLayout;

struct _MatOps {
  PetscErrorCode (*mult)(Mat, Vec, Vec);
  PetscErrorCode (*multadd)(Mat, Vec, Vec, Vec);
  PetscErrorCode (*multtranspose)(Mat, Vec, Vec);
  PetscErrorCode (*multtransposeadd)(Mat, Vec, Vec, Vec);
  PetscErrorCode (*assemblyend)(Mat, MatAssemblyType);
  PetscErrorCode (*destroy)(Mat);
  PetscErrorCode (*getdiagonal)(Mat, Vec);
  PetscErrorCode (*setvalues)(Mat, PetscInt, const PetscInt[], PetscInt, const PetscInt[], const PetscScalar[], InsertMode);
  PetscErrorCode (*zerorowscolumns)(Mat, PetscInt, const PetscInt[], PetscScalar, Vec, Vec);
  PetscErrorCode (*scale)(Mat, PetscScalar);
};

struct _p_Mat {
  struct _MatOps ops[1];
  MatType        type_name;
  PetscLayout    rmap, cmap;
  void          *data;     /* Mat_SeqAIJ*                                                     */
  void          *spptr;    /* device residency record (PETSc: "for external packages")        */
  PetscBool      assembled;
  PetscInt       state;    /* PetscObjectState: bumped on every value/pattern change          */
  struct { PetscReal nz_unneeded, mallocs; } info;
};

typedef struct {
  PetscBool use;
  PetscInt  nrows;
  PetscInt *i;
  PetscInt *rindex;
} Mat_CompressedRow;

typedef struct {
  PetscInt         *i, *j;       /* row pointers, column indices                              */
  MatScalar        *a;           /* values                                                    */
  PetscInt          nz, maxnz;
  PetscInt         *imax, *ilen; /* reserved / used slots per row                             */
  PetscInt          nonzerorowcnt, rmax;
  PetscInt         *diag;        /* position of the diagonal entry of each row (or -1)        */
  Mat_CompressedRow compressedrow;
  PetscInt          reallocs, nounused;
  PetscBool         singlemalloc, roworiented;
} Mat_SeqAIJ;

struct _p_Vec {
  PetscInt     n;
  PetscScalar *array;    /* page-locked host memory                                           */
  PetscScalar *d_array;  /* device mirror (allocated on first device access)                  */
  PetscBool    host_valid, dev_valid;
  PetscInt     state;
};

/* private helpers that PETSc also has under these names */
PetscErrorCode MatMarkDiagonal_SeqAIJ(Mat A);
PetscErrorCode MatCheckCompressedRow(Mat A, PetscInt nrows, Mat_CompressedRow *c, PetscInt *ai, PetscInt mbs, PetscReal ratio);
PetscErrorCode MatAssemblyEnd_SeqAIJ_Inode(Mat A, MatAssemblyType mode);
PetscErrorCode MatSeqAIJInvalidateDiagonal(Mat A);
PetscErrorCode MatSeqXAIJFreeAIJ(Mat A, MatScalar **a, PetscInt **j, PetscInt **i);

#ifdef __cplusplus
}
#endif
#endif
