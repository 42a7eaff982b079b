/*
 * b200_petsc_symbols.h -- the PETSc-named entry points of the drop-in (libb200petsc.so), i.e.
 * exactly the symbols PETSc's MATSEQAIJ operator table binds and the reference substitutes at
 * link time (Makefile:153-158 lists the replacement objects ahead of libpetsc.a):
 *
 *   MatMult_SeqAIJ          src/openacc-step3/MatMult_SeqAIJ.patch:12
 *   MatAssemblyEnd_SeqAIJ   src/openacc-step2/MatAssemblyEnd_SeqAIJ.patch:12
 *   MatDestroy_SeqAIJ       src/openacc-step2/MatDestroy_SeqAIJ.patch:12
 *   MatMultAdd_SeqAIJ, MatMultTranspose_SeqAIJ, MatMultTransposeAdd_SeqAIJ
 *                           PETSc 3.7.6 aij.c (named by the north star; absent from the reference)
 *
 * and the residency hooks underneath them, which take only raw arrays so that a real-PETSc build
 * of the same functions can call them from C.
 */
#ifndef B200_PETSC_SYMBOLS_H
#define B200_PETSC_SYMBOLS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef B200_PETSC_H
typedef struct _p_Mat *Mat;
typedef struct _p_Vec *Vec;
typedef int            PetscErrorCode;
typedef enum { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 } MatAssemblyType;
#endif

PetscErrorCode MatMult_SeqAIJ(Mat A, Vec xx, Vec yy);
PetscErrorCode MatMultAdd_SeqAIJ(Mat A, Vec xx, Vec yy, Vec zz);
PetscErrorCode MatMultTranspose_SeqAIJ(Mat A, Vec xx, Vec yy);
PetscErrorCode MatMultTransposeAdd_SeqAIJ(Mat A, Vec xx, Vec zz, Vec yy);
PetscErrorCode MatAssemblyEnd_SeqAIJ(Mat A, MatAssemblyType mode);
PetscErrorCode MatDestroy_SeqAIJ(Mat A);

/* ---- residency hooks (replace acc_is_present / enter data / exit data,
 *      src/openacc-step2/MatAssemblyEnd_SeqAIJ.patch:21-29,42-44, MatDestroy...:26-34) -------- */
/* *slot is the per-matrix record kept in Mat->spptr.  `state` is the PetscObjectState of the
 * matrix: a changed state with an unchanged pattern re-uploads the values only.                */
int b200_petsc_ensure_resident(void **slot, int32_t m, int32_t n, const int32_t *ai,
                               const int32_t *aj, const double *aa, int64_t state);
int b200_petsc_invalidate(void **slot);            /* before host compaction (AssemblyEnd)      */
int b200_petsc_release(void **slot);               /* MatDestroy                                */
/* op: 0 = y=Ax, 1 = z=y+Ax, 2 = y=A'x, 3 = y=z+A'x ; host pointers (PETSc 3.7.6 Vec arrays)     */
int b200_petsc_apply_host(void **slot, int op, const double *x, const double *yin, double *yout);
/* same on device pointers (device-resident Vec extension)                                      */
int b200_petsc_apply_device(void **slot, int op, const double *d_x, const double *d_yin,
                            double *d_yout);
void *b200_petsc_handle(void **slot);             /* the b200_csr_t of a resident matrix, or NULL  */
int b200_petsc_mode(void);                         /* B200_MODE env: exact (default)|fast|exact_fma */

#ifdef __cplusplus
}
#endif
#endif
