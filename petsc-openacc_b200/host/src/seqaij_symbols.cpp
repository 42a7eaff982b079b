// seqaij_symbols.cpp -- the PETSc-named symbols of the drop-in and the residency hooks under them.
//
// MatMult_SeqAIJ / MatAssemblyEnd_SeqAIJ / MatDestroy_SeqAIJ are the three functions the reference
// cuts out of PETSc 3.7.6's aij.c and replaces (scripts/petsc.sh:81-89, Makefile:153-158);
// MatMultAdd_SeqAIJ / MatMultTranspose[Add]_SeqAIJ are the other MatOps the north star names.
// The bodies use only fields and calls that exist under the same names in PETSc's private aij.h
// (a->i, a->j, a->a, a->nz, a->nonzerorowcnt, a->ilen, a->imax, a->compressedrow, A->rmap->n,
// A->spptr, VecGetArrayRead ...), so that with -DB200_WITH_PETSC the same file compiles against a
// real PETSc 3.7.6 source tree.
#ifdef B200_WITH_PETSC
#include <../src/mat/impls/aij/seq/aij.h>
#else
#include "b200_aij.h"
#endif

#include <cstdlib>
#include <cstring>

#include "../../../include/b200_petsc_symbols.h"
#include "../../../include/b200_seqaij.h"

namespace {

// what Mat->spptr points to
struct Resident {
  b200_csr_t     h     = nullptr;
  const int32_t *ai    = nullptr, *aj = nullptr;
  const double  *aa    = nullptr;
  int32_t        m = 0, n = 0, nz = 0;
  int64_t        state = -1;
};

PetscErrorCode map_err(int rc)
{
  // b200 error codes were chosen inside PETSc's error-code range; pass them through
  return rc;
}

}  // namespace

extern "C" int b200_petsc_mode(void)
{
  const char *s = getenv("B200_MODE");
  if (!s || !*s || !strcmp(s, "exact")) return B200_MODE_EXACT;
  if (!strcmp(s, "fast")) return B200_MODE_FAST;
  if (!strcmp(s, "exact_fma")) return B200_MODE_EXACT_FMA;
  return B200_MODE_EXACT;
}

extern "C" int b200_petsc_ensure_resident(void **slot, int32_t m, int32_t n, const int32_t *ai,
                                          const int32_t *aj, const double *aa, int64_t state)
{
  if (!slot) return B200_ERR_ARG;
  Resident *r = static_cast<Resident *>(*slot);
  const int32_t nz = ai ? ai[m] : 0;
  if (r && r->h && (r->ai != ai || r->aj != aj || r->aa != aa || r->m != m || r->n != n || r->nz != nz)) {
    // the host arrays moved or changed size (re-assembly with new non-zeros): start over
    b200_csr_destroy(r->h);
    r->h = nullptr;
  }
  if (!r) { r = new Resident; *slot = r; }
  if (!r->h) {
    int rc = b200_csr_create(&r->h, m, n, ai, aj, aa);
    if (rc) return rc;
    r->ai = ai; r->aj = aj; r->aa = aa; r->m = m; r->n = n; r->nz = nz; r->state = state;
  } else if (r->state != state) {
    // same pattern, values touched since the last upload (MatSetValues into existing slots,
    // MatZeroRowsColumns, MatScale ...): the reference would silently use stale device values here
    int rc = b200_csr_update_values(r->h, aa);
    if (rc) return rc;
    r->state = state;
  }
  return B200_OK;
}

extern "C" void *b200_petsc_handle(void **slot)
{
  return (slot && *slot) ? (void *)static_cast<Resident *>(*slot)->h : nullptr;
}

extern "C" int b200_petsc_invalidate(void **slot)
{
  if (!slot || !*slot) return B200_OK;
  Resident *r = static_cast<Resident *>(*slot);
  if (r->h) { b200_csr_destroy(r->h); r->h = nullptr; }
  return B200_OK;
}

extern "C" int b200_petsc_release(void **slot)
{
  if (!slot || !*slot) return B200_OK;
  b200_petsc_invalidate(slot);
  delete static_cast<Resident *>(*slot);
  *slot = nullptr;
  return B200_OK;
}

extern "C" int b200_petsc_apply_host(void **slot, int op, const double *x, const double *yin, double *yout)
{
  Resident *r = slot ? static_cast<Resident *>(*slot) : nullptr;
  if (!r || !r->h) return B200_ERR_STATE;
  const int mode = b200_petsc_mode();
  switch (op) {
    case 0: return b200_spmv_host(r->h, x, yout, mode);
    case 1: return b200_spmv_add_host(r->h, x, yin, yout, mode);
    case 2: return b200_spmv_transpose_host(r->h, x, yout, mode);
    case 3: return b200_spmv_transpose_add_host(r->h, x, yin, yout, mode);
  }
  return B200_ERR_ARG;
}

extern "C" int b200_petsc_apply_device(void **slot, int op, const double *x, const double *yin, double *yout)
{
  Resident *r = slot ? static_cast<Resident *>(*slot) : nullptr;
  if (!r || !r->h) return B200_ERR_STATE;
  const int mode = b200_petsc_mode();
  switch (op) {
    case 0: return b200_spmv(r->h, x, yout, mode, nullptr);
    case 1: return b200_spmv_add(r->h, x, yin, yout, mode, nullptr);
    case 2: return b200_spmv_transpose(r->h, x, yout, mode, nullptr);
    case 3: return b200_spmv_transpose_add(r->h, x, yin, yout, mode, nullptr);
  }
  return B200_ERR_ARG;
}

// ---------------------------------------------------------------------------------------------
// the operator-table entries
// ---------------------------------------------------------------------------------------------
#ifndef B200_WITH_PETSC
#define B200_STATE(A) ((int64_t)(A)->state)
#else
#define B200_STATE(A) ((int64_t)((PetscObject)(A))->state)
#endif

// common body: op as in b200_petsc_apply_*; in = the vector read by the product, acc = the
// accumulator input (NULL for op 0/2), out = the result
static PetscErrorCode seqaij_apply(Mat A, int op, Vec in, Vec acc, Vec out)
{
  Mat_SeqAIJ    *a = (Mat_SeqAIJ *)A->data;
  PetscErrorCode ierr;
  int            rc;

  rc = b200_petsc_ensure_resident(&A->spptr, A->rmap->n, A->cmap->n, a->i, a->j, a->a, B200_STATE(A));
  if (rc) return map_err(rc);
#ifndef B200_WITH_PETSC
  {
    // device-resident Vec extension: stay in HBM when the input's device copy is current
    PetscBool on_dev = PETSC_FALSE;
    ierr = VecB200HasDevice(in, &on_dev);CHKERRQ(ierr);
    if (on_dev) {
      const PetscScalar *dx, *dacc = NULL;
      PetscScalar       *dout;
      ierr = VecB200GetDeviceArrayRead(in, &dx);CHKERRQ(ierr);
      if (acc && acc != out) { ierr = VecB200GetDeviceArrayRead(acc, &dacc);CHKERRQ(ierr); }
      if (acc && acc == out) { ierr = VecB200GetDeviceArray(out, &dout);CHKERRQ(ierr); dacc = dout; }
      else { ierr = VecB200GetDeviceArrayWrite(out, &dout);CHKERRQ(ierr); }
      rc = b200_petsc_apply_device(&A->spptr, op, dx, dacc, dout);
      if (rc) return map_err(rc);
      goto logflops;
    }
  }
#endif
  {
    const PetscScalar *x, *yin = NULL;
    PetscScalar       *y;
    ierr = VecGetArrayRead(in, &x);CHKERRQ(ierr);
    if (acc && acc != out) { ierr = VecGetArrayRead(acc, &yin);CHKERRQ(ierr); }
    ierr = VecGetArray(out, &y);CHKERRQ(ierr);
    if (acc && acc == out) yin = y;
    rc = b200_petsc_apply_host(&A->spptr, op, x, yin, y);
    ierr = VecRestoreArrayRead(in, &x);CHKERRQ(ierr);
    if (acc && acc != out) { ierr = VecRestoreArrayRead(acc, &yin);CHKERRQ(ierr); }
    ierr = VecRestoreArray(out, &y);CHKERRQ(ierr);
    if (rc) return map_err(rc);
  }
#ifndef B200_WITH_PETSC
logflops:
#endif
  // flop accounting of the original functions (src/openacc-step2/MatMult_SeqAIJ.patch:47)
  if (op == 0) { ierr = PetscLogFlops(2.0 * a->nz - a->nonzerorowcnt);CHKERRQ(ierr); }
  else { ierr = PetscLogFlops(2.0 * a->nz);CHKERRQ(ierr); }
  return 0;
}

extern "C" PetscErrorCode MatMult_SeqAIJ(Mat A, Vec xx, Vec yy) { return seqaij_apply(A, 0, xx, NULL, yy); }
extern "C" PetscErrorCode MatMultAdd_SeqAIJ(Mat A, Vec xx, Vec yy, Vec zz) { return seqaij_apply(A, 1, xx, yy, zz); }
extern "C" PetscErrorCode MatMultTranspose_SeqAIJ(Mat A, Vec xx, Vec yy) { return seqaij_apply(A, 2, xx, NULL, yy); }
extern "C" PetscErrorCode MatMultTransposeAdd_SeqAIJ(Mat A, Vec xx, Vec zz, Vec yy) { return seqaij_apply(A, 3, xx, zz, yy); }

#ifndef B200_WITH_PETSC
// The two functions below are compiled for the OFFLINE API slice only.  Against a real PETSc 3.7.6
// tree (-DB200_WITH_PETSC) they must stay PETSc's own bodies -- MatDestroy_SeqAIJ also destroys the
// row/col index sets, diag/imax/ilen/idiag/solve_work/saved_values, the inode data and the composed
// functions; MatAssemblyEnd_SeqAIJ also sets rmax / the Info counters / the inode check -- with only
// the residency hook lines added, exactly as the reference patches them
// (src/openacc-step2/MatDestroy_SeqAIJ.patch, MatAssemblyEnd_SeqAIJ.patch; the hook lines are spelled
// out in INTEGRATION.md section 3).  Defining look-alikes here would leak every SeqAIJ's private data.
//
// MatAssemblyEnd_SeqAIJ: PETSc 3.7.6 aij.c:973-1032 (scripts/petsc.sh:81).  The compaction below
// follows the published algorithm (head visible at src/openacc-step2/
// MatAssemblyEnd_SeqAIJ.patch:34-37: rows keep imax[i] reserved slots of which ilen[i] are used;
// every row is moved back by the unused slots before it).  Residency: the reference drops the
// device aj/aa before and re-uploads after when present (:21-29, :42-44); here the device mirror
// is invalidated before the host arrays are rewritten and rebuilt lazily by the next MatMult.
extern "C" PetscErrorCode MatAssemblyEnd_SeqAIJ(Mat A, MatAssemblyType mode)
{
  Mat_SeqAIJ    *a = (Mat_SeqAIJ *)A->data;
  PetscErrorCode ierr;
  PetscInt       fshift = 0, i, j, *ai = a->i, *aj = a->j, *imax = a->imax;
  PetscInt       m = A->rmap->n, *ip, N, *ailen = a->ilen, rmax = 0;
  MatScalar     *aa = a->a, *ap;
  PetscReal      ratio = 0.6;

  if (mode == MAT_FLUSH_ASSEMBLY) return 0;

  // is anything going to move?  (pure value updates keep the mirror and only bump the state)
  PetscBool moves = PETSC_FALSE;
  for (i = 0; i < m; i++) if (imax[i] != ailen[i]) { moves = PETSC_TRUE; break; }
  if (moves) { int rc = b200_petsc_invalidate(&A->spptr); if (rc) return rc; }

  if (m) rmax = ailen[0];
  for (i = 1; i < m; i++) {
    fshift += imax[i - 1] - ailen[i - 1];
    rmax = PetscMax(rmax, ailen[i]);
    if (fshift) {
      ip = aj + ai[i];
      ap = aa + ai[i];
      N  = ailen[i];
      for (j = 0; j < N; j++) {
        ip[j - fshift] = ip[j];
        ap[j - fshift] = ap[j];
      }
    }
    ai[i] = ai[i - 1] + ailen[i - 1];
  }
  if (m) {
    fshift += imax[m - 1] - ailen[m - 1];
    ai[m] = ai[m - 1] + ailen[m - 1];
  }
  a->nonzerorowcnt = 0;
  for (i = 0; i < m; i++) {
    ailen[i] = imax[i] = ai[i + 1] - ai[i];
    a->nonzerorowcnt += ((ai[i + 1] - ai[i]) > 0);
  }
  a->nz = m ? ai[m] : 0;
  if (fshift && a->nounused == -1)
    SETERRQ3(PETSC_COMM_SELF, 77, "Unused space detected in matrix: %d X %d, %d unneeded", m, A->cmap->n, fshift);

  ierr = MatMarkDiagonal_SeqAIJ(A);CHKERRQ(ierr);
  A->info.mallocs += a->reallocs;
  a->reallocs = 0;
  A->info.nz_unneeded = (PetscReal)fshift;
  a->rmax = rmax;

  ierr = MatCheckCompressedRow(A, a->nonzerorowcnt, &a->compressedrow, a->i, m, ratio);CHKERRQ(ierr);
  ierr = MatAssemblyEnd_SeqAIJ_Inode(A, mode);CHKERRQ(ierr);
  ierr = MatSeqAIJInvalidateDiagonal(A);CHKERRQ(ierr);
  return 0;
}

// MatDestroy_SeqAIJ: PETSc 3.7.6 aij.c:1076-1121 (scripts/petsc.sh:83).  Device mirrors go first
// (src/openacc-step2/MatDestroy_SeqAIJ.patch:26-34), then the host CSR (:36).
extern "C" PetscErrorCode MatDestroy_SeqAIJ(Mat A)
{
  Mat_SeqAIJ    *a = (Mat_SeqAIJ *)A->data;
  PetscErrorCode ierr;
  int            rc = b200_petsc_release(&A->spptr);
  if (rc) return rc;
  ierr = MatSeqXAIJFreeAIJ(A, &a->a, &a->j, &a->i);CHKERRQ(ierr);
  free(a->diag); free(a->imax); free(a->ilen);
  free(a->compressedrow.i); free(a->compressedrow.rindex);
  free(a);
  A->data = NULL;
  return 0;
}
#endif  // !B200_WITH_PETSC
