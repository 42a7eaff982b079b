/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY.  Compiles the reference's OWN MatMult row
 * loops (cut from its patch files at build time by extract_ref_loops.py into oracle/_ref as .inc files,
 * never committed) into oracle/_ref/libref_matmult.so, so that the restatement in
 * seqaij_oracle.c (orc_matmult) can be checked against the reference's text bit for bit and the
 * CPU baseline of bench.py can run the reference's loop rather than a port of it.
 *
 * What is the reference's: the loop bodies -- src/openacc-step1/MatMult_SeqAIJ.patch:19-32, old
 * side = PETSc 3.7.6's loop as the patch shows it; src/openacc-step3/MatMult_SeqAIJ.patch:36-70,
 * new side = the author's host loop followed by the device loop; src/openacc-step4/
 * MatMult_SeqAIJ.patch:36-88 = the same with the device loop cut into row blocks.  A plain C
 * compiler ignores their `# pragma acc` lines.
 * What is NOT in the reference and is supplied here: the declarations around the fragments (they
 * are PETSc's, visible only as names in the patches) and the PetscSparseDensePlusDot macro, which
 * lives in PETSc 3.7.6's private headers [P376]; its default variant is restated below, as in
 * seqaij_oracle.c.  The reference's full translation unit cannot be compiled here (PETSc is not
 * available offline).
 *
 * Build: oracle/Makefile (same flags as the oracle: -O2 -ffp-contract=off, no -march).
 */
#include <pthread.h>
#include <stdlib.h>

typedef int    PetscInt;
typedef double PetscScalar;
typedef double MatScalar;

/* [P376] petsc/private/kernels: default variant (scripts/petsc-release.sh sets no unroll macro) */
#define PetscSparseDensePlusDot(sum, r, xv, xi, nnz) \
  { PetscInt __i; for (__i = 0; __i < nnz; __i++) sum += xv[__i] * r[xi[__i]]; }

/* the Mat_SeqAIJ fields the original loop reads through `a->` */
typedef struct { const PetscInt *j; const MatScalar *a; PetscInt nz; } ref_aij_t;

/* PETSc 3.7.6 MatMult_SeqAIJ row loop: text from the old side of the step-1 patch */
void ref_matmult_original(PetscInt m, const PetscInt *ii, const PetscInt *cols, const MatScalar *data,
                          const PetscScalar *x, PetscScalar *y)
{
  ref_aij_t        A_ = {cols, data, 0}, *a = &A_;
  PetscInt         i, n;
  const PetscInt  *aj;
  const MatScalar *aa;
  PetscScalar      sum;
#include "matmult_original_loop.inc"
}

/* Step 3: the author's host loop runs rows 0.. until the asynchronous transfers are done, then the
 * device loop does the rest.  Here the transfers "complete" after host_rows rows. */
#define acc_async_test_all() (offset >= host_rows)
void ref_matmult_step3(PetscInt m, const PetscInt *ii, const PetscInt *cols, const MatScalar *data,
                       const PetscScalar *x, PetscScalar *y, PetscInt host_rows)
{
  PetscInt         i, n;
  const PetscInt  *aj;
  const MatScalar *aa;
  PetscScalar      sum;
#include "matmult_step3_host.inc"
}

/* Step 4: as step 3, but the device part is cut into blocks of 983,040 rows (one kernel launch and
 * one y download per block, src/openacc-step4/MatMult_SeqAIJ.patch:51-72) plus the remaining rows. */
void ref_matmult_step4(PetscInt m, const PetscInt *ii, const PetscInt *cols, const MatScalar *data,
                       const PetscScalar *x, PetscScalar *y, PetscInt host_rows)
{
  PetscInt         i, n;
  const PetscInt  *aj;
  const MatScalar *aa;
  PetscScalar      sum;
#include "matmult_step4_blocked.inc"
}
#undef acc_async_test_all

/* One row block per thread (an MPI rank per core in the reference's runs,
 * runs/single-node-scaling.pbs:56-64), blocks balanced by non-zeros; every block runs the
 * reference's loop. */
typedef struct { PetscInt r0, r1; const PetscInt *ii, *cols; const MatScalar *data; const PetscScalar *x; PetscScalar *y; } ref_job_t;
static void *ref_worker(void *p)
{
  ref_job_t *j = (ref_job_t *)p;
  ref_matmult_original(j->r1 - j->r0, j->ii + j->r0, j->cols, j->data, j->x, j->y + j->r0);
  return NULL;
}
void ref_matmult_mt(int nthreads, PetscInt m, const PetscInt *ii, const PetscInt *cols, const MatScalar *data,
                    const PetscScalar *x, PetscScalar *y)
{
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  ref_job_t job[256];
  const long long nz = m ? ii[m] : 0;
  PetscInt r = 0;
  for (int t = 0; t < nthreads; t++) {
    const long long want = nz * (t + 1) / nthreads;
    PetscInt r1 = r;
    if (t == nthreads - 1) r1 = m;
    else while (r1 < m && ii[r1] < want) r1++;
    job[t] = (ref_job_t){r, r1, ii, cols, data, x, y};
    r = r1;
  }
  for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, ref_worker, &job[t]);
  ref_worker(&job[0]);
  for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
}
