/* spmv_from_c.c -- the C ABI used from plain C99, the way a C caller (PETSc itself is C) binds it:
 * build the 1-D Laplacian tridiag(-1, 2, -1), run y = A x on the GPU through host vectors and through
 * device vectors, compare with the loop of src/openacc-step3/MatMult_SeqAIJ.patch:38-48.
 *   gcc -std=c99 -Iinclude examples/spmv_from_c.c -Lpetsc-openacc_b200 -lb200aij -o spmv_from_c   */
#include <stdio.h>
#include <stdlib.h>

#include "b200_mpiaij.h" /* pulls in b200_seqaij.h; included to prove both headers are C */
#include "b200_petsc_symbols.h"

int main(void)
{
  const int32_t m = 100000;
  int32_t      *ai = malloc(sizeof(int32_t) * (m + 1)), *aj = malloc(sizeof(int32_t) * 3 * m);
  double       *aa = malloc(sizeof(double) * 3 * m), *x = malloc(sizeof(double) * m), *y = malloc(sizeof(double) * m);
  int32_t       nz = 0, i, k;
  int           rc, bad = 0;
  b200_csr_t    A = NULL;
  b200_csr_info_t info;

  ai[0] = 0;
  for (i = 0; i < m; i++) {
    if (i > 0) { aj[nz] = i - 1; aa[nz++] = -1.0; }
    aj[nz] = i; aa[nz++] = 2.0;
    if (i < m - 1) { aj[nz] = i + 1; aa[nz++] = -1.0; }
    ai[i + 1] = nz;
  }
  b200_gen_vector(x, m, 0xB200);
  if ((rc = b200_init(0)) || (rc = b200_csr_create(&A, m, m, ai, aj, aa))) {
    fprintf(stderr, "b200 error %d: %s\n", rc, b200_last_error());
    return 2; /* no GPU: there is no CPU fallback */
  }
  if ((rc = b200_spmv_host(A, x, y, B200_MODE_EXACT))) { fprintf(stderr, "b200 error %d: %s\n", rc, b200_last_error()); return 3; }
  for (i = 0; i < m; i++) {
    double sum = 0.0;
    for (k = ai[i]; k < ai[i + 1]; k++) sum += aa[k] * x[aj[k]];
    if (sum != y[i]) bad++;
  }
  b200_csr_get_info(A, &info);
  printf("spmv_from_c: %d rows, %d nnz, plan kernel %d, %d diagonal codes, mismatches %d, launches %llu\n", (int)info.m,
         (int)info.nz, (int)info.kernel_exact, (int)info.index8_diagonals, bad, (unsigned long long)b200_launch_count());
  b200_csr_destroy(A);
  free(ai); free(aj); free(aa); free(x); free(y);
  return bad ? 1 : 0;
}
