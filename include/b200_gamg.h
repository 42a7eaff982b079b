/*
 * b200_gamg.h -- host-side building blocks of the multigrid set-up behind a C ABI (libb200petsc.so):
 * a host CSR handle with the sparse products the Galerkin operator needs, and one coarsening step
 * (strength graph, aggregates, smoothed prolongator) on a square block.  They are the kernels of
 * `-pc_type gamg` (petsc-openacc_b200/host/src/pcgamg.cpp, which restates PCGAMG "agg" of PETSc
 * 3.7.6 [P376] for the reference's configs/PETSc_SolverOptions_GAMG.info) exposed so that the
 * row-partitioned set-up (petsc-openacc_b200/dgamg.py, one process per GPU, BASELINE configs[2]) can
 * run them per rank.  No device code; every function returns 0 or a PetscErrorCode.
 */
#ifndef B200_GAMG_H
#define B200_GAMG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_hcsr_s *b200_hcsr_t;   /* host CSR: int32 indices, fp64 values, ascending columns */

int b200_hcsr_create(b200_hcsr_t *out, int32_t m, int32_t n, const int32_t *ai, const int32_t *aj,
                     const double *aa);                                   /* arrays are copied      */
int b200_hcsr_destroy(b200_hcsr_t h);
int b200_hcsr_shape(b200_hcsr_t h, int32_t *m, int32_t *n, int32_t *nz);
int b200_hcsr_arrays(b200_hcsr_t h, const int32_t **ai, const int32_t **aj, const double **aa); /* borrowed */
int b200_hcsr_spgemm(b200_hcsr_t X, b200_hcsr_t Y, b200_hcsr_t *out);   /* X Y, entries summed in storage order */
int b200_hcsr_transpose(b200_hcsr_t X, b200_hcsr_t *out);               /* ascending rows inside a column */
int b200_hcsr_add(b200_hcsr_t X, b200_hcsr_t Y, b200_hcsr_t *out);      /* X + Y, same shape         */
int b200_hcsr_abs_row_sums(b200_hcsr_t A, double *out);                 /* sum_j |a_ij| per row      */

/* One coarsening step on the square block A with near-null-space vector B[m]:
 * agg[m] (aggregate of each vertex, -1 = no strong neighbour), *nagg, P (m x nagg; smoothed with
 * 1.4/emax D^-1 A when emax > 0, tentative otherwise), Bc[nagg] (the coarse near-null-space vector;
 * the caller provides room for m entries). */
int b200_gamg_coarsen_block(b200_hcsr_t A, const double *B, double threshold, int square, double emax,
                            int32_t *agg, int32_t *nagg, b200_hcsr_t *P, double *Bc);

#ifdef __cplusplus
}
#endif
#endif /* B200_GAMG_H */
