// pc_impl.h -- private interface between the KSP (dmda_ksp.cpp) and the multigrid
// preconditioner (pcgamg.cpp).
#ifndef B200_PC_IMPL_H
#define B200_PC_IMPL_H
#include "b200_aij.h"

struct B200PCGamg;

// PCSetUp_GAMG + PCSetUp_MG: reads the -pc_gamg_* / -mg_levels_* / -mg_coarse_* options, builds
// the smoothed-aggregation hierarchy of `A` on the host and the level objects (Mat/Vec).
PetscErrorCode b200_pcgamg_setup(Mat A, B200PCGamg **out);
// PCApply_MG: one multiplicative V-cycle, z = M^{-1} r, every matrix operation through the
// SeqAIJ hot path on the device.
PetscErrorCode b200_pcgamg_apply(B200PCGamg *mg, Vec r, Vec z);
PetscErrorCode b200_pcgamg_destroy(B200PCGamg **mg);

// read access for the tests (level 0 = finest)
PetscInt       b200_pcgamg_num_levels(const B200PCGamg *mg);
PetscErrorCode b200_pcgamg_level(const B200PCGamg *mg, PetscInt level, Mat *A, Mat *P, Vec *dinv, const PetscInt **agg,
                                 PetscInt *nagg, PetscReal *emax);

// KSPCG with PCJACOBI run for a fixed number of iterations with KSP_NORM_NONE on a fixed
// pseudo-random right-hand side; the largest eigenvalue of the Lanczos tridiagonal is the
// estimate of lambda_max(D^-1 A) that PCGAMGOptProlongator_AGG uses [P376].  (dmda_ksp.cpp)
PetscErrorCode b200_ksp_estimate_emax(Mat A, PetscInt its, PetscReal *emax);

#endif
