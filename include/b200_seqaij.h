/*
 * b200_seqaij.h -- C ABI of the Blackwell-native (sm_100a) SeqAIJ sparse mat-vec hot path.
 *
 * This is the drop-in boundary underneath PETSc's per-type operator table: the five C symbols the
 * reference (olcf/PETSC-OpenACC) substitutes at link time --
 *     MatMult_SeqAIJ        src/openacc-step3/MatMult_SeqAIJ.patch:12
 *     MatAssemblyEnd_SeqAIJ src/openacc-step2/MatAssemblyEnd_SeqAIJ.patch:12
 *     MatDestroy_SeqAIJ     src/openacc-step2/MatDestroy_SeqAIJ.patch:12
 * plus MatMultAdd_SeqAIJ / MatMultTranspose_SeqAIJ (PETSc 3.7.6 aij.c, named by the north star) --
 * are thin shims over the functions below (see include/b200_petsc_symbols.h and INTEGRATION.md).
 *
 * Conventions
 *   - plain C, plain pointers and sizes; PetscInt = int32_t, PetscScalar = MatScalar = double
 *     (scripts/petsc-release.sh:6,62 and no --with-64-bit-indices);
 *   - every function returns 0 on success or a non-zero b200 error code (never aborts), mirroring
 *     PetscErrorCode propagation (src/openacc-step1/MatMult_SeqAIJ.patch:16);
 *   - "h_" pointers are host memory, "d_" pointers are device memory of the current device;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - there is NO CPU fallback: without a usable sm_100 device every compute entry returns
 *     B200_ERR_NO_DEVICE.
 */
#ifndef B200_SEQAIJ_H
#define B200_SEQAIJ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes ------------------------------------------------------------------------- */
enum {
  B200_OK             = 0,
  B200_ERR_ARG        = 60,  /* bad argument (PETSC_ERR_ARG_* range is 56..63 in PETSc)        */
  B200_ERR_NO_DEVICE  = 92,  /* no CUDA device / not sm_100 (PETSC_ERR_LIB-like)               */
  B200_ERR_CUDA       = 76,  /* a CUDA runtime call failed; see b200_last_error()              */
  B200_ERR_MEM        = 55,  /* allocation failed (PETSC_ERR_MEM)                              */
  B200_ERR_STATE      = 73,  /* object in wrong state (PETSC_ERR_ARG_WRONGSTATE)               */
  B200_ERR_TIMEOUT    = 77   /* a peer flag was not seen within the spin budget (halo)         */
};

/* ---- summation-order modes ---------------------------------------------------------------- */
/* The reference sums each row strictly left to right (PetscSparseDensePlusDot,
 * src/openacc-step1/MatMult_SeqAIJ.patch:30).
 *   EXACT      one accumulator per row, left to right, multiply and add rounded separately
 *              (bit-exact against the oracle built with -ffp-contract=off);
 *   EXACT_FMA  same order with a fused multiply-add (bit-exact against the oracle's *_fma);
 *   FAST       no order promised (|y - y_ref| <= 1e-13 * sum_j |a_ij x_j| per row).  With the
 *              default plans FAST runs the same one-accumulator-per-row kernels as EXACT_FMA /
 *              EXACT -- they are also the fastest measured -- so it is bit-reproducible too; only
 *              the VECTOR / split-row MERGE overrides sum a row with several lanes.            */
enum { B200_MODE_FAST = 0, B200_MODE_EXACT = 1, B200_MODE_EXACT_FMA = 2 };

/* Kernel override for tests and sweeps (AUTO = choose from the histogram). */
enum {
  B200_KERNEL_AUTO   = 0,
  B200_KERNEL_ROW    = 1,  /* thread per row straight from global memory (the reference's shape) */
  B200_KERNEL_STREAM = 2,  /* TMA-bulk-staged CSR tiles, thread per row out of shared memory     */
  B200_KERNEL_VECTOR = 3,  /* sub-warp per row, __shfl_xor reduction                             */
  B200_KERNEL_MERGE  = 4,  /* skewed rows: warp-granular chunks of <= 128 non-zeros, rows summed in CSR
                              order (k_wmerge; in column blocks when x exceeds 60 MB);
                              B200_MERGE_SPLIT=1 + MODE_FAST: split-row tiles with carry fix-up   */
  B200_KERNEL_CPROW  = 5,  /* compressed-row: only the non-empty rows                            */
  B200_KERNEL_SELL   = 6   /* the optional SELL-32-sigma copy (b200_csr_build_sell); never chosen by
                              AUTO                                                               */
};

typedef struct b200_csr_s *b200_csr_t;

typedef struct {
  int32_t m, n, nz;
  int32_t nonzerorowcnt;     /* rows with at least one entry (a->nonzerorowcnt)                 */
  int32_t rmax;              /* longest row (a->rmax)                                           */
  int32_t compressedrow_use; /* MatCheckCompressedRow verdict, ratio 0.6                        */
  int32_t cprow_nrows;
  int32_t kernel_fast;       /* B200_KERNEL_* chosen for MODE_FAST                              */
  int32_t kernel_exact;      /* B200_KERNEL_* chosen for MODE_EXACT*                            */
  int32_t vector_lanes;      /* lanes per row of the VECTOR kernel                              */
  int32_t stream_tiles;      /* number of tiles of the STREAM kernel (0 = not applicable)       */
  int32_t merge_tiles;
  int32_t has_transpose;     /* explicit transpose copy is resident                             */
  int32_t index8_diagonals;  /* > 0: column indices stream as 1-byte codes over this many diagonals */
  int32_t hist[16];          /* row-length histogram: [0],[1],[2],[3-4],[5-8],...,[>16384]      */
  uint64_t device_bytes;     /* HBM held by this handle                                         */
  int32_t sell_chunks;       /* > 0: a SELL-32-sigma copy is resident (chunks of 32 rows)          */
  int32_t sell_sigma;        /* its sorting window                                               */
  uint64_t sell_padded_nnz;  /* entries it stores, padding included                              */
  int32_t rowlen8;           /* 1: the stream plan reads 1-byte row lengths instead of 4-byte row pointers */
  int32_t reserved;
} b200_csr_info_t;

/* ---- runtime ------------------------------------------------------------------------------ */
int         b200_init(int device);           /* cudaSetDevice + capability check (sm_100)       */
const char *b200_last_error(void);           /* text of the last failure on this thread         */
uint64_t    b200_launch_count(void);         /* kernels of this library launched so far         */
int         b200_device_sm_count(void);
const char *b200_version(void);

/* ---- CSR residency (replaces the OpenACC enter/exit data directives,
 *      src/openacc-step2/MatMult_SeqAIJ.patch:19-21, MatAssemblyEnd...:28-29,42-44,
 *      MatDestroy...:26-34) ---------------------------------------------------------------- */
/* Mirror host CSR (a->i, a->j, a->a) once into device-resident, 128-bit aligned arrays and
 * build the per-matrix kernel plan (row histogram, tile tables, compressed-row index).       */
int b200_csr_create(b200_csr_t *out, int32_t m, int32_t n, const int32_t *h_ai,
                    const int32_t *h_aj, const double *h_aa);
/* Same from arrays that already live on the device (copied; the caller keeps ownership).     */
int b200_csr_create_from_device(b200_csr_t *out, int32_t m, int32_t n, const int32_t *d_ai,
                                const int32_t *d_aj, const double *d_aa);
/* Values changed, pattern did not (MatAssemblyEnd after MatSetValues into existing slots,
 * MatZeroRowsColumns, MatScale ...): re-upload aa only.                                      */
int b200_csr_update_values(b200_csr_t A, const double *h_aa);
int b200_csr_destroy(b200_csr_t A);
int b200_csr_get_info(b200_csr_t A, b200_csr_info_t *info);
int b200_csr_set_kernel(b200_csr_t A, int kernel); /* B200_KERNEL_* override, AUTO resets     */
/* Build (or drop) the explicit transpose used by the deterministic MatMultTranspose.         */
int b200_csr_build_transpose(b200_csr_t A);
/* Optional SELL-32-sigma (sliced ELLPACK) copy for stencil-like matrices: chunks of 32 rows
 * stored column-major and padded to the chunk's longest row; inside windows of `sigma` rows the
 * rows are ordered by decreasing length first (sigma = 1: natural order, no permutation array).
 * With a byte-code plan (<= 254 diagonals) the copy holds 1-byte codes too.  Used by
 * b200_csr_set_kernel(A, B200_KERNEL_SELL); dropped by b200_csr_update_values.  Padding is
 * skipped, so every mode sums a row left to right exactly like the CSR kernels.               */
int b200_csr_build_sell(b200_csr_t A, int32_t sigma);
/* The packing itself, host only (no device needed): sizes first, then the arrays
 * cs[nchunks + 1] (chunk starts, in entries), perm[nchunks * 32] (slot -> row, -1 = padding slot),
 * val[padded], col[padded] (-1 = padding entry).                                              */
int b200_sell_pack_size(int32_t m, const int32_t *h_ai, int32_t sigma, int32_t *nchunks,
                        uint64_t *padded);
int b200_sell_pack(int32_t m, const int32_t *h_ai, const int32_t *h_aj, const double *h_aa,
                   int32_t sigma, uint32_t *cs, int32_t *perm, double *val, int32_t *col);
/* The plan of the warp-granular exact-order kernel (k_wmerge, skewed row lengths), host only:
 * chunks4[4*c..] = {first row, rows (>= 0) | -2 first / -1 middle / -3 last piece of a row longer
 * than 128 entries, first non-zero, last non-zero + 1}; blk[b], blk[b+1] = the chunks of work block b
 * (a long row's pieces never straddle two blocks).  Sizes first, then the arrays.               */
int b200_wmerge_plan_size(int32_t m, const int32_t *h_ai, int32_t *nchunks, int32_t *nblocks);
int b200_wmerge_plan(int32_t m, const int32_t *h_ai, int32_t *chunks4, int32_t *blk);
/* Column blocks of the skewed plan (x larger than 1.5 x limit_bytes: A = [A_0 | A_1 | ...], y = A_0 x,
 * y += A_1 x, ...), host only: *nblocks (0 = no blocking) and, when split != NULL, for every block b
 * and row r the first entry of row r that belongs to a later block: split[b * m + r].           */
int b200_colblock_split(int32_t m, int32_t n, const int32_t *h_ai, const int32_t *h_aj,
                        int64_t limit_bytes, int32_t *nblocks, int32_t *split);
/* Device pointers of the mirrors (for tests / composition), any may be NULL.                 */
int b200_csr_device_arrays(b200_csr_t A, const int32_t **d_ai, const int32_t **d_aj,
                           const double **d_aa);

/* ---- the hot path, device-resident vectors ------------------------------------------------ */
/* y = A x                      MatMult_SeqAIJ                                                 */
int b200_spmv(b200_csr_t A, const double *d_x, double *d_y, int mode, void *stream);
/* z = y + A x (z may alias y)  MatMultAdd_SeqAIJ                                              */
int b200_spmv_add(b200_csr_t A, const double *d_x, const double *d_y, double *d_z, int mode,
                  void *stream);
/* y = A^T x                    MatMultTranspose_SeqAIJ                                        */
int b200_spmv_transpose(b200_csr_t A, const double *d_x, double *d_y, int mode, void *stream);
/* y = z + A^T x                MatMultTransposeAdd_SeqAIJ                                     */
int b200_spmv_transpose_add(b200_csr_t A, const double *d_x, const double *d_z, double *d_y,
                            int mode, void *stream);

/* Fused epilogues for the multigrid levels the solver options ask for (richardson(1) + jacobi,
 * configs/PETSc_SolverOptions_GAMG.info:15-21): one pass instead of MatMult + VecAYPX (+
 * VecPointwiseMult + VecAXPY).  Every step rounds separately, like the separate PETSc calls.   */
/* r = b - A x                                                                                  */
int b200_spmv_residual(b200_csr_t A, const double *d_x, const double *d_b, double *d_r, int mode,
                       void *stream);
/* xnew = x + dinv .* (b - A x)   (xnew must not alias x)                                       */
int b200_spmv_jacobi_sweep(b200_csr_t A, const double *d_x, const double *d_b,
                           const double *d_dinv, double *d_xnew, int mode, void *stream);

/* ---- the hot path, HOST vectors (what MatMult_SeqAIJ(Mat,Vec,Vec) sees in PETSc 3.7.6) ----
 * x is uploaded, y downloaded, synchronous on return (the reference ends with `acc wait`,
 * src/openacc-step4/MatMult_SeqAIJ.patch:91).  Row-blocked and pipelined over three streams
 * like the reference's step 4 (:51-72) when the matrix is banded.                             */
int b200_spmv_host(b200_csr_t A, const double *h_x, double *h_y, int mode);
int b200_spmv_add_host(b200_csr_t A, const double *h_x, const double *h_y, double *h_z, int mode);
int b200_spmv_transpose_host(b200_csr_t A, const double *h_x, double *h_y, int mode);
int b200_spmv_transpose_add_host(b200_csr_t A, const double *h_x, const double *h_z, double *h_y,
                                int mode);
/* Pinned host allocation for Vec arrays (page-locked memory makes the copies above async).   */
int b200_host_alloc(void **p, size_t bytes);
int b200_host_free(void *p);
int b200_host_register(void *p, size_t bytes);
int b200_host_unregister(void *p);

/* ---- CG vector kernels (KSPSolve_CG's VecDot/VecNorm/VecAXPY/VecAYPX, fused) -------------- */
/* Results of reductions are written to device memory (d_out) so a solve never syncs the host
 * unless asked to.  All reductions are deterministic (fixed grid, fixed tree).               */
int b200_vec_set(double *d_x, double a, int64_t n, void *stream);
int b200_vec_copy(double *d_y, const double *d_x, int64_t n, void *stream);
int b200_vec_axpy(double *d_y, double a, const double *d_x, int64_t n, void *stream);   /* y+=a x */
int b200_vec_aypx(double *d_y, double a, const double *d_x, int64_t n, void *stream);   /* y=x+a y */
int b200_vec_pointwise_mult(double *d_w, const double *d_x, const double *d_y, int64_t n,
                            void *stream);
int b200_vec_dot(const double *d_x, const double *d_y, int64_t n, double *d_out, void *stream);
int b200_vec_norm2(const double *d_x, int64_t n, double *d_out, void *stream);
int b200_vec_norm_inf(const double *d_x, int64_t n, double *d_out, void *stream);
int b200_vec_sum(const double *d_x, int64_t n, double *d_out, void *stream);

/* PETSc-free CG (KSPCG semantics: left-preconditioned, preconditioned-residual norm, Jacobi),
 * everything device-resident; only the scalar norm is read back once per iteration.           */
typedef struct {
  int32_t its;        /* iterations done                                                       */
  int32_t reason;     /* >0 converged (2 = rtol, 3 = atol), <0 diverged (-3 = max_it)          */
  double  rnorm;      /* last preconditioned residual norm                                     */
  double  rnorm0;
  double  solve_ms;   /* device time of the solve                                              */
  uint64_t launches;
} b200_cg_result_t;
int b200_cg_jacobi(b200_csr_t A, const double *d_b, double *d_x, double rtol, double atol,
                   int32_t max_it, int mode, b200_cg_result_t *res, void *stream);

/* ---- synthetic workloads (device-independent, counter based) ----------------- */
/* uniform [-1,1) from splitmix64(seed ^ i) */
int b200_gen_vector(double *h_x, int64_t n, uint64_t seed);
/* The reference problem (src/helper.cpp:161-279: generateA + setRefPoint, :78-157 rhs/exact) for
 * one rank of the DMDA decomposition PETSC_DECIDE picks for `size` ranks; PETSc ordering, global
 * column ids ascending per row.  info[12] = m n p xs ys zs xm ym zm nloc rstart nnz.          */
int b200_gen_poisson7_info(int M, int N, int P, int size, int rank, int32_t *info);
int b200_gen_poisson7_bases(int M, int N, int P, int size, int32_t *base /* size+1 */);
int b200_gen_poisson7(int M, int N, int P, int size, int rank, int refpoint, int32_t *ai,
                      int32_t *aj_global, double *aa, double *rhs, double *exact);

/* BASELINE configs[4]: power-law row lengths 1..lmax (alpha = 2 -> mean ~ ln lmax), columns = sorted
 * unique splitmix64 draws mod n, values uniform [-1,1).  First the row pointers (ai[m+1]), then the
 * fill into arrays of ai[m] entries.                                                              */
int b200_gen_powerlaw_rowptr(int32_t m, int32_t n, double alpha, int32_t lmax, uint64_t seed,
                             int32_t *ai);
int b200_gen_powerlaw_fill(int32_t m, int32_t n, double alpha, int32_t lmax, uint64_t seed,
                           const int32_t *ai, int32_t *aj, double *aa);

/* BASELINE configs[3]: 27-point box stencil on an N^3 grid (non-periodic, natural ordering, nnz =
 * (3N-2)^3): off-diagonal -1, diagonal = number of neighbours, or uniform [-1,1) values when seed != 0.
 * ai[N^3 + 1] is always filled; aj / aa may be NULL for a sizing call.                               */
int b200_gen_stencil27(int32_t N, uint64_t seed, int32_t *ai, int32_t *aj, double *aa);

#ifdef __cplusplus
}
#endif
#endif /* B200_SEQAIJ_H */
