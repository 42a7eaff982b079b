/*
 * oracle/seqaij_oracle.c -- TEST INFRASTRUCTURE ONLY.  Not shipped, not on the product path.
 *
 * A plain-C CPU restatement of the reference's SeqAIJ sparse mat-vec hot path and of the
 * host-side integer work around it.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the checker or as the
 * timed CPU baseline.  The product (petsc-openacc_b200/) never links or calls it.
 *
 * PINNED: orc_matmult equals, bit for bit, the reference's own row loops compiled from its patch
 * files (oracle/_ref/libref_matmult.so: extract_ref_loops.py + ref_harness.c; tests/test_oracle.py);
 * orc_matmultadd, orc_matmulttranspose, orc_matmulttransposeadd and the compressed-row variants
 * equal that same compiled loop run on equivalent matrices built independently with scipy -- A^T as
 * CSR with ascending columns, [I | A] applied to [y; x], the non-empty rows scattered through rindex
 * (tests/test_oracle.py::test_transpose_add_and_compressed_row_pinned_to_the_reference_loop);
 * and the generator below equals the reference's own src/helper.cpp compiled from where it lies
 * (oracle/_ref/libref_helper.so; tests/test_host_layer.py).
 * PARITY UNPINNED for everything else (MPIAIJ set-up, DMDA grids, the multigrid, CG's reduction
 * order): the reference (olcf/PETSC-OpenACC) ships no golden
 * vectors, no tests and no logs, and the rest of the arithmetic lives in PETSc 3.7.6
 * (petsc-lite-3.7.6.tar.gz, sha1 f2310cc0663848cbdcdf2ddf8ac48246a43d336b, scripts/petsc.sh:39,45)
 * which is downloaded at build time and is neither under /root/reference nor installable here (no
 * network, no MPI).  What this file follows:
 *   - the MatMult loop is visible verbatim as context lines of the reference's own patches
 *     (src/openacc-step1/MatMult_SeqAIJ.patch:22-32) and is restated by the reference author in
 *     plain C at src/openacc-step3/MatMult_SeqAIJ.patch:38-48; orc_matmult follows those lines;
 *   - the flop count follows src/openacc-step2/MatMult_SeqAIJ.patch:47;
 *   - the assembly compaction follows the context lines of
 *     src/openacc-step2/MatAssemblyEnd_SeqAIJ.patch:31-37,77-79;
 *   - the problem generator follows src/helper.cpp:78-279 line by line;
 *   - the reference's only known-answer test (analytic solution, src/main_ksp.cpp:5-15,120-121)
 *     is run against this generator in tests/test_oracle.py (test_known_answer_analytic_solution).
 * MatMultAdd / MatMultTranspose / compressed-row / MPIAIJ setup have NO text in the reference;
 * they restate PETSc 3.7.6's published algorithm from its documented behaviour ("[P376]" below);
 * the first three are tied to the reference's MatMult loop as said above.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: strict left-to-right, unfused).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------- */
/* A3: PetscSparseDensePlusDot, default variant: sum += xv[k]*r[xi[k]] for k = 0..n-1           */
/* (used at src/openacc-step1/MatMult_SeqAIJ.patch:30).  Canonical semantics here = UNFUSED      */
/* multiply then add (this TU is compiled with -ffp-contract=off).                               */
/* ------------------------------------------------------------------------------------------- */
static inline double sparse_dense_plus_dot(double sum, const double *r, const double *xv,
                                           const int *xi, int n)
{
  for (int k = 0; k < n; k++) sum += xv[k] * r[xi[k]];
  return sum;
}

/* Same traversal with a fused multiply-add: what a contracting compiler (PGI -fast on an FMA4
 * Opteron, gcc -ffp-contract=fast) would emit for the reference loop.  Reported separately. */
#if defined(__x86_64__)
__attribute__((target("fma")))
static double spdot_fma_hw(double sum, const double *r, const double *xv, const int *xi, int n)
{
  for (int k = 0; k < n; k++) sum = __builtin_fma(xv[k], r[xi[k]], sum);
  return sum;
}
#endif
static double spdot_fma_sw(double sum, const double *r, const double *xv, const int *xi, int n)
{
  for (int k = 0; k < n; k++) sum = fma(xv[k], r[xi[k]], sum);
  return sum;
}
typedef double (*spdot_fn)(double, const double *, const double *, const int *, int);
static spdot_fn pick_fma(void)
{
#if defined(__x86_64__)
  if (__builtin_cpu_supports("fma")) return spdot_fma_hw;
#endif
  return spdot_fma_sw;
}

/* ------------------------------------------------------------------------------------------- */
/* A1: MatMult_SeqAIJ, non-compressed branch.                                                    */
/* Follows src/openacc-step1/MatMult_SeqAIJ.patch:22-32 (context lines = PETSc 3.7.6 text) and   */
/* the reference author's host restatement src/openacc-step3/MatMult_SeqAIJ.patch:38-48.         */
/* ------------------------------------------------------------------------------------------- */
void orc_matmult(int m, const int *ii, const int *aj, const double *aa, const double *x, double *y)
{
  for (int i = 0; i < m; i++) {
    int           n   = ii[i + 1] - ii[i];
    const int    *cj  = aj + ii[i];
    const double *cv  = aa + ii[i];
    double        sum = 0.0;
    sum  = sparse_dense_plus_dot(sum, x, cv, cj, n);
    y[i] = sum;
  }
}

void orc_matmult_fma(int m, const int *ii, const int *aj, const double *aa, const double *x,
                     double *y)
{
  spdot_fn f = pick_fma();
  for (int i = 0; i < m; i++) y[i] = f(0.0, x, aa + ii[i], aj + ii[i], ii[i + 1] - ii[i]);
}

/* A2: MatMult_SeqAIJ, compressed-row branch [P376]: zero y, then only the non-empty rows. */
void orc_matmult_cprow(int m, int nrows, const int *cpi, const int *ridx, const int *aj,
                       const double *aa, const double *x, double *y)
{
  memset(y, 0, (size_t)m * sizeof(double));
  for (int i = 0; i < nrows; i++) {
    int    n   = cpi[i + 1] - cpi[i];
    double sum = 0.0;
    sum        = sparse_dense_plus_dot(sum, x, aa + cpi[i], aj + cpi[i], n);
    y[ridx[i]] = sum;
  }
}

/* A7: MatMultAdd_SeqAIJ [P376]: z = y + A x; the accumulator STARTS at y[i]. */
void orc_matmultadd(int m, const int *ii, const int *aj, const double *aa, const double *x,
                    const double *y, double *z)
{
  for (int i = 0; i < m; i++) {
    double sum = y[i];
    sum  = sparse_dense_plus_dot(sum, x, aa + ii[i], aj + ii[i], ii[i + 1] - ii[i]);
    z[i] = sum;
  }
}

void orc_matmultadd_fma(int m, const int *ii, const int *aj, const double *aa, const double *x,
                        const double *y, double *z)
{
  spdot_fn f = pick_fma();
  for (int i = 0; i < m; i++) z[i] = f(y[i], x, aa + ii[i], aj + ii[i], ii[i + 1] - ii[i]);
}

/* A7, compressed-row variant [P376]: copy y to z when they differ, then update non-empty rows. */
void orc_matmultadd_cprow(int m, int nrows, const int *cpi, const int *ridx, const int *aj,
                          const double *aa, const double *x, const double *y, double *z)
{
  if (z != y) memcpy(z, y, (size_t)m * sizeof(double));
  for (int i = 0; i < nrows; i++) {
    double sum = y[ridx[i]];
    sum        = sparse_dense_plus_dot(sum, x, aa + cpi[i], aj + cpi[i], cpi[i + 1] - cpi[i]);
    z[ridx[i]] = sum;
  }
}

/* A8: MatMultTranspose_SeqAIJ = VecSet(y,0) + MatMultTransposeAdd [P376]:
 * for each row i in ascending order: alpha = x[i]; y[aj[k]] += alpha*aa[k]. */
void orc_matmulttransposeadd(int m, int n, const int *ii, const int *aj, const double *aa,
                             const double *x, const double *z, double *y)
{
  if (z != y) memcpy(y, z, (size_t)n * sizeof(double));
  for (int i = 0; i < m; i++) {
    double alpha = x[i];
    for (int k = ii[i]; k < ii[i + 1]; k++) y[aj[k]] += alpha * aa[k];
  }
}

void orc_matmulttranspose(int m, int n, const int *ii, const int *aj, const double *aa,
                          const double *x, double *y)
{
  memset(y, 0, (size_t)n * sizeof(double));
  orc_matmulttransposeadd(m, n, ii, aj, aa, x, y, y);
}

void orc_matmulttranspose_fma(int m, int n, const int *ii, const int *aj, const double *aa,
                              const double *x, double *y)
{
  memset(y, 0, (size_t)n * sizeof(double));
  for (int i = 0; i < m; i++) {
    double alpha = x[i];
    for (int k = ii[i]; k < ii[i + 1]; k++) y[aj[k]] = fma(alpha, aa[k], y[aj[k]]);
  }
}

/* The level smoother the reference's options select (configs/PETSc_SolverOptions_GAMG.info:15-21:
 * richardson, max_it 1, bjacobi/jacobi) as PETSc runs it [P376]: r = b - A x (KSP_MatMult then
 * VecAYPX(r,-1,b)), z = dinv .* r (VecPointwiseMult), x = x + z (VecAXPY, scale 1).  Separate
 * calls, separately rounded. */
void orc_residual(int m, const int *ii, const int *aj, const double *aa, const double *x,
                  const double *b, double *r)
{
  orc_matmult(m, ii, aj, aa, x, r);
  for (int i = 0; i < m; i++) r[i] = b[i] + (-1.0) * r[i];
}
void orc_jacobi_sweep(int m, const int *ii, const int *aj, const double *aa, const double *x,
                      const double *b, const double *dinv, double *xnew)
{
  orc_residual(m, ii, aj, aa, x, b, xnew);
  for (int i = 0; i < m; i++) {
    double z = dinv[i] * xnew[i];
    xnew[i]  = x[i] + z;
  }
}

/* PetscLogFlops(2.0*a->nz - a->nonzerorowcnt), src/openacc-step2/MatMult_SeqAIJ.patch:47 */
double orc_matmult_flops(int nz, int nonzerorowcnt) { return 2.0 * nz - nonzerorowcnt; }

/* Per-row bound used by the fast-kernel tolerance: s[i] = sum_k |aa[k]*x[aj[k]]| */
void orc_row_abs_sum(int m, const int *ii, const int *aj, const double *aa, const double *x,
                     double *s)
{
  for (int i = 0; i < m; i++) {
    double t = 0.0;
    for (int k = ii[i]; k < ii[i + 1]; k++) t += fabs(aa[k] * x[aj[k]]);
    s[i] = t;
  }
}

/* ------------------------------------------------------------------------------------------- */
/* CPU baseline: the same kernel with one contiguous row block per thread, mimicking "one MPI   */
/* rank per core" (runs/single-node-scaling.pbs:56-64).  Row blocks are balanced by nnz.        */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
  int           r0, r1, fma;
  const int    *ii, *aj;
  const double *aa, *x;
  double       *y;
} mt_job;

static void *mt_worker(void *p)
{
  mt_job *j = (mt_job *)p;
  int     m = j->r1 - j->r0;
  /* shift so that the block looks like its own SeqAIJ matrix with global column ids */
  if (j->fma) orc_matmult_fma(m, j->ii + j->r0, j->aj, j->aa, j->x, j->y + j->r0);
  else orc_matmult(m, j->ii + j->r0, j->aj, j->aa, j->x, j->y + j->r0);
  return NULL;
}

void orc_matmult_mt(int nthreads, int use_fma, int m, const int *ii, const int *aj,
                    const double *aa, const double *x, double *y)
{
  if (nthreads < 1) nthreads = 1;
  pthread_t *th  = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
  mt_job    *job = (mt_job *)malloc(sizeof(mt_job) * (size_t)nthreads);
  long long  nz  = ii[m];
  int        r   = 0;
  for (int t = 0; t < nthreads; t++) {
    long long target = nz * (t + 1) / nthreads;
    int       r1     = r;
    if (t == nthreads - 1) r1 = m;
    else {
      /* first row whose end offset reaches the target (binary search on ii) */
      int lo = r, hi = m;
      while (lo < hi) {
        int mid = lo + (hi - lo) / 2;
        if (ii[mid] < target) lo = mid + 1; else hi = mid;
      }
      r1 = lo;
    }
    job[t] = (mt_job){r, r1, use_fma, ii, aj, aa, x, y};
    r      = r1;
  }
  for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, mt_worker, &job[t]);
  mt_worker(&job[0]);
  for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th);
  free(job);
}

/* ------------------------------------------------------------------------------------------- */
/* A5: MatAssemblyEnd_SeqAIJ -- the compaction loop.                                             */
/* Context lines src/openacc-step2/MatAssemblyEnd_SeqAIJ.patch:34-37 show its head ("if (m) rmax */
/* = ailen[0]", "for (i=1; i<m; i++)", "move each row back by the amount of empty slots          */
/* (fshift) before it"); the rest is [P376].  Rows are stored with imax[i] reserved slots of     */
/* which ailen[i] are used; after the call ai is the packed CSR row-pointer array.               */
/* Returns fshift (the number of unneeded slots).                                                */
/* ------------------------------------------------------------------------------------------- */
int orc_assembly_end(int m, int *ai, int *aj, double *aa, int *imax, int *ailen, int *nz_out,
                     int *nonzerorowcnt_out, int *rmax_out)
{
  int fshift = 0, rmax = 0;
  if (m) rmax = ailen[0];
  for (int i = 1; i < m; i++) {
    fshift += imax[i - 1] - ailen[i - 1];
    if (ailen[i] > rmax) rmax = ailen[i];
    if (fshift) {
      int    *ip = aj + ai[i];
      double *ap = aa + ai[i];
      int     N  = ailen[i];
      for (int j = 0; j < N; j++) {
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
  int nzr = 0;
  for (int i = 0; i < m; i++) {
    ailen[i] = imax[i] = ai[i + 1] - ai[i];
    nzr += ((ai[i + 1] - ai[i]) > 0);
  }
  *nz_out            = m ? ai[m] : 0;
  *nonzerorowcnt_out = nzr;
  *rmax_out          = rmax;
  return fshift;
}

/* MatCheckCompressedRow [P376] (called at src/openacc-step2/MatAssemblyEnd_SeqAIJ.patch:77 with
 * ratio = 0.6): use the compressed-row structure when the zero-row count is >= ratio*m.
 * cpi must hold nonzerorowcnt+1 ints, ridx nonzerorowcnt ints.  Returns use (0/1). */
int orc_check_compressed_row(int m, const int *ai, int nonzerorowcnt, double ratio, int *cpi,
                             int *ridx, int *nrows_out)
{
  int nzero = m - nonzerorowcnt;
  *nrows_out = 0;
  if (nzero < ratio * m) return 0;
  int row = 0;
  cpi[0]  = 0;
  for (int i = 0; i < m; i++) {
    int nz = ai[i + 1] - ai[i];
    if (nz == 0) continue;
    cpi[row + 1] = ai[i + 1];
    ridx[row++]  = i;
  }
  *nrows_out = row;
  return 1;
}

/* ------------------------------------------------------------------------------------------- */
/* A9: the problem generator, src/helper.cpp.                                                    */
/* DMDA bits are [P376] (un-vendored): PETSC_DECIDE process grid, ownership split, PETSc         */
/* ordering (rank-major, then i fastest / j / k inside a rank's sub-box).                        */
/* ------------------------------------------------------------------------------------------- */

/* DMSetUp_DA_3D's PETSC_DECIDE heuristic [P376], called from src/helper.cpp:31-36. */
void orc_dmda_decide(int M, int N, int P, int size, int *m_out, int *n_out, int *p_out)
{
  int m, n, p, pm;
  n = (int)(0.5 + pow(((double)N * N) * ((double)size) / ((double)P * M), 1. / 3.));
  if (!n) n = 1;
  while (n > 0) {
    pm = size / n;
    if (n * pm == size) break;
    n--;
  }
  if (!n) n = 1;
  m = (int)(0.5 + sqrt(((double)M) * ((double)size) / ((double)P * n)));
  if (!m) m = 1;
  p = 1;
  while (m > 0) {
    p = size / (m * n);
    if (m * n * p == size) break;
    m--;
  }
  if (M > P && m < p) { int t = m; m = p; p = t; }
  *m_out = m; *n_out = n; *p_out = p;
}

/* ownership along one dimension: process q of np gets M/np + ((M % np) > q) cells [P376] */
static void own_range(int M, int np, int q, int *start, int *len)
{
  int s = 0;
  for (int t = 0; t < q; t++) s += M / np + ((M % np) > t);
  *start = s;
  *len   = M / np + ((M % np) > q);
}

/* out[9] = m n p xs ys zs xm ym zm for this rank (rank = pi + pj*m + pk*m*n [P376]) */
void orc_dmda_info(int M, int N, int P, int size, int rank, int *out)
{
  int m, n, p;
  orc_dmda_decide(M, N, P, size, &m, &n, &p);
  int pi = rank % m, pj = (rank / m) % n, pk = rank / (m * n);
  out[0] = m; out[1] = n; out[2] = p;
  own_range(M, m, pi, &out[3], &out[6]);
  own_range(N, n, pj, &out[4], &out[7]);
  own_range(P, p, pk, &out[5], &out[8]);
}

/* base[r] = first global (PETSc-ordering) row of rank r; base[size] = M*N*P */
void orc_dmda_bases(int M, int N, int P, int size, int *base)
{
  int info[9];
  base[0] = 0;
  for (int r = 0; r < size; r++) {
    orc_dmda_info(M, N, P, size, r, info);
    base[r + 1] = base[r] + info[6] * info[7] * info[8];
  }
}

typedef struct {
  int M, N, P, m, n, p, size;
  int *xs, *xl, *ys, *yl, *zs, *zl; /* per process-coordinate start / length */
  int *base;
} dmda_t;

static void dmda_init(dmda_t *d, int M, int N, int P, int size)
{
  d->M = M; d->N = N; d->P = P; d->size = size;
  orc_dmda_decide(M, N, P, size, &d->m, &d->n, &d->p);
  d->xs = malloc(sizeof(int) * d->m); d->xl = malloc(sizeof(int) * d->m);
  d->ys = malloc(sizeof(int) * d->n); d->yl = malloc(sizeof(int) * d->n);
  d->zs = malloc(sizeof(int) * d->p); d->zl = malloc(sizeof(int) * d->p);
  for (int q = 0; q < d->m; q++) own_range(M, d->m, q, &d->xs[q], &d->xl[q]);
  for (int q = 0; q < d->n; q++) own_range(N, d->n, q, &d->ys[q], &d->yl[q]);
  for (int q = 0; q < d->p; q++) own_range(P, d->p, q, &d->zs[q], &d->zl[q]);
  d->base = malloc(sizeof(int) * (size + 1));
  orc_dmda_bases(M, N, P, size, d->base);
}
static void dmda_free(dmda_t *d)
{
  free(d->xs); free(d->xl); free(d->ys); free(d->yl); free(d->zs); free(d->zl); free(d->base);
}
static int find_owner(const int *s, const int *l, int np, int c)
{
  for (int q = 0; q < np; q++) if (c >= s[q] && c < s[q] + l[q]) return q;
  return -1;
}
/* DMDAConvertToCell + ISLocalToGlobalMappingApply (src/helper.cpp:217-226): the global PETSc
 * index of cell (i,j,k), or -1 for a ghost cell outside the domain (DM_BOUNDARY_GHOSTED). */
static int cell_global(const dmda_t *d, int i, int j, int k)
{
  if (i < 0 || j < 0 || k < 0 || i >= d->M || j >= d->N || k >= d->P) return -1;
  int pi = find_owner(d->xs, d->xl, d->m, i);
  int pj = find_owner(d->ys, d->yl, d->n, j);
  int pk = find_owner(d->zs, d->zl, d->p, k);
  int r  = pi + pj * d->m + pk * d->m * d->n;
  return d->base[r] + (i - d->xs[pi]) + (j - d->ys[pj]) * d->xl[pi] +
         (k - d->zs[pk]) * d->xl[pi] * d->yl[pj];
}

/* c1 and c2 exactly as the (unparenthesised) macros of src/helper.cpp:14-18 expand */
#define ORC_C1 2.0 * 1.0 * M_PI
#define ORC_C2 -3.0 * ORC_C1 * ORC_C1

/* generateRHS (src/helper.cpp:78-116) and generateExt (:120-157) for this rank's sub-box. */
void orc_poisson7_vectors(int M, int N, int P, int size, int rank, double *rhs, double *exact)
{
  int info[9];
  orc_dmda_info(M, N, P, size, rank, info);
  int xs = info[3], ys = info[4], zs = info[5], xm = info[6], ym = info[7], zm = info[8];
  double dx = 1.0 / M, dy = 1.0 / N, dz = 1.0 / P;
  size_t r = 0;
  for (int k = zs; k < zs + zm; ++k)
    for (int j = ys; j < ys + ym; ++j)
      for (int i = xs; i < xs + xm; ++i, ++r) {
        double cx = cos(ORC_C1 * (i + 0.5) * dx);
        double cy = cos(ORC_C1 * (j + 0.5) * dy);
        double cz = cos(ORC_C1 * (k + 0.5) * dz);
        if (rhs) rhs[r] = ORC_C2 * cx * cy * cz;
        if (exact) exact[r] = cx * cy * cz;
      }
}

/* generateA (src/helper.cpp:161-246): this rank's rows, global PETSc column ids, columns
 * ascending inside a row (MatSetValues keeps rows sorted [P376]); absent neighbours (col -1)
 * are dropped (:233 and MatSetValues ignoring negative columns [P376]).
 * ai[nloc+1], aj/aa[7*nloc].  Returns nnz. */
int orc_poisson7_rows(int M, int N, int P, int size, int rank, int *ai, int *aj, double *aa)
{
  dmda_t d;
  dmda_init(&d, M, N, P, size);
  int info[9];
  orc_dmda_info(M, N, P, size, rank, info);
  int xs = info[3], ys = info[4], zs = info[5], xm = info[6], ym = info[7], zm = info[8];
  double dx = 1.0 / M, dy = 1.0 / N, dz = 1.0 / P;
  double values[7];
  int    cols[7];
  values[1] = values[2] = 1.0 / (dx * dx);
  values[3] = values[4] = 1.0 / (dy * dy);
  values[5] = values[6] = 1.0 / (dz * dz);
  int row = 0, nz = 0;
  ai[0] = 0;
  for (int k = zs; k < zs + zm; ++k)
    for (int j = ys; j < ys + ym; ++j)
      for (int i = xs; i < xs + xm; ++i) {
        cols[0] = cell_global(&d, i, j, k);
        cols[1] = cell_global(&d, i - 1, j, k);
        cols[2] = cell_global(&d, i + 1, j, k);
        cols[3] = cell_global(&d, i, j - 1, k);
        cols[4] = cell_global(&d, i, j + 1, k);
        cols[5] = cell_global(&d, i, j, k - 1);
        cols[6] = cell_global(&d, i, j, k + 1);
        values[0] = 0.0;
        for (int idx = 1; idx < 7; ++idx)
          if (cols[idx] > -1) values[0] -= values[idx];
        /* sorted insertion of the present columns */
        int cnt = 0;
        for (int idx = 0; idx < 7; ++idx) {
          if (cols[idx] < 0) continue;
          int pos = cnt;
          while (pos > 0 && aj[nz + pos - 1] > cols[idx]) {
            aj[nz + pos] = aj[nz + pos - 1];
            aa[nz + pos] = aa[nz + pos - 1];
            pos--;
          }
          aj[nz + pos] = cols[idx];
          aa[nz + pos] = values[idx];
          cnt++;
        }
        nz += cnt;
        ai[++row] = nz;
      }
  dmda_free(&d);
  return nz;
}

/* setRefPoint's scale (src/helper.cpp:262-270): MatGetDiagonal, VecSum, divide by the global
 * size.  VecSum [P376] = sequential local sum, then MPI_Allreduce(SUM); the all-reduce order is
 * not defined by MPI, here the rank partial sums are added in rank order. */
double orc_poisson7_diag_scale(int M, int N, int P, int size)
{
  double dx = 1.0 / M, dy = 1.0 / N, dz = 1.0 / P;
  double v[7];
  v[1] = v[2] = 1.0 / (dx * dx);
  v[3] = v[4] = 1.0 / (dy * dy);
  v[5] = v[6] = 1.0 / (dz * dz);
  double total = 0.0;
  for (int r = 0; r < size; r++) {
    int info[9];
    orc_dmda_info(M, N, P, size, r, info);
    int xs = info[3], ys = info[4], zs = info[5], xm = info[6], ym = info[7], zm = info[8];
    double lsum = 0.0;
    for (int k = zs; k < zs + zm; ++k)
      for (int j = ys; j < ys + ym; ++j)
        for (int i = xs; i < xs + xm; ++i) {
          int present[7] = {1, i > 0, i < M - 1, j > 0, j < N - 1, k > 0, k < P - 1};
          double dg = 0.0;
          for (int idx = 1; idx < 7; ++idx) if (present[idx]) dg -= v[idx];
          lsum += dg;
        }
    total += lsum;
  }
  return total / (double)((long long)M * N * P);
}

/* MatZeroRowsColumns(A, 1, {0}, scale, exact, rhs) (src/helper.cpp:272) on this rank's rows
 * [P376]: the pattern is kept; row 0's values become 0 with `scale` on the diagonal and
 * rhs[0] = scale*exact[0]; every other row with an entry in column 0 gets
 * rhs[i] -= a_i0*exact[0] and a_i0 = 0.  rstart = this rank's first global row. */
void orc_poisson7_refpoint(int nloc, int rstart, const int *ai, const int *aj, double *aa,
                           double *rhs, double exact0, double scale)
{
  for (int i = 0; i < nloc; i++) {
    int grow = rstart + i;
    if (grow == 0) {
      for (int k = ai[i]; k < ai[i + 1]; k++) aa[k] = 0.0;
      rhs[i] = scale * exact0;
    } else {
      for (int k = ai[i]; k < ai[i + 1]; k++)
        if (aj[k] == 0) {
          rhs[i] -= aa[k] * exact0;
          aa[k] = 0.0;
        }
    }
  }
  if (rstart == 0 && nloc > 0)
    for (int k = ai[0]; k < ai[1]; k++) if (aj[k] == 0) aa[k] = scale;
}

/* exact[0] of the global vector = cell (0,0,0) */
double orc_poisson7_exact0(int M, int N, int P)
{
  double dx = 1.0 / M, dy = 1.0 / N, dz = 1.0 / P;
  return cos(ORC_C1 * (0 + 0.5) * dx) * cos(ORC_C1 * (0 + 0.5) * dy) * cos(ORC_C1 * (0 + 0.5) * dz);
}

/* ------------------------------------------------------------------------------------------- */
/* A10: Mat_MPIAIJ {A, B, garray} and the VecScatter lists [P376].                               */
/* ------------------------------------------------------------------------------------------- */

/* MatSetValues_MPIAIJ's split: columns in [cstart,cend) go to the diagonal block A with local
 * ids, the rest to the off-diagonal block B (global ids until MatSetUpMultiply).
 * Output arrays sized like the input.  Returns nnz(A); *bnz_out = nnz(B). */
int orc_mpiaij_split(int nloc, int cstart, int cend, const int *ai, const int *aj,
                     const double *aa, int *Ai, int *Aj, double *Aa, int *Bi, int *Bj, double *Ba,
                     int *bnz_out)
{
  int an = 0, bn = 0;
  Ai[0] = Bi[0] = 0;
  for (int i = 0; i < nloc; i++) {
    for (int k = ai[i]; k < ai[i + 1]; k++) {
      if (aj[k] >= cstart && aj[k] < cend) { Aj[an] = aj[k] - cstart; Aa[an++] = aa[k]; }
      else { Bj[bn] = aj[k]; Ba[bn++] = aa[k]; }
    }
    Ai[i + 1] = an;
    Bi[i + 1] = bn;
  }
  *bnz_out = bn;
  return an;
}

static int cmp_int(const void *a, const void *b)
{
  int x = *(const int *)a, y = *(const int *)b;
  return (x > y) - (x < y);
}

/* MatSetUpMultiply_MPIAIJ: garray = sorted unique global column ids of B; B's column ids are
 * rewritten in place to positions in garray.  garray must hold bnz ints.  Returns nghost. */
int orc_mpiaij_setup_multiply(int bnz, int *Bj, int *garray)
{
  if (bnz == 0) return 0;
  int *tmp = (int *)malloc(sizeof(int) * (size_t)bnz);
  memcpy(tmp, Bj, sizeof(int) * (size_t)bnz);
  qsort(tmp, (size_t)bnz, sizeof(int), cmp_int);
  int ng = 0;
  for (int k = 0; k < bnz; k++) if (k == 0 || tmp[k] != tmp[k - 1]) garray[ng++] = tmp[k];
  free(tmp);
  for (int k = 0; k < bnz; k++) {
    int lo = 0, hi = ng - 1;
    while (lo < hi) {
      int mid = (lo + hi) / 2;
      if (garray[mid] < Bj[k]) lo = mid + 1; else hi = mid;
    }
    Bj[k] = lo;
  }
  return ng;
}

/* VecScatter lists for Mvctx: rank `me` needs garray[] (sorted).  For each owner rank q the
 * receive segment is the contiguous run of garray inside [base[q], base[q+1]); the matching send
 * list on q is those ids minus base[q], in the same (ascending) order.
 * recv_off[size+1]: lvec offsets per owner.  Returns 0. */
int orc_scatter_recv_offsets(int size, const int *base, int ng, const int *garray, int *recv_off)
{
  int k = 0;
  for (int q = 0; q < size; q++) {
    recv_off[q] = k;
    while (k < ng && garray[k] < base[q + 1]) k++;
  }
  recv_off[size] = k;
  return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* A11: CG vector ops [P376] (sequential BLAS-1 semantics) and a reference CG used to pin the    */
/* PETSc-free solver driver: KSPCG, left preconditioning, preconditioned-residual norm,          */
/* diagonal (Jacobi) preconditioner.                                                             */
/* ------------------------------------------------------------------------------------------- */
double orc_vecdot(int n, const double *x, const double *y)
{
  double s = 0.0;
  for (int i = 0; i < n; i++) s += x[i] * y[i];
  return s;
}
double orc_vecnorm2(int n, const double *x) { return sqrt(orc_vecdot(n, x, x)); }
double orc_vecnorm_inf(int n, const double *x)
{
  double s = 0.0;
  for (int i = 0; i < n; i++) { double a = fabs(x[i]); if (a > s) s = a; }
  return s;
}
/* y = y + a x */
void orc_vecaxpy(int n, double a, const double *x, double *y) { for (int i = 0; i < n; i++) y[i] += a * x[i]; }
/* y = x + a y */
void orc_vecaypx(int n, double a, const double *x, double *y) { for (int i = 0; i < n; i++) y[i] = x[i] + a * y[i]; }

/* KSPSolve_CG [P376] with PCJACOBI, KSP_NORM_PRECONDITIONED (the CG default), zero initial
 * guess, convergence test rnorm < max(rtol*rnorm0, atol) (KSPConvergedDefault).
 * Returns the iteration count (negative when max_it was hit); *rnorm_out = last norm. */
int orc_cg_jacobi(int m, const int *ii, const int *aj, const double *aa, const double *b,
                  double *x, double rtol, double atol, int max_it, double *rnorm_out)
{
  double *r = malloc(sizeof(double) * m), *z = malloc(sizeof(double) * m);
  double *p = malloc(sizeof(double) * m), *w = malloc(sizeof(double) * m);
  double *dinv = malloc(sizeof(double) * m);
  for (int i = 0; i < m; i++) {
    double d = 0.0;
    for (int k = ii[i]; k < ii[i + 1]; k++) if (aj[k] == i) d = aa[k];
    dinv[i] = (d != 0.0) ? 1.0 / d : 1.0;
  }
  memset(x, 0, sizeof(double) * m);
  memcpy(r, b, sizeof(double) * m);
  for (int i = 0; i < m; i++) z[i] = dinv[i] * r[i];
  double dp = orc_vecnorm2(m, z), rnorm0 = dp, beta = 0.0, betaold = 1.0;
  double ttol = rtol * rnorm0 > atol ? rtol * rnorm0 : atol;
  int    it = 0, conv = (dp < ttol);
  beta = orc_vecdot(m, z, r);
  while (!conv && it < max_it) {
    if (it == 0) memcpy(p, z, sizeof(double) * m);
    else orc_vecaypx(m, beta / betaold, z, p);
    betaold = beta;
    orc_matmult(m, ii, aj, aa, p, w);
    double dpi = orc_vecdot(m, p, w);
    double a   = beta / dpi;
    orc_vecaxpy(m, a, p, x);
    orc_vecaxpy(m, -a, w, r);
    for (int i = 0; i < m; i++) z[i] = dinv[i] * r[i];
    dp = orc_vecnorm2(m, z);
    it++;
    if (dp < ttol) { conv = 1; break; }
    beta = orc_vecdot(m, z, r);
  }
  *rnorm_out = dp;
  free(r); free(z); free(p); free(w); free(dinv);
  return conv ? it : -it;
}

/* ------------------------------------------------------------------------------------------- */
/* The multigrid preconditioner the reference's options select                                  */
/* (configs/PETSc_SolverOptions_GAMG.info:6-20): PCMG multiplicative V-cycle [P376,             */
/* PCMGMCycle_Private] over a given hierarchy -- level l has the operator A_l (m[l] x m[l]) and, */
/* except on the coarsest, the prolongator P_l (m[l] x m[l+1]).  Smoother: richardson(sweeps) +  */
/* jacobi before and after; coarse "solve": one Jacobi application (preonly + jacobi).           */
/* Each step is the separate PETSc call, separately rounded:                                    */
/*   first pre-smoothing step from the zero guess   x = dinv .* b      (VecPointwiseMult)        */
/*   residual                                       r = b - A x        (MatMult, VecAYPX)        */
/*   restriction                                    b_c = P^T r        (MatMultTranspose)        */
/*   interpolation                                  x = x + P x_c      (MatMultAdd)              */
/*   smoothing                                      orc_jacobi_sweep                             */
/* PCGAMG itself (the hierarchy) has no text in the reference; oracle/gamg.py restates it.       */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
  int            nlev, sweeps;
  const int     *m;
  int *const    *ai, *const *aj, *const *pi, *const *pj;
  double *const *aa, *const *pa;
  double       **dinv;
} orc_mg_t;

static void orc_mg_cycle(const orc_mg_t *g, int l, const double *b, double *x)
{
  const int     m = g->m[l];
  const double *dinv = g->dinv[l];
  if (l == g->nlev - 1) {
    for (int i = 0; i < m; i++) x[i] = b[i] * dinv[i];
    return;
  }
  const int mc = g->m[l + 1];
  double   *t = malloc(sizeof(double) * (m > 0 ? m : 1)), *r = malloc(sizeof(double) * (m > 0 ? m : 1));
  double   *bc = malloc(sizeof(double) * (mc > 0 ? mc : 1)), *xc = malloc(sizeof(double) * (mc > 0 ? mc : 1));
  for (int i = 0; i < m; i++) x[i] = b[i] * dinv[i];
  for (int s = 1; s < g->sweeps; s++) {
    orc_jacobi_sweep(m, g->ai[l], g->aj[l], g->aa[l], x, b, dinv, t);
    memcpy(x, t, sizeof(double) * m);
  }
  orc_residual(m, g->ai[l], g->aj[l], g->aa[l], x, b, r);
  orc_matmulttranspose(m, mc, g->pi[l], g->pj[l], g->pa[l], r, bc);
  orc_mg_cycle(g, l + 1, bc, xc);
  orc_matmultadd(m, g->pi[l], g->pj[l], g->pa[l], xc, x, x);
  for (int s = 0; s < g->sweeps; s++) {
    orc_jacobi_sweep(m, g->ai[l], g->aj[l], g->aa[l], x, b, dinv, t);
    memcpy(x, t, sizeof(double) * m);
  }
  free(t); free(r); free(bc); free(xc);
}

static void orc_mg_init(orc_mg_t *g, int nlev, int sweeps, const int *m, int *const *ai, int *const *aj,
                        double *const *aa, int *const *pi, int *const *pj, double *const *pa)
{
  g->nlev = nlev; g->sweeps = sweeps; g->m = m;
  g->ai = ai; g->aj = aj; g->aa = aa; g->pi = pi; g->pj = pj; g->pa = pa;
  g->dinv = malloc(sizeof(double *) * nlev);
  for (int l = 0; l < nlev; l++) {  /* PCSetUp_Jacobi [P376]: 1/diagonal, zero -> 1 */
    g->dinv[l] = malloc(sizeof(double) * (m[l] > 0 ? m[l] : 1));
    for (int i = 0; i < m[l]; i++) {
      double d = 0.0;
      for (int k = ai[l][i]; k < ai[l][i + 1]; k++) if (aj[l][k] == i) { d = aa[l][k]; break; }
      g->dinv[l][i] = (d != 0.0) ? 1.0 / d : 1.0;
    }
  }
}
static void orc_mg_free(orc_mg_t *g)
{
  for (int l = 0; l < g->nlev; l++) free(g->dinv[l]);
  free(g->dinv);
}

/* z = M^{-1} r: one V-cycle */
void orc_mg_apply(int nlev, int sweeps, const int *m, int *const *ai, int *const *aj, double *const *aa,
                  int *const *pi, int *const *pj, double *const *pa, const double *r, double *z)
{
  orc_mg_t g;
  orc_mg_init(&g, nlev, sweeps, m, ai, aj, aa, pi, pj, pa);
  orc_mg_cycle(&g, 0, r, z);
  orc_mg_free(&g);
}

/* KSPSolve_CG [P376] preconditioned by the V-cycle; same conventions as orc_cg_jacobi. */
int orc_cg_mg(int nlev, int sweeps, const int *m, int *const *ai, int *const *aj, double *const *aa,
              int *const *pi, int *const *pj, double *const *pa, const double *b, double *x,
              double rtol, double atol, int max_it, double *rnorm_out)
{
  orc_mg_t g;
  orc_mg_init(&g, nlev, sweeps, m, ai, aj, aa, pi, pj, pa);
  const int n = m[0];
  double *r = malloc(sizeof(double) * n), *z = malloc(sizeof(double) * n);
  double *p = malloc(sizeof(double) * n), *w = malloc(sizeof(double) * n);
  memset(x, 0, sizeof(double) * n);
  memcpy(r, b, sizeof(double) * n);
  orc_mg_cycle(&g, 0, r, z);
  double dp = orc_vecnorm2(n, z), rnorm0 = dp, beta = 0.0, betaold = 1.0;
  double ttol = rtol * rnorm0 > atol ? rtol * rnorm0 : atol;
  int    it = 0, conv = (dp < ttol);
  beta = orc_vecdot(n, z, r);
  while (!conv && it < max_it) {
    if (it == 0) memcpy(p, z, sizeof(double) * n);
    else orc_vecaypx(n, beta / betaold, z, p);
    betaold = beta;
    orc_matmult(n, ai[0], aj[0], aa[0], p, w);
    double dpi = orc_vecdot(n, p, w);
    double a   = beta / dpi;
    orc_vecaxpy(n, a, p, x);
    orc_vecaxpy(n, -a, w, r);
    orc_mg_cycle(&g, 0, r, z);
    dp = orc_vecnorm2(n, z);
    it++;
    if (dp < ttol) { conv = 1; break; }
    beta = orc_vecdot(n, z, r);
  }
  *rnorm_out = dp;
  free(r); free(z); free(p); free(w);
  orc_mg_free(&g);
  return conv ? it : -it;
}
