// pcgamg.cpp -- `-pc_type gamg` for the reference's solver options
// (configs/PETSc_SolverOptions_GAMG.info): smoothed-aggregation algebraic multigrid whose every
// level operation in the solve runs through the SeqAIJ hot path on the device.
//
// The reference takes PCGAMG from PETSc 3.7.6, which is not under /root/reference
// (scripts/petsc.sh:38-40 downloads it), so there is no text to follow; this file restates the
// published structure of PCGAMG "agg" + PCMG [P376]:
//
//   setup (host, like PETSc's -- it is integer and sparse-product work done once per solve)
//     graph       |a_ij| / sqrt(|a_ii a_jj|) > -pc_gamg_threshold, diagonal and zeros dropped
//                 (PCGAMGGraph_AGG + PCGAMGFilterGraph)
//     coarsen     greedy maximal independent set; on the first -pc_gamg_square_graph levels (1)
//                 of the SQUARED graph, aggregate = root + the vertices it removes
//                 (PCGAMGCoarsen_AGG / MatCoarsen MIS with strict aggregates); vertices without a
//                 strong neighbour stay out of every aggregate (empty prolongator row)
//     prolongator tentative P0 from the near-null-space vector B (constants on the fine grid,
//                 carried down as the R factor of the per-aggregate QR = its norm), then one
//                 damped-Jacobi smoothing step P = (I - 1.4/emax D^-1 A) P0
//                 (-pc_gamg_agg_nsmooths 1, PCGAMGOptProlongator_AGG), emax from 10 iterations of
//                 Jacobi-CG on the device
//     coarse op   Galerkin A_c = P^T A P (MatPtAP); stop at -pc_gamg_coarse_eq_limit (50) rows
//   solve (device): PCMG multiplicative V-cycle, richardson(k)+jacobi pre/post smoothing,
//     one Jacobi application as the coarse "solve" -- exactly what the options file asks for:
//       residual    b - A x            b200_spmv_residual (MatMult + VecAYPX in one pass)
//       restrict    P^T r              MatMultTranspose_SeqAIJ
//       interpolate x + P x_c          MatMultAdd_SeqAIJ
//       smooth      x + D^-1 (b - A x) b200_spmv_jacobi_sweep (MatMult+VecAYPX+PointwiseMult+AXPY)
//
// Deliberate differences from PETSc, none of which can be pinned without its sources: the MIS
// visits vertices in natural order (PETSc permutes them with PetscRandom), PETSc's smoothAggs
// clean-up of squared-graph aggregates is not applied, the right-hand side of the eigenvalue
// estimate is splitmix64 noise, and the first pre-smoothing step starts from the known zero guess
// without multiplying by A.  Iteration counts are therefore this implementation's own (oracle:
// oracle/gamg.py + orc_mg_*), NOT PETSc's -- see DESIGN.md "parity unpinned".
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "pc_impl.h"
// (after b200_aij.h: the symbols header only forward-declares Mat/Vec when PETSc types are absent)
#include "../../../include/b200_petsc_symbols.h"
#include "../../../include/b200_seqaij.h"

namespace {

// ---------------------------------------------------------------------------------------------
// host sparse kernels of the setup
// ---------------------------------------------------------------------------------------------
struct CsrView {
  PetscInt         m = 0, n = 0;
  const PetscInt  *i = nullptr, *j = nullptr;
  const MatScalar *a = nullptr;
};
struct Csr {
  PetscInt               m = 0, n = 0;
  std::vector<PetscInt>  i, j;
  std::vector<MatScalar> a;
  CsrView view() const { return {m, n, i.data(), j.data(), a.data()}; }
};

struct Stopwatch {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double lap()
  {
    auto   t1 = std::chrono::steady_clock::now();
    double s  = std::chrono::duration<double>(t1 - t0).count();
    t0 = t1;
    return s;
  }
};

int setup_threads()
{
  const char *s = getenv("B200_SETUP_THREADS");
  int         t = s ? atoi(s) : (int)std::thread::hardware_concurrency();
  return std::max(1, std::min(t, 64));
}

template <class F>
void parallel_blocks(PetscInt m, F f)
{
  const int nt = (m < 20000) ? 1 : setup_threads();
  if (nt == 1) { f(0, 0, m); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t) {
    const PetscInt r0 = (PetscInt)((long long)m * t / nt), r1 = (PetscInt)((long long)m * (t + 1) / nt);
    th.emplace_back(f, t, r0, r1);
  }
  for (auto &x : th) x.join();
}

// Rows of a CSR built independently: `row(t, r, j, a)` appends the entries of row r (ascending
// columns) to the calling thread's buffers.  Row blocks run on the set-up threads; the blocks are
// then laid end to end, so the result does not depend on the thread count.
template <class RowFn, class BoundFn>
PetscErrorCode build_rows(PetscInt m, PetscInt n, bool values, Csr &C, RowFn row, BoundFn bound)
{
  const int nt = (m < 20000) ? 1 : setup_threads();
  struct Part { std::vector<PetscInt> len, j; std::vector<MatScalar> a; };
  std::vector<Part> parts(nt);
  parallel_blocks(m, [&](int t, PetscInt r0, PetscInt r1) {
    Part &p = parts[t];
    p.len.resize((size_t)(r1 - r0));
    // room for the most this block can produce: untouched pages cost nothing, regrowing does
    const size_t most = bound(r0, r1);
    p.j.reserve(most);
    if (values) p.a.reserve(most);
    for (PetscInt r = r0; r < r1; ++r) {
      const size_t before = p.j.size();
      row(t, r, p.j, p.a);
      p.len[r - r0] = (PetscInt)(p.j.size() - before);
    }
  });
  std::vector<long long> first((size_t)nt + 1, 0);
  for (int t = 0; t < nt; ++t) first[t + 1] = first[t] + (long long)parts[t].j.size();
  if (first[nt] > 2147483647LL) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "sparse product exceeds int32 indices");
  C.m = m; C.n = n;
  C.i.assign((size_t)m + 1, 0);
  C.j.resize((size_t)first[nt]);
  C.a.resize(values ? (size_t)first[nt] : 0);
  parallel_blocks(m, [&](int t, PetscInt r0, PetscInt r1) {
    Part    &p   = parts[t];
    PetscInt pos = (PetscInt)first[t];
    for (PetscInt r = r0; r < r1; ++r) { C.i[r] = pos; pos += p.len[r - r0]; }
    std::copy(p.j.begin(), p.j.end(), C.j.begin() + first[t]);
    if (values) std::copy(p.a.begin(), p.a.end(), C.a.begin() + first[t]);
    p = Part();
  });
  C.i[m] = (PetscInt)first[nt];
  return 0;
}
template <class RowFn>
PetscErrorCode build_rows(PetscInt m, PetscInt n, bool values, Csr &C, RowFn row)
{
  return build_rows(m, n, values, C, row, [](PetscInt, PetscInt) { return (size_t)0; });
}

// C = X * Y, row by row with a dense accumulator; every entry is summed over X's row in storage
// order (then over Y's row).
PetscErrorCode spgemm(const CsrView &X, const CsrView &Y, Csr &C)
{
  if (X.n != Y.m) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_SIZ, "spgemm: inner dimensions differ");
  struct Scratch { std::vector<PetscInt> mark, cols; std::vector<MatScalar> acc; };
  std::vector<Scratch> scratch((size_t)setup_threads());
  return build_rows(X.m, Y.n, true, C, [&](int t, PetscInt r, std::vector<PetscInt> &cj, std::vector<MatScalar> &ca) {
    Scratch &w = scratch[t];
    if (w.mark.empty()) { w.mark.assign((size_t)std::max(Y.n, 1), -1); w.acc.assign((size_t)std::max(Y.n, 1), 0.0); }
    w.cols.clear();
    for (PetscInt k = X.i[r]; k < X.i[r + 1]; ++k) {
      const MatScalar xv = X.a[k];
      const PetscInt  q  = X.j[k];
      for (PetscInt l = Y.i[q]; l < Y.i[q + 1]; ++l) {
        const PetscInt c = Y.j[l];
        if (w.mark[c] != r) { w.mark[c] = r; w.cols.push_back(c); w.acc[c] = xv * Y.a[l]; }
        else w.acc[c] += xv * Y.a[l];
      }
    }
    std::sort(w.cols.begin(), w.cols.end());
    for (PetscInt c : w.cols) { cj.push_back(c); ca.push_back(w.acc[c]); }
  }, [&](PetscInt r0, PetscInt r1) {
    size_t most = 0;   // every term of the block's rows lands in a column of its own
    for (PetscInt k = X.i[r0]; k < X.i[r1]; ++k) most += (size_t)(Y.i[X.j[k] + 1] - Y.i[X.j[k]]);
    return std::min(most, (size_t)(r1 - r0) * (size_t)std::max(Y.n, 1));
  });
}

// transpose: row c of the result lists column c's entries by ascending row.  Row blocks count
// their columns on the set-up threads; block t's entries of a column go after those of the blocks
// before it, so the order (and the result) is the sequential one.
void transpose(const CsrView &X, Csr &T)
{
  T.m = X.n; T.n = X.m;
  const PetscInt nz = X.i[X.m];
  T.i.assign((size_t)X.n + 1, 0);
  T.j.resize((size_t)nz); T.a.resize((size_t)nz);
  int nt = (X.m < 20000) ? 1 : setup_threads();
  while (nt > 1 && (size_t)nt * (size_t)X.n > (size_t)256 << 20) nt /= 2;   // cap the count tables at 1 GB
  std::vector<std::vector<PetscInt>> next((size_t)nt, std::vector<PetscInt>((size_t)X.n, 0));
  auto block = [&](int t) { return std::make_pair((PetscInt)((long long)X.m * t / nt), (PetscInt)((long long)X.m * (t + 1) / nt)); };
  auto each = [&](auto f) {
    if (nt == 1) { f(0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(f, t);
    for (auto &x : th) x.join();
  };
  each([&](int t) {
    const auto [r0, r1] = block(t);
    for (PetscInt k = X.i[r0]; k < X.i[r1]; ++k) next[t][X.j[k]]++;
  });
  PetscInt pos = 0;
  for (PetscInt c = 0; c < X.n; ++c) {   // next[t][c] <- where block t starts writing column c
    T.i[c] = pos;
    for (int t = 0; t < nt; ++t) { const PetscInt cnt = next[t][c]; next[t][c] = pos; pos += cnt; }
  }
  T.i[X.n] = pos;
  each([&](int t) {
    const auto [r0, r1] = block(t);
    for (PetscInt r = r0; r < r1; ++r)
      for (PetscInt k = X.i[r]; k < X.i[r + 1]; ++k) {
        const PetscInt p = next[t][X.j[k]]++;
        T.j[p] = r; T.a[p] = X.a[k];
      }
  });
}

void diagonal(const CsrView &A, std::vector<MatScalar> &d)
{
  d.assign((size_t)A.m, 0.0);
  for (PetscInt r = 0; r < A.m; ++r)
    for (PetscInt k = A.i[r]; k < A.i[r + 1]; ++k) if (A.j[k] == r) { d[r] = A.a[k]; break; }
}

// PCGAMGGraph_AGG + PCGAMGFilterGraph [P376]: pattern of the strong off-diagonal connections
PetscErrorCode strength_graph(const CsrView &A, const std::vector<MatScalar> &d, PetscReal threshold, Csr &G)
{
  std::vector<PetscReal> s((size_t)A.m);
  for (PetscInt r = 0; r < A.m; ++r) s[r] = (d[r] != 0.0) ? 1.0 / std::sqrt(std::fabs(d[r])) : 1.0;
  return build_rows(A.m, A.m, false, G, [&](int, PetscInt r, std::vector<PetscInt> &gj, std::vector<MatScalar> &) {
    for (PetscInt k = A.i[r]; k < A.i[r + 1]; ++k) {
      const PetscInt c = A.j[k];
      if (c == r || c >= A.m) continue;
      if (std::fabs(A.a[k]) * s[r] * s[c] > threshold) gj.push_back(c);
    }
  });
}

// greedy MIS in natural order on G (square = false) or on G*G (square = true, never formed);
// aggregate = root + the undecided vertices it removes.  agg[v] = -1: isolated vertex.
PetscInt aggregate(const Csr &G, bool square, std::vector<PetscInt> &agg)
{
  enum : char { UNDECIDED = 0, TAKEN = 1 };
  std::vector<char> state((size_t)G.m, UNDECIDED);
  agg.assign((size_t)G.m, -1);
  PetscInt nagg = 0;
  for (PetscInt v = 0; v < G.m; ++v) {
    if (state[v] != UNDECIDED) continue;
    state[v] = TAKEN;
    if (G.i[v + 1] == G.i[v]) continue;  // no strong neighbour: not aggregated
    const PetscInt id = nagg++;
    agg[v] = id;
    for (PetscInt k = G.i[v]; k < G.i[v + 1]; ++k) {
      const PetscInt w = G.j[k];
      if (state[w] == UNDECIDED) { state[w] = TAKEN; agg[w] = id; }
    }
    if (square)
      for (PetscInt k = G.i[v]; k < G.i[v + 1]; ++k) {
        const PetscInt w = G.j[k];
        for (PetscInt l = G.i[w]; l < G.i[w + 1]; ++l) {
          const PetscInt u = G.j[l];
          if (state[u] == UNDECIDED) { state[u] = TAKEN; agg[u] = id; }
        }
      }
  }
  return nagg;
}

// formProl0 [P376] for one near-null-space vector: column a of P0 = B restricted to aggregate a,
// normalised; the norm is the coarse B (the 1x1 R factor of the QR).
void tentative_prolongator(const std::vector<PetscInt> &agg, PetscInt nagg, const std::vector<MatScalar> &B, Csr &P0,
                           std::vector<MatScalar> &Bc)
{
  const PetscInt m = (PetscInt)agg.size();
  Bc.assign((size_t)nagg, 0.0);
  for (PetscInt v = 0; v < m; ++v) if (agg[v] >= 0) Bc[agg[v]] += B[v] * B[v];
  for (PetscInt a = 0; a < nagg; ++a) Bc[a] = std::sqrt(Bc[a]);
  P0.m = m; P0.n = nagg;
  P0.i.assign((size_t)m + 1, 0);
  P0.j.clear(); P0.a.clear();
  for (PetscInt v = 0; v < m; ++v) {
    if (agg[v] >= 0) { P0.j.push_back(agg[v]); P0.a.push_back(B[v] / Bc[agg[v]]); }
    P0.i[v + 1] = (PetscInt)P0.j.size();
  }
}

// P = P0 + alpha * D^-1 (A P0), alpha = -1.4 / emax (PCGAMGOptProlongator_AGG [P376])
PetscErrorCode smooth_prolongator(const Csr &AP0, const Csr &P0, const std::vector<MatScalar> &d, PetscReal alpha, Csr &P)
{
  return build_rows(P0.m, P0.n, true, P, [&](int, PetscInt r, std::vector<PetscInt> &pj, std::vector<MatScalar> &pa) {
    const MatScalar dinv = (d[r] != 0.0) ? 1.0 / d[r] : 1.0;
    PetscInt        k = AP0.i[r], l = P0.i[r];
    const PetscInt  ke = AP0.i[r + 1], le = P0.i[r + 1];
    while (k < ke || l < le) {  // merge of two ascending rows
      if (l >= le || (k < ke && AP0.j[k] < P0.j[l])) { pj.push_back(AP0.j[k]); pa.push_back(alpha * (dinv * AP0.a[k])); ++k; }
      else if (k >= ke || P0.j[l] < AP0.j[k]) { pj.push_back(P0.j[l]); pa.push_back(P0.a[l]); ++l; }
      else { pj.push_back(P0.j[l]); pa.push_back(P0.a[l] + alpha * (dinv * AP0.a[k])); ++k; ++l; }
    }
  });
}

// upper bound of lambda_max(D^-1 A) from the absolute row sums (-pc_gamg_b200_esteig gershgorin)
PetscReal gershgorin_emax(const CsrView &A, const std::vector<MatScalar> &d)
{
  PetscReal emax = 0.0;
  for (PetscInt r = 0; r < A.m; ++r) {
    PetscReal s = 0.0;
    for (PetscInt k = A.i[r]; k < A.i[r + 1]; ++k) s += std::fabs(A.a[k]);
    if (d[r] != 0.0) emax = std::max(emax, s / std::fabs(d[r]));
  }
  return emax;
}

PetscErrorCode mat_view(Mat A, CsrView &v)
{
  PetscInt nz;
  return MatSeqAIJGetCSRB200(A, &v.m, &v.n, &nz, &v.i, &v.j, &v.a);
}

std::string option(const char *name, const char *dflt)
{
  char      buf[256];
  PetscBool set = PETSC_FALSE;
  PetscOptionsGetString(NULL, NULL, name, buf, sizeof buf, &set);
  return set ? std::string(buf) : std::string(dflt);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// the hierarchy
// ---------------------------------------------------------------------------------------------
struct B200PCGamg {
  struct Level {
    Mat  A = NULL;      // level operator (level 0: the caller's matrix, not owned)
    Mat  P = NULL;      // prolongator from level+1 to this level (NULL on the coarsest)
    Vec  dinv = NULL;   // PCJACOBI of the level smoother / coarse solver
    Vec  b = NULL, x = NULL, xt = NULL, r = NULL;  // b and x of level 0 are the caller's
    std::vector<PetscInt> agg;
    PetscInt  nagg = 0;
    PetscReal emax = 0.0;
  };
  std::vector<Level> lv;
  PetscInt sweeps = 1;
};

PetscInt b200_pcgamg_num_levels(const B200PCGamg *mg) { return (PetscInt)mg->lv.size(); }

PetscErrorCode b200_pcgamg_level(const B200PCGamg *mg, PetscInt level, Mat *A, Mat *P, Vec *dinv, const PetscInt **agg,
                                 PetscInt *nagg, PetscReal *emax)
{
  if (level < 0 || level >= (PetscInt)mg->lv.size()) SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "no level %d", level);
  const B200PCGamg::Level &L = mg->lv[level];
  if (A) *A = L.A;
  if (P) *P = L.P;
  if (dinv) *dinv = L.dinv;
  if (agg) *agg = L.agg.empty() ? NULL : L.agg.data();
  if (nagg) *nagg = L.nagg;
  if (emax) *emax = L.emax;
  return 0;
}

static PetscErrorCode level_handle(Mat A, b200_csr_t *h)
{
  Mat_SeqAIJ *a  = (Mat_SeqAIJ *)A->data;
  int         rc = b200_petsc_ensure_resident(&A->spptr, A->rmap->n, A->cmap->n, a->i, a->j, a->a, (int64_t)A->state);
  if (rc) return PetscError(PETSC_COMM_SELF, __LINE__, __func__, __FILE__, rc, "level operator is not resident");
  *h = (b200_csr_t)b200_petsc_handle(&A->spptr);
  return 0;
}

static PetscErrorCode check_level_options(B200PCGamg *mg)
{
  // the smoother / coarse-solver combination the reference's options file selects
  // (configs/PETSc_SolverOptions_GAMG.info:9-20); anything else is refused, not approximated
  struct { const char *name, *dflt, *allowed; } want[] = {
      {"-pc_gamg_type", "agg", "agg"},
      {"-mg_levels_ksp_type", "richardson", "richardson"},
      {"-mg_levels_pc_type", "bjacobi", "bjacobi jacobi"},
      {"-mg_levels_sub_ksp_type", "preonly", "preonly"},
      {"-mg_levels_sub_pc_type", "jacobi", "jacobi"},
      {"-mg_coarse_ksp_type", "preonly", "preonly"},
      {"-mg_coarse_pc_type", "bjacobi", "bjacobi jacobi"},
      {"-mg_coarse_sub_ksp_type", "preonly", "preonly"},
      {"-mg_coarse_sub_pc_type", "jacobi", "jacobi"},
  };
  for (auto &w : want) {
    const std::string v = option(w.name, w.dflt);
    const std::string allowed = std::string(" ") + w.allowed + " ";
    if (allowed.find(" " + v + " ") == std::string::npos)
      SETERRQ3(PETSC_COMM_SELF, PETSC_ERR_SUP, "%s %s is not supported (supported: %s)", w.name, v.c_str(), w.allowed);
  }
  PetscErrorCode ierr;
  mg->sweeps = 1;
  ierr = PetscOptionsGetInt(NULL, NULL, "-mg_levels_ksp_max_it", &mg->sweeps, NULL);CHKERRQ(ierr);
  if (mg->sweeps < 1) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_OUTOFRANGE, "-mg_levels_ksp_max_it must be at least 1");
  return 0;
}

static PetscErrorCode jacobi_vec(Mat A, Vec *dinv)
{
  // PCSetUp_Jacobi [P376]: reciprocal of the diagonal, zero entries -> 1
  PetscErrorCode ierr;
  PetscScalar   *d;
  PetscInt       n;
  ierr = MatCreateVecs(A, NULL, dinv);CHKERRQ(ierr);
  ierr = MatGetDiagonal(A, *dinv);CHKERRQ(ierr);
  ierr = VecGetLocalSize(*dinv, &n);CHKERRQ(ierr);
  ierr = VecGetArray(*dinv, &d);CHKERRQ(ierr);
  for (PetscInt i = 0; i < n; ++i) d[i] = (d[i] != 0.0) ? 1.0 / d[i] : 1.0;
  return VecRestoreArray(*dinv, &d);
}

PetscErrorCode b200_pcgamg_setup(Mat Afine, B200PCGamg **out)
{
  PetscErrorCode ierr;
  B200PCGamg    *mg = new B200PCGamg;
  *out = mg;
  ierr = check_level_options(mg);CHKERRQ(ierr);

  PetscInt  nsmooths = 1, coarse_eq_limit = 50, max_levels = 30, square_graph = 1, est_its = 10;
  PetscReal threshold = 0.0, emax_given = 0.0;
  PetscBool emax_set = PETSC_FALSE;
  ierr = PetscOptionsGetInt(NULL, NULL, "-pc_gamg_agg_nsmooths", &nsmooths, NULL);CHKERRQ(ierr);
  ierr = PetscOptionsGetReal(NULL, NULL, "-pc_gamg_threshold", &threshold, NULL);CHKERRQ(ierr);
  ierr = PetscOptionsGetInt(NULL, NULL, "-pc_gamg_coarse_eq_limit", &coarse_eq_limit, NULL);CHKERRQ(ierr);
  ierr = PetscOptionsGetInt(NULL, NULL, "-pc_mg_levels", &max_levels, NULL);CHKERRQ(ierr);
  ierr = PetscOptionsGetInt(NULL, NULL, "-pc_gamg_square_graph", &square_graph, NULL);CHKERRQ(ierr);
  ierr = PetscOptionsGetInt(NULL, NULL, "-pc_gamg_b200_esteig_its", &est_its, NULL);CHKERRQ(ierr);
  ierr = PetscOptionsGetReal(NULL, NULL, "-pc_gamg_b200_emax", &emax_given, &emax_set);CHKERRQ(ierr);
  const std::string esteig = option("-pc_gamg_b200_esteig", "cg");
  if (nsmooths < 0 || nsmooths > 1) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_SUP, "-pc_gamg_agg_nsmooths: only 0 and 1");
  if (esteig != "cg" && esteig != "gershgorin") SETERRQ(PETSC_COMM_SELF, PETSC_ERR_SUP, "-pc_gamg_b200_esteig: cg or gershgorin");
  if (max_levels < 1) max_levels = 1;

  mg->lv.emplace_back();
  mg->lv[0].A = Afine;
  double    t_graph = 0, t_agg = 0, t_prol = 0, t_ptap = 0, t_mat = 0;
  Stopwatch sw;
  std::vector<MatScalar> B;  // near-null-space vector of the current level
  {
    PetscInt m;
    ierr = MatGetLocalSize(Afine, &m, NULL);CHKERRQ(ierr);
    B.assign((size_t)m, 1.0);
  }
  while ((PetscInt)mg->lv.size() < max_levels) {
    const size_t l = mg->lv.size() - 1;
    CsrView      A;
    ierr = mat_view(mg->lv[l].A, A);CHKERRQ(ierr);
    if (A.m != A.n) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_SIZ, "PCGAMG needs a square operator");
    if (l > 0 && A.m <= coarse_eq_limit) break;

    std::vector<MatScalar> d;
    diagonal(A, d);
    Csr G;
    sw.lap();
    ierr = strength_graph(A, d, threshold, G);CHKERRQ(ierr);
    t_graph += sw.lap();
    std::vector<PetscInt> agg;
    const PetscInt        nagg = aggregate(G, (PetscInt)l < square_graph, agg);
    t_agg += sw.lap();
    G = Csr();
    if (nagg == 0 || nagg >= A.m) break;  // nothing to coarsen

    Csr                    P0, P;
    std::vector<MatScalar> Bc;
    tentative_prolongator(agg, nagg, B, P0, Bc);
    PetscReal emax = 0.0;
    if (nsmooths == 1) {
      if (emax_set) emax = emax_given;
      else if (esteig == "gershgorin") emax = gershgorin_emax(A, d);
      else { ierr = b200_ksp_estimate_emax(mg->lv[l].A, est_its, &emax);CHKERRQ(ierr); }
      if (!(emax > 0.0)) SETERRQ1(PETSC_COMM_SELF, PETSC_ERR_CONV_FAILED, "eigenvalue estimate %g is not positive", emax);
      Csr AP0;
      ierr = spgemm(A, P0.view(), AP0);CHKERRQ(ierr);
      ierr = smooth_prolongator(AP0, P0, d, -1.4 / emax, P);CHKERRQ(ierr);
    } else P = std::move(P0);
    t_prol += sw.lap();

    Csr AP, PT, Ac;
    ierr = spgemm(A, P.view(), AP);CHKERRQ(ierr);
    transpose(P.view(), PT);
    ierr = spgemm(PT.view(), AP.view(), Ac);CHKERRQ(ierr);
    AP = Csr(); PT = Csr();
    t_ptap += sw.lap();

    mg->lv[l].agg  = std::move(agg);
    mg->lv[l].nagg = nagg;
    mg->lv[l].emax = emax;
    ierr = MatCreateSeqAIJFromCSRB200(P.m, P.n, P.i.data(), P.j.data(), P.a.data(), &mg->lv[l].P);CHKERRQ(ierr);
    mg->lv.emplace_back();
    ierr = MatCreateSeqAIJFromCSRB200(Ac.m, Ac.n, Ac.i.data(), Ac.j.data(), Ac.a.data(), &mg->lv[l + 1].A);CHKERRQ(ierr);
    B = std::move(Bc);
    t_mat += sw.lap();
  }

  for (size_t l = 0; l < mg->lv.size(); ++l) {
    B200PCGamg::Level &L = mg->lv[l];
    ierr = jacobi_vec(L.A, &L.dinv);CHKERRQ(ierr);
    if (l > 0) {
      ierr = VecDuplicate(L.dinv, &L.b);CHKERRQ(ierr);
      ierr = VecDuplicate(L.dinv, &L.x);CHKERRQ(ierr);
    }
    if (l + 1 < mg->lv.size()) {
      ierr = VecDuplicate(L.dinv, &L.xt);CHKERRQ(ierr);
      ierr = VecDuplicate(L.dinv, &L.r);CHKERRQ(ierr);
    }
  }
  // PCSetUp ends with everything the solve needs in HBM: level operators, prolongators and their
  // explicit transposes (the restriction), so that KSPSolve's time is the solve alone
  double t_dev = 0.0;
  if (b200_device_sm_count() > 0) {
    sw.lap();
    for (size_t l = 0; l < mg->lv.size(); ++l) {
      b200_csr_t h;
      ierr = level_handle(mg->lv[l].A, &h);CHKERRQ(ierr);
      if (mg->lv[l].P) {
        ierr = level_handle(mg->lv[l].P, &h);CHKERRQ(ierr);
        int rc = b200_csr_build_transpose(h);
        if (rc) return PetscError(PETSC_COMM_SELF, __LINE__, __func__, __FILE__, rc, "b200_csr_build_transpose");
      }
    }
    t_dev = sw.lap();
  }
  PetscBool view = PETSC_FALSE;
  ierr = PetscOptionsGetString(NULL, NULL, "-pc_gamg_b200_view", NULL, 0, &view);CHKERRQ(ierr);
  if (view) {
    for (size_t l = 0; l < mg->lv.size(); ++l) {
      PetscInt m, nz;
      ierr = MatSeqAIJGetCSRB200(mg->lv[l].A, &m, NULL, &nz, NULL, NULL, NULL);CHKERRQ(ierr);
      ierr = PetscPrintf(PETSC_COMM_WORLD, "[b200] gamg level %d: rows %d nz %d aggregates %d emax %.6f\n", (int)l, m, nz,
                         mg->lv[l].nagg, mg->lv[l].emax);CHKERRQ(ierr);
    }
    ierr = PetscPrintf(PETSC_COMM_WORLD, "[b200] gamg setup seconds: graph %.2f aggregate %.2f prolongator %.2f PtAP %.2f level objects %.2f upload %.2f (%d threads)\n",
                       t_graph, t_agg, t_prol, t_ptap, t_mat, t_dev, setup_threads());CHKERRQ(ierr);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// PCApply_MG: multiplicative V-cycle (PCMGMCycle_Private [P376])
// ---------------------------------------------------------------------------------------------
namespace {

// xnew = x + dinv .* (b - A x): KSPRICHARDSON (scale 1) + PCJACOBI, one iteration
PetscErrorCode jacobi_sweep(Mat A, Vec dinv, Vec b, Vec x, Vec xnew)
{
  PetscErrorCode     ierr;
  b200_csr_t         h;
  const PetscScalar *dx, *db, *dd;
  PetscScalar       *dn;
  ierr = level_handle(A, &h);CHKERRQ(ierr);
  ierr = VecB200GetDeviceArrayRead(x, &dx);CHKERRQ(ierr);
  ierr = VecB200GetDeviceArrayRead(b, &db);CHKERRQ(ierr);
  ierr = VecB200GetDeviceArrayRead(dinv, &dd);CHKERRQ(ierr);
  ierr = VecB200GetDeviceArrayWrite(xnew, &dn);CHKERRQ(ierr);
  int rc = b200_spmv_jacobi_sweep(h, dx, db, dd, dn, b200_petsc_mode(), NULL);
  if (rc) return PetscError(PETSC_COMM_SELF, __LINE__, __func__, __FILE__, rc, "b200_spmv_jacobi_sweep");
  Mat_SeqAIJ *a = (Mat_SeqAIJ *)A->data;
  return PetscLogFlops(2.0 * a->nz + 2.0 * A->rmap->n);
}

// r = b - A x: PCMGResidualDefault (MatMult + VecAYPX)
PetscErrorCode residual(Mat A, Vec b, Vec x, Vec r)
{
  PetscErrorCode     ierr;
  b200_csr_t         h;
  const PetscScalar *dx, *db;
  PetscScalar       *dr;
  ierr = level_handle(A, &h);CHKERRQ(ierr);
  ierr = VecB200GetDeviceArrayRead(x, &dx);CHKERRQ(ierr);
  ierr = VecB200GetDeviceArrayRead(b, &db);CHKERRQ(ierr);
  ierr = VecB200GetDeviceArrayWrite(r, &dr);CHKERRQ(ierr);
  int rc = b200_spmv_residual(h, dx, db, dr, b200_petsc_mode(), NULL);
  if (rc) return PetscError(PETSC_COMM_SELF, __LINE__, __func__, __FILE__, rc, "b200_spmv_residual");
  Mat_SeqAIJ *a = (Mat_SeqAIJ *)A->data;
  return PetscLogFlops(2.0 * a->nz);
}

PetscErrorCode vcycle(B200PCGamg *mg, size_t l, Vec b, Vec x)
{
  PetscErrorCode     ierr;
  B200PCGamg::Level &L = mg->lv[l];
  if (l + 1 == mg->lv.size()) return VecPointwiseMult(x, b, L.dinv);  // mg_coarse: preonly + jacobi

  // `sweeps` smoothing steps alternate between the two buffers; the post-smoother must end in x
  const PetscInt k = mg->sweeps;
  Vec            cur = (k % 2) ? L.xt : x, other = (k % 2) ? x : L.xt;  // where the pre-smoothed iterate ends up
  {
    // step 1 starts from the zero guess: x1 = 0 + dinv .* (b - A 0)
    Vec w = (k % 2) ? cur : other;
    ierr = VecPointwiseMult(w, b, L.dinv);CHKERRQ(ierr);
    for (PetscInt s = 1; s < k; ++s) {
      Vec nxt = (w == cur) ? other : cur;
      ierr = jacobi_sweep(L.A, L.dinv, b, w, nxt);CHKERRQ(ierr);
      w = nxt;
    }
    // k odd: written cur, then (k-1) swaps (even) -> cur.  k even: written other, odd swaps -> cur.
  }
  ierr = residual(L.A, b, cur, L.r);CHKERRQ(ierr);
  B200PCGamg::Level &C = mg->lv[l + 1];
  ierr = MatMultTranspose(L.P, L.r, C.b);CHKERRQ(ierr);        // MatRestrict
  ierr = vcycle(mg, l + 1, C.b, C.x);CHKERRQ(ierr);            // (zero coarse guess is implicit)
  ierr = MatMultAdd(L.P, C.x, cur, cur);CHKERRQ(ierr);         // MatInterpolateAdd
  for (PetscInt s = 0; s < k; ++s) {
    ierr = jacobi_sweep(L.A, L.dinv, b, cur, other);CHKERRQ(ierr);
    std::swap(cur, other);
  }
  // cur == x here by construction
  return 0;
}

}  // namespace

PetscErrorCode b200_pcgamg_apply(B200PCGamg *mg, Vec r, Vec z)
{
  if (!mg || mg->lv.empty()) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_WRONGSTATE, "PCGAMG is not set up");
  if (r == z) SETERRQ(PETSC_COMM_SELF, 61, "PCApply: r and z must be different vectors");
  return vcycle(mg, 0, r, z);
}

PetscErrorCode b200_pcgamg_destroy(B200PCGamg **pmg)
{
  if (!pmg || !*pmg) return 0;
  PetscErrorCode ierr;
  B200PCGamg    *mg = *pmg;
  for (size_t l = 0; l < mg->lv.size(); ++l) {
    B200PCGamg::Level &L = mg->lv[l];
    if (l > 0) { ierr = MatDestroy(&L.A);CHKERRQ(ierr); }
    ierr = MatDestroy(&L.P);CHKERRQ(ierr);
    ierr = VecDestroy(&L.dinv);CHKERRQ(ierr);
    ierr = VecDestroy(&L.b);CHKERRQ(ierr);
    ierr = VecDestroy(&L.x);CHKERRQ(ierr);
    ierr = VecDestroy(&L.xt);CHKERRQ(ierr);
    ierr = VecDestroy(&L.r);CHKERRQ(ierr);
  }
  delete mg;
  *pmg = NULL;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Host CSR utilities and one coarsening step behind a C ABI (include/b200_gamg.h): the pieces the
// row-partitioned multigrid set-up (petsc-openacc_b200/dgamg.py, one process per GPU) composes.
// Same kernels as the single-process set-up above; no device code.
// ---------------------------------------------------------------------------------------------
#include "../../../include/b200_gamg.h"

struct b200_hcsr_s { Csr c; };

namespace {
int hcsr_fail(int code, const char *what)
{
  PetscError(PETSC_COMM_SELF, __LINE__, "b200_hcsr", __FILE__, code, "%s", what);
  return code;
}
}  // namespace

extern "C" int b200_hcsr_create(b200_hcsr_t *out, int32_t m, int32_t n, const int32_t *ai, const int32_t *aj, const double *aa)
{
  if (!out || m < 0 || n < 0 || !ai || ai[0] != 0) return hcsr_fail(PETSC_ERR_ARG_OUTOFRANGE, "bad CSR");
  for (int32_t r = 0; r < m; ++r) {
    if (ai[r + 1] < ai[r]) return hcsr_fail(PETSC_ERR_ARG_OUTOFRANGE, "row pointers decrease");
    for (int32_t k = ai[r]; k < ai[r + 1]; ++k)
      if (aj[k] < 0 || aj[k] >= n || (k > ai[r] && aj[k] <= aj[k - 1])) return hcsr_fail(PETSC_ERR_ARG_WRONG, "columns out of range or not strictly ascending");
  }
  b200_hcsr_s *h = new b200_hcsr_s;
  h->c.m = m; h->c.n = n;
  h->c.i.assign(ai, ai + m + 1);
  h->c.j.assign(aj, aj + ai[m]);
  h->c.a.assign(aa, aa + ai[m]);
  *out = h;
  return 0;
}
extern "C" int b200_hcsr_destroy(b200_hcsr_t h) { delete h; return 0; }
extern "C" int b200_hcsr_shape(b200_hcsr_t h, int32_t *m, int32_t *n, int32_t *nz)
{
  if (!h) return hcsr_fail(PETSC_ERR_ARG_WRONG, "null handle");
  if (m) *m = h->c.m;
  if (n) *n = h->c.n;
  if (nz) *nz = h->c.i.empty() ? 0 : h->c.i[h->c.m];
  return 0;
}
extern "C" int b200_hcsr_arrays(b200_hcsr_t h, const int32_t **ai, const int32_t **aj, const double **aa)
{
  if (!h) return hcsr_fail(PETSC_ERR_ARG_WRONG, "null handle");
  if (ai) *ai = h->c.i.data();
  if (aj) *aj = h->c.j.data();
  if (aa) *aa = h->c.a.data();
  return 0;
}
extern "C" int b200_hcsr_spgemm(b200_hcsr_t X, b200_hcsr_t Y, b200_hcsr_t *out)
{
  if (!X || !Y || !out) return hcsr_fail(PETSC_ERR_ARG_WRONG, "null handle");
  b200_hcsr_s *h = new b200_hcsr_s;
  int rc = spgemm(X->c.view(), Y->c.view(), h->c);
  if (rc) { delete h; return rc; }
  *out = h;
  return 0;
}
extern "C" int b200_hcsr_transpose(b200_hcsr_t X, b200_hcsr_t *out)
{
  if (!X || !out) return hcsr_fail(PETSC_ERR_ARG_WRONG, "null handle");
  b200_hcsr_s *h = new b200_hcsr_s;
  transpose(X->c.view(), h->c);
  *out = h;
  return 0;
}
// X + Y (same shape): an entry present in both is x + y, otherwise the one that is there
extern "C" int b200_hcsr_add(b200_hcsr_t X, b200_hcsr_t Y, b200_hcsr_t *out)
{
  if (!X || !Y || !out) return hcsr_fail(PETSC_ERR_ARG_WRONG, "null handle");
  if (X->c.m != Y->c.m || X->c.n != Y->c.n) return hcsr_fail(PETSC_ERR_ARG_SIZ, "b200_hcsr_add: shapes differ");
  b200_hcsr_s *h = new b200_hcsr_s;
  const Csr &x = X->c, &y = Y->c;
  int rc = build_rows(x.m, x.n, true, h->c, [&](int, PetscInt r, std::vector<PetscInt> &cj, std::vector<MatScalar> &ca) {
    PetscInt k = x.i[r], l = y.i[r];
    const PetscInt ke = x.i[r + 1], le = y.i[r + 1];
    while (k < ke || l < le) {
      if (l >= le || (k < ke && x.j[k] < y.j[l])) { cj.push_back(x.j[k]); ca.push_back(x.a[k]); ++k; }
      else if (k >= ke || y.j[l] < x.j[k]) { cj.push_back(y.j[l]); ca.push_back(y.a[l]); ++l; }
      else { cj.push_back(x.j[k]); ca.push_back(x.a[k] + y.a[l]); ++k; ++l; }
    }
  });
  if (rc) { delete h; return rc; }
  *out = h;
  return 0;
}
extern "C" int b200_hcsr_abs_row_sums(b200_hcsr_t A, double *out)
{
  if (!A || (!out && A->c.m)) return hcsr_fail(PETSC_ERR_ARG_WRONG, "null argument");
  for (PetscInt r = 0; r < A->c.m; ++r) {
    double s = 0.0;
    for (PetscInt k = A->c.i[r]; k < A->c.i[r + 1]; ++k) s += std::fabs(A->c.a[k]);
    out[r] = s;
  }
  return 0;
}

// One coarsening step on a square block (a whole matrix, or one rank's diagonal block): strength
// graph, greedy MIS aggregates (of the squared graph when `square`), tentative prolongator from B,
// and, when emax > 0, one smoothing step P = (I - 1.4/emax D^-1 A) P0 with this block as A.
extern "C" int b200_gamg_coarsen_block(b200_hcsr_t A, const double *B, double threshold, int square, double emax,
                                       int32_t *agg_out, int32_t *nagg_out, b200_hcsr_t *P_out, double *Bc_out)
{
  if (!A || !agg_out || !nagg_out || !P_out || (A->c.m && (!B || !Bc_out))) return hcsr_fail(PETSC_ERR_ARG_WRONG, "null argument");
  if (A->c.m != A->c.n) return hcsr_fail(PETSC_ERR_ARG_SIZ, "b200_gamg_coarsen_block: square block expected");
  const CsrView          Av = A->c.view();
  std::vector<MatScalar> d;
  diagonal(Av, d);
  Csr G;
  int rc = strength_graph(Av, d, threshold, G);
  if (rc) return rc;
  std::vector<PetscInt> agg;
  const PetscInt        nagg = aggregate(G, square != 0, agg);
  std::copy(agg.begin(), agg.end(), agg_out);
  *nagg_out = nagg;
  std::vector<MatScalar> Bv(B, B + A->c.m), Bc;
  b200_hcsr_s *h = new b200_hcsr_s;
  Csr          P0;
  tentative_prolongator(agg, nagg, Bv, P0, Bc);
  if (emax > 0.0 && nagg > 0) {
    Csr AP0;
    rc = spgemm(Av, P0.view(), AP0);
    if (!rc) rc = smooth_prolongator(AP0, P0, d, -1.4 / emax, h->c);
    if (rc) { delete h; return rc; }
  } else h->c = std::move(P0);
  std::copy(Bc.begin(), Bc.end(), Bc_out);
  *P_out = h;
  return 0;
}
