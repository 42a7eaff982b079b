// b200_workloads.cpp -- host-side generators of the benchmark matrices (product side).
//
// b200_gen_poisson7 builds, for one rank of a DMDA-decomposed N^3 grid, exactly the rows that
// the reference's generateA + setRefPoint produce (src/helper.cpp:161-279) without going through
// MatSetValues: a direct, multi-threaded two-pass construction used by bench.py and by callers
// that only need the assembled CSR.  The MatSetValues/MatAssemblyEnd route over the PETSc-shaped
// API lives in host/ and is tested to give identical arrays.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "b200_common.h"

using namespace b200;

namespace {

struct Decomp {
  int M, N, P, m, n, p, size;
  std::vector<int> xs, xl, ys, yl, zs, zl, base;
};

// PETSC_DECIDE factorisation of DMSetUp_DA_3D [P376] (src/helper.cpp:31-36 passes PETSC_DECIDE)
void decide(int M, int N, int P, int size, int &m, int &n, int &p)
{
  n = (int)(0.5 + std::pow(((double)N * N) * ((double)size) / ((double)P * M), 1. / 3.));
  if (n < 1) n = 1;
  for (; n > 0; --n) if (size % n == 0) break;
  if (n < 1) n = 1;
  m = (int)(0.5 + std::sqrt(((double)M) * ((double)size) / ((double)P * n)));
  if (m < 1) m = 1;
  p = 1;
  for (; m > 0; --m) { p = size / (m * n); if (m * n * p == size) break; }
  if (M > P && m < p) std::swap(m, p);
}

void split(int len, int parts, std::vector<int> &s, std::vector<int> &l)
{
  s.resize(parts); l.resize(parts);
  int at = 0;
  for (int q = 0; q < parts; ++q) {
    l[q] = len / parts + ((len % parts) > q);
    s[q] = at;
    at += l[q];
  }
}

int make_decomp(Decomp &d, int M, int N, int P, int size)
{
  if (M < 1 || N < 1 || P < 1 || size < 1) return set_error(B200_ERR_ARG, "bad grid/size");
  if ((long long)M * N * P > 2147483647LL / 7) return set_error(B200_ERR_ARG, "grid too large for int32 indices");
  d.M = M; d.N = N; d.P = P; d.size = size;
  decide(M, N, P, size, d.m, d.n, d.p);
  if (d.m * d.n * d.p != size) return set_error(B200_ERR_ARG, "cannot factor %d ranks", size);
  split(M, d.m, d.xs, d.xl); split(N, d.n, d.ys, d.yl); split(P, d.p, d.zs, d.zl);
  d.base.assign(size + 1, 0);
  for (int r = 0; r < size; ++r) {
    int pi = r % d.m, pj = (r / d.m) % d.n, pk = r / (d.m * d.n);
    d.base[r + 1] = d.base[r] + d.xl[pi] * d.yl[pj] * d.zl[pk];
  }
  return B200_OK;
}

inline int owner(const std::vector<int> &s, const std::vector<int> &l, int c)
{
  int q = 0;
  while (c >= s[q] + l[q]) ++q;
  return q;
}

// global PETSc-ordering index of an in-domain cell
inline int gid(const Decomp &d, int i, int j, int k)
{
  int pi = owner(d.xs, d.xl, i), pj = owner(d.ys, d.yl, j), pk = owner(d.zs, d.zl, k);
  int r  = pi + pj * d.m + pk * d.m * d.n;
  return d.base[r] + (i - d.xs[pi]) + (j - d.ys[pj]) * d.xl[pi] + (k - d.zs[pk]) * d.xl[pi] * d.yl[pj];
}

// Neumann diagonal of cell (i,j,k): 0 - v[present neighbours] in the order idx = 1..6 of
// src/helper.cpp:229-233 (i-1, i+1, j-1, j+1, k-1, k+1)
inline double diag_of(const Decomp &d, const double v[7], int i, int j, int k)
{
  double dg = 0.0;
  if (i > 0) dg -= v[1];
  if (i < d.M - 1) dg -= v[2];
  if (j > 0) dg -= v[3];
  if (j < d.N - 1) dg -= v[4];
  if (k > 0) dg -= v[5];
  if (k < d.P - 1) dg -= v[6];
  return dg;
}

void parallel_for(int n, const std::function<void(int, int)> &fn)
{
  int nt = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
  nt     = std::min(nt, std::max(1, n));
  if (nt <= 1) { fn(0, n); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t) {
    int a = (int)((long long)n * t / nt), b = (int)((long long)n * (t + 1) / nt);
    th.emplace_back([=, &fn] { fn(a, b); });
  }
  for (auto &t : th) t.join();
}
}  // namespace

// out[13] = m n p  xs ys zs  xm ym zm  nloc rstart nnz_local nghost_upper
extern "C" int b200_gen_poisson7_info(int M, int N, int P, int size, int rank, int32_t *out)
{
  Decomp d;
  B200_TRY(make_decomp(d, M, N, P, size));
  if (rank < 0 || rank >= size) return set_error(B200_ERR_ARG, "rank out of range");
  int pi = rank % d.m, pj = (rank / d.m) % d.n, pk = rank / (d.m * d.n);
  int xs = d.xs[pi], ys = d.ys[pj], zs = d.zs[pk], xm = d.xl[pi], ym = d.yl[pj], zm = d.zl[pk];
  long long nnz = 0;
  // 7 per cell minus one for every face of the global domain the sub-box touches
  nnz = 7LL * xm * ym * zm;
  if (xs == 0) nnz -= (long long)ym * zm;
  if (xs + xm == M) nnz -= (long long)ym * zm;
  if (ys == 0) nnz -= (long long)xm * zm;
  if (ys + ym == N) nnz -= (long long)xm * zm;
  if (zs == 0) nnz -= (long long)xm * ym;
  if (zs + zm == P) nnz -= (long long)xm * ym;
  int32_t v[12] = {d.m, d.n, d.p, xs, ys, zs, xm, ym, zm, xm * ym * zm, d.base[rank], (int32_t)nnz};
  memcpy(out, v, sizeof v);
  return B200_OK;
}

extern "C" int b200_gen_poisson7_bases(int M, int N, int P, int size, int32_t *base)
{
  Decomp d;
  B200_TRY(make_decomp(d, M, N, P, size));
  memcpy(base, d.base.data(), sizeof(int32_t) * (size + 1));
  return B200_OK;
}

// Rows of this rank with GLOBAL column ids ascending inside each row; optional rhs / exact.
// refpoint != 0 applies setRefPoint (src/helper.cpp:250-279).
extern "C" int b200_gen_poisson7(int M, int N, int P, int size, int rank, int refpoint,
                                 int32_t *ai, int32_t *aj, double *aa, double *rhs, double *exact)
{
  Decomp d;
  B200_TRY(make_decomp(d, M, N, P, size));
  if (rank < 0 || rank >= size || !ai || !aj || !aa) return set_error(B200_ERR_ARG, "bad argument");
  const int pi = rank % d.m, pj = (rank / d.m) % d.n, pk = rank / (d.m * d.n);
  const int xs = d.xs[pi], ys = d.ys[pj], zs = d.zs[pk], xm = d.xl[pi], ym = d.yl[pj], zm = d.zl[pk];
  const double dx = 1.0 / M, dy = 1.0 / N, dz = 1.0 / P;
  double v[7];
  v[0] = 0.0;
  v[1] = v[2] = 1.0 / (dx * dx);
  v[3] = v[4] = 1.0 / (dy * dy);
  v[5] = v[6] = 1.0 / (dz * dz);
  const int rstart = d.base[rank];

  // pass 1: row lengths -> ai
  ai[0] = 0;
  parallel_for(zm, [&](int ka, int kb) {
    for (int kk = ka; kk < kb; ++kk) {
      const int k = zs + kk;
      for (int jj = 0; jj < ym; ++jj) {
        const int j = ys + jj;
        int32_t  *row = ai + 1 + ((size_t)kk * ym + jj) * xm;
        const int base = 1 + (j > 0) + (j < N - 1) + (k > 0) + (k < P - 1);
        for (int ii = 0; ii < xm; ++ii) {
          const int i = xs + ii;
          row[ii]     = base + (i > 0) + (i < M - 1);
        }
      }
    }
  });
  const int nloc = xm * ym * zm;
  for (int r = 0; r < nloc; ++r) ai[r + 1] += ai[r];

  // pass 2: fill
  const double c1 = 2.0 * 1.0 * M_PI;
  const double c2 = -3.0 * 2.0 * 1.0 * M_PI * 2.0 * 1.0 * M_PI;  // the macro expansion of src/helper.cpp:17-18
  parallel_for(zm, [&](int ka, int kb) {
    std::vector<double> cx(xm);
    for (int ii = 0; ii < xm; ++ii) cx[ii] = std::cos(c1 * ((xs + ii) + 0.5) * dx);
    for (int kk = ka; kk < kb; ++kk) {
      const int    k  = zs + kk;
      const double cz = std::cos(c1 * (k + 0.5) * dz);
      for (int jj = 0; jj < ym; ++jj) {
        const int    j  = ys + jj;
        const double cy = std::cos(c1 * (j + 0.5) * dy);
        for (int ii = 0; ii < xm; ++ii) {
          const int    i = xs + ii;
          const size_t r = ((size_t)kk * ym + jj) * xm + ii;
          int          c[7];
          double       w[7];
          int          cnt = 0;
          auto put = [&](int col, double val) {
            int pos = cnt++;
            while (pos > 0 && c[pos - 1] > col) { c[pos] = c[pos - 1]; w[pos] = w[pos - 1]; --pos; }
            c[pos] = col; w[pos] = val;
          };
          put(rstart + (int)r, diag_of(d, v, i, j, k));
          if (i > 0) put(gid(d, i - 1, j, k), v[1]);
          if (i < M - 1) put(gid(d, i + 1, j, k), v[2]);
          if (j > 0) put(gid(d, i, j - 1, k), v[3]);
          if (j < N - 1) put(gid(d, i, j + 1, k), v[4]);
          if (k > 0) put(gid(d, i, j, k - 1), v[5]);
          if (k < P - 1) put(gid(d, i, j, k + 1), v[6]);
          int32_t *cj = aj + ai[r];
          double  *cv = aa + ai[r];
          for (int q = 0; q < cnt; ++q) { cj[q] = c[q]; cv[q] = w[q]; }
          if (rhs) rhs[r] = c2 * cx[ii] * cy * cz;
          if (exact) exact[r] = cx[ii] * cy * cz;
        }
      }
    }
  });

  if (refpoint) {
    // scale = VecSum(diag)/n: sequential sum per rank in row order, rank partials added in rank
    // order (MPI_Allreduce order is unspecified; documented in DESIGN.md)
    double total = 0.0;
    for (int r = 0; r < size; ++r) {
      int qi = r % d.m, qj = (r / d.m) % d.n, qk = r / (d.m * d.n);
      double lsum = 0.0;
      for (int k = d.zs[qk]; k < d.zs[qk] + d.zl[qk]; ++k)
        for (int j = d.ys[qj]; j < d.ys[qj] + d.yl[qj]; ++j)
          for (int i = d.xs[qi]; i < d.xs[qi] + d.xl[qi]; ++i) lsum += diag_of(d, v, i, j, k);
      total += lsum;
    }
    const double scale  = total / (double)((long long)M * N * P);
    const double exact0 = std::cos(c1 * 0.5 * dx) * std::cos(c1 * 0.5 * dy) * std::cos(c1 * 0.5 * dz);
    // MatZeroRowsColumns(A, 1, {0}, scale, exact, rhs): only rows adjacent to cell (0,0,0) hold
    // column 0, but scan all local rows like the reference does
    parallel_for(nloc, [&](int ra, int rb) {
      for (int r = ra; r < rb; ++r) {
        if (rstart + r == 0) continue;
        for (int q = ai[r]; q < ai[r + 1]; ++q)
          if (aj[q] == 0) {
            if (rhs) rhs[r] -= aa[q] * exact0;
            aa[q] = 0.0;
          }
      }
    });
    if (rstart == 0 && nloc > 0) {
      for (int q = ai[0]; q < ai[1]; ++q) aa[q] = (aj[q] == 0) ? scale : 0.0;
      if (rhs) rhs[0] = scale * exact0;
    }
  }
  return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// Synthetic irregular matrix of BASELINE configs[4]: m x n, row length
// L_i = min(lmax, max(1, floor(u_i^(-1/(alpha-1))))), columns = sorted unique splitmix64 draws mod n,
// values uniform [-1,1) indexed by the final non-zero position.  Counter based: identical to
// tests/gen.py::powerlaw for the same arguments.  Two calls: sizes (ai filled), then fill.
// ---------------------------------------------------------------------------------------------
namespace {
inline uint64_t splitmix64(uint64_t z)
{
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline int powerlaw_row(uint64_t seed, int64_t n, double alpha, int lmax, int64_t row, std::vector<int32_t> &cols)
{
  double u = (double)(splitmix64(seed ^ (uint64_t)row) >> 11) / 9007199254740992.0;
  if (u < 1e-300) u = 1e-300;
  double Ld = std::floor(std::pow(u, -1.0 / (alpha - 1.0)));
  int64_t L = (int64_t)std::min<double>((double)lmax, std::max(1.0, Ld));
  L = std::min<int64_t>(L, n);
  cols.resize((size_t)L);
  const uint64_t s2 = seed * 0x100000001B3ull;
  for (int64_t k = 0; k < L; ++k) cols[(size_t)k] = (int32_t)(splitmix64(s2 ^ ((uint64_t)row << 20) ^ (uint64_t)k) % (uint64_t)n);
  std::sort(cols.begin(), cols.end());
  cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
  return (int)cols.size();
}
}  // namespace

extern "C" int b200_gen_powerlaw_rowptr(int32_t m, int32_t n, double alpha, int32_t lmax, uint64_t seed, int32_t *ai)
{
  if (m < 0 || n < 1 || !ai || alpha <= 1.0 || lmax < 1) return set_error(B200_ERR_ARG, "b200_gen_powerlaw_rowptr: bad argument");
  parallel_for(m, [&](int a, int b) {
    std::vector<int32_t> cols;
    for (int r = a; r < b; ++r) ai[r + 1] = powerlaw_row(seed, n, alpha, lmax, r, cols);
  });
  ai[0] = 0;
  long long tot = 0;
  for (int r = 0; r < m; ++r) {
    tot += ai[r + 1];
    if (tot > 2147483647LL) return set_error(B200_ERR_ARG, "power-law matrix exceeds int32 non-zeros");
    ai[r + 1] = (int32_t)tot;
  }
  return B200_OK;
}

extern "C" int b200_gen_powerlaw_fill(int32_t m, int32_t n, double alpha, int32_t lmax, uint64_t seed, const int32_t *ai,
                                      int32_t *aj, double *aa)
{
  if (m < 0 || n < 1 || !ai || !aj || !aa) return set_error(B200_ERR_ARG, "b200_gen_powerlaw_fill: bad argument");
  const uint64_t vseed = seed ^ 0xABCDEFull;
  parallel_for(m, [&](int a, int b) {
    std::vector<int32_t> cols;
    for (int r = a; r < b; ++r) {
      powerlaw_row(seed, n, alpha, lmax, r, cols);
      for (size_t k = 0; k < cols.size(); ++k) {
        const int64_t p = (int64_t)ai[r] + (int64_t)k;
        aj[p] = cols[k];
        aa[p] = (double)(splitmix64(vseed ^ (uint64_t)p) >> 11) * (2.0 / 9007199254740992.0) - 1.0;
      }
    }
  });
  return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// BASELINE configs[3]: 27-point box stencil on an N^3 grid, non-periodic, natural
// ordering, columns ascending; off-diagonal -1, diagonal = number of neighbours; nnz = (3N-2)^3.
// seed != 0 replaces the values by uniform [-1,1) draws indexed by the non-zero position (identical
// to tests/gen.py::stencil27).  ai[N^3 + 1]; aj / aa may be NULL for a sizing call.
// ---------------------------------------------------------------------------------------------
extern "C" int b200_gen_stencil27(int32_t N, uint64_t seed, int32_t *ai, int32_t *aj, double *aa)
{
  if (N < 1 || !ai) return set_error(B200_ERR_ARG, "b200_gen_stencil27: bad argument");
  const long long n = (long long)N * N * N;
  if (n > 2147483647LL / 27) return set_error(B200_ERR_ARG, "27-point matrix exceeds int32 non-zeros");
  auto span = [N](int c) { return (c > 0) + 1 + (c < N - 1); };
  ai[0] = 0;
  for (int k = 0; k < N; ++k)
    for (int j = 0; j < N; ++j)
      for (int i = 0; i < N; ++i) {
        const long long r = ((long long)k * N + j) * N + i;
        ai[r + 1] = ai[r] + span(i) * span(j) * span(k);
      }
  if (!aj || !aa) return B200_OK;
  parallel_for((int)n, [&](int a, int b) {
    for (int r = a; r < b; ++r) {
      const int i = r % N, j = (r / N) % N, k = r / (N * N);
      long long p = ai[r];
      const int cnt = ai[r + 1] - ai[r];
      for (int dk = -1; dk <= 1; ++dk)
        for (int dj = -1; dj <= 1; ++dj)
          for (int di = -1; di <= 1; ++di) {
            if (i + di < 0 || i + di >= N || j + dj < 0 || j + dj >= N || k + dk < 0 || k + dk >= N) continue;
            aj[p] = r + di + dj * N + dk * N * N;
            if (seed) aa[p] = (double)(splitmix64(seed ^ (uint64_t)p) >> 11) * (2.0 / 9007199254740992.0) - 1.0;
            else aa[p] = (di == 0 && dj == 0 && dk == 0) ? (double)(cnt - 1) : -1.0;
            ++p;
          }
    }
  });
  return B200_OK;
}
