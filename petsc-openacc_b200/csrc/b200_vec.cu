// b200_vec.cu -- KSPSolve_CG's vector operations as fused, device-resident kernels, and a
// PETSc-free CG driver.
//
// What it replaces: PETSc 3.7.6 VecDot/VecNorm/VecAXPY/VecAYPX (un-vendored; selected by
// configs/PETSc_SolverOptions_GAMG.info:1 and src/main_ksp.cpp:94, also used directly at
// src/main_ksp.cpp:120-121), each of which is a separate pass over host memory in the reference.
// Here the vectors never leave HBM, the CG scalars live in device memory, and the two updates
// + preconditioner + both reductions of an iteration are one kernel.
#include <algorithm>
#include <cmath>
#include <vector>

#include "b200_common.h"

using namespace b200;

namespace {

constexpr int RED_THREADS  = 256;
constexpr int RED_MAX_GRID = 1024;  // partial sums per reduction
constexpr int RED_SLOTS    = 16;

struct RedScratch {
  double   *partials = nullptr;  // RED_SLOTS * 2 * RED_MAX_GRID
  unsigned *counters = nullptr;  // RED_SLOTS
  int       next     = 0;
};
RedScratch g_red;

int red_scratch(double **partials, unsigned **counter)
{
  if (!g_red.partials) {
    B200_CUDA_TRY(cudaMalloc((void **)&g_red.partials, sizeof(double) * RED_SLOTS * 2 * RED_MAX_GRID));
    B200_CUDA_TRY(cudaMalloc((void **)&g_red.counters, sizeof(unsigned) * RED_SLOTS));
    B200_CUDA_TRY(cudaMemset(g_red.counters, 0, sizeof(unsigned) * RED_SLOTS));
  }
  int s    = g_red.next;
  g_red.next = (s + 1) % RED_SLOTS;
  *partials = g_red.partials + (size_t)s * 2 * RED_MAX_GRID;
  *counter  = g_red.counters + s;
  return B200_OK;
}

int red_grid(int64_t n)
{
  int64_t want = (n + RED_THREADS * 4 - 1) / (RED_THREADS * 4);
  int     cap  = std::min(RED_MAX_GRID, std::max(1, sm_count() * 4));
  return (int)std::max<int64_t>(1, std::min<int64_t>(want, cap));
}

enum { OP_DOT = 0, OP_SUM = 1, OP_MAXABS = 2 };

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

// Block reduction of up to two values; result valid in thread 0.
template <bool MAX>
__device__ __forceinline__ void block_reduce2(double &a, double &b)
{
  __shared__ double sa[RED_THREADS / 32], sb[RED_THREADS / 32];
  a = MAX ? warp_max(a) : warp_sum(a);
  b = MAX ? warp_max(b) : warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sa[w] = a; sb[w] = b; }
  __syncthreads();
  if (w == 0) {
    a = (l < RED_THREADS / 32) ? sa[l] : (MAX ? 0.0 : 0.0);
    b = (l < RED_THREADS / 32) ? sb[l] : 0.0;
    a = MAX ? warp_max(a) : warp_sum(a);
    b = MAX ? warp_max(b) : warp_sum(b);
  }
}

// Deterministic grid reduction: every block writes its partial, the last block to finish sums
// the partials in index order.  post: 0 = raw, 1 = sqrt.
template <bool MAX>
__device__ __forceinline__ void grid_finish2(double a, double b, double *partials,
                                             unsigned *counter, double *out_a, double *out_b,
                                             int post_a)
{
  __shared__ bool last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x]                = a;
    partials[RED_MAX_GRID + blockIdx.x] = b;
    __threadfence();
    unsigned t = atomicAdd(counter, 1u);
    last       = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double ta = 0.0, tb = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += RED_THREADS) {
    // strided by thread but combined in a fixed tree below: deterministic for a fixed grid
    double va = partials[i], vb = partials[RED_MAX_GRID + i];
    if (MAX) { ta = fmax(ta, va); tb = fmax(tb, vb); } else { ta += va; tb += vb; }
  }
  __syncthreads();
  block_reduce2<MAX>(ta, tb);
  if (threadIdx.x == 0) {
    if (out_a) *out_a = post_a ? sqrt(ta) : ta;
    if (out_b) *out_b = tb;
    *counter = 0;
  }
}

template <int OP>
__global__ void __launch_bounds__(RED_THREADS) k_reduce(const double *__restrict__ x,
                                                        const double *__restrict__ y, long long n,
                                                        double *partials, unsigned *counter,
                                                        double *out, int post)
{
  double    a = 0.0, b = 0.0;
  long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x;
  long long s = (long long)gridDim.x * RED_THREADS;
  for (; i < n; i += s) {
    if (OP == OP_DOT) a = __fma_rn(x[i], y[i], a);
    else if (OP == OP_SUM) a += x[i];
    else a = fmax(a, fabs(x[i]));
  }
  block_reduce2<OP == OP_MAXABS>(a, b);
  grid_finish2<OP == OP_MAXABS>(a, b, partials, counter, out, nullptr, post);
}

__global__ void k_set(double *x, double a, long long n)
{
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, s = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += s) x[i] = a;
}
__global__ void k_vcopy(double *__restrict__ y, const double *__restrict__ x, long long n)
{
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, s = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += s) y[i] = x[i];
}
__global__ void k_axpy(double *__restrict__ y, double a, const double *__restrict__ x, long long n)
{
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, s = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += s) y[i] = __fma_rn(a, x[i], y[i]);
}
__global__ void k_aypx(double *__restrict__ y, double a, const double *__restrict__ x, long long n)
{
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, s = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += s) y[i] = __fma_rn(a, y[i], x[i]);
}
__global__ void k_pmult(double *__restrict__ w, const double *__restrict__ x, const double *__restrict__ y, long long n)
{
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, s = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += s) w[i] = x[i] * y[i];
}

// ------------------------------- CG pieces ---------------------------------------------------
// scalars in device memory: sc[0]=beta sc[1]=betaold sc[2]=dpi sc[3]=dp (norm) sc[4]=beta_new
enum { S_BETA = 0, S_BETAOLD = 1, S_DPI = 2, S_DP = 3, S_BETANEW = 4, S_ZZ = 5, S_ZR = 6, S_COUNT = 8 };  // S_ZZ, S_ZR, S_DPI: raw (all-reducible) sums

// dinv[i] = 1/a_ii (PCJACOBI [P376]: zero diagonal -> 1)
__global__ void k_diag_inv(int m, const int *__restrict__ ii, const int *__restrict__ aj,
                           const double *__restrict__ aa, double *dinv)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double d = 0.0;
  for (int k = ii[i]; k < ii[i + 1]; ++k) if (aj[k] == i) d = aa[k];
  dinv[i] = (d != 0.0) ? 1.0 / d : 1.0;
}

// z = dinv .* r ; dp = ||z|| ; beta = (z,r)      (start of the solve)
__global__ void __launch_bounds__(RED_THREADS) k_cg_init(long long n, const double *__restrict__ r,
                                                         const double *__restrict__ dinv,
                                                         double *__restrict__ z, double *partials,
                                                         unsigned *counter, double *sc)
{
  double    zz = 0.0, zr = 0.0;
  long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x, s = (long long)gridDim.x * RED_THREADS;
  for (; i < n; i += s) {
    double ri = r[i], zi = dinv[i] * ri;
    z[i] = zi;
    zz   = __fma_rn(zi, zi, zz);
    zr   = __fma_rn(zi, ri, zr);
  }
  block_reduce2<false>(zz, zr);
  grid_finish2<false>(zz, zr, partials, counter, sc + S_ZZ, sc + S_ZR, 0);
}

// after the (all-reduced) sums of k_cg_init: dp = sqrt(zz), beta = (z,r)
__global__ void k_cg_post_init(double *sc)
{
  sc[S_DP]   = sqrt(sc[S_ZZ]);
  sc[S_BETA] = sc[S_ZR];
}

// p = z + (beta/betaold) p   (first iteration: p = z);  betaold <- beta happens in k_cg_step
__global__ void k_cg_update_p(long long n, const double *__restrict__ z, double *__restrict__ p,
                              const double *sc, int first)
{
  const double b = first ? 0.0 : sc[S_BETA] / sc[S_BETAOLD];
  long long    i = (long long)blockIdx.x * blockDim.x + threadIdx.x, s = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += s) p[i] = first ? z[i] : __fma_rn(b, p[i], z[i]);
}

// a = beta/dpi; x += a p; r -= a w; z = dinv.*r; dp = ||z||; beta_new = (z,r)
__global__ void __launch_bounds__(RED_THREADS)
    k_cg_step(long long n, double *__restrict__ x, double *__restrict__ r, double *__restrict__ z,
              const double *__restrict__ p, const double *__restrict__ w,
              const double *__restrict__ dinv, double *partials, unsigned *counter, double *sc)
{
  const double a  = sc[S_BETA] / sc[S_DPI];
  double       zz = 0.0, zr = 0.0;
  long long    i = (long long)blockIdx.x * RED_THREADS + threadIdx.x, s = (long long)gridDim.x * RED_THREADS;
  for (; i < n; i += s) {
    x[i]      = __fma_rn(a, p[i], x[i]);
    double ri = __fma_rn(-a, w[i], r[i]);
    r[i]      = ri;
    double zi = dinv[i] * ri;
    z[i]      = zi;
    zz        = __fma_rn(zi, zi, zz);
    zr        = __fma_rn(zi, ri, zr);
  }
  block_reduce2<false>(zz, zr);
  grid_finish2<false>(zz, zr, partials, counter, sc + S_ZZ, sc + S_ZR, 0);
}

// after the (all-reduced) sums of k_cg_step: dp = sqrt(zz); betaold <- beta; beta <- (z,r)
__global__ void k_cg_rotate(double *sc)
{
  sc[S_DP]      = sqrt(sc[S_ZZ]);
  sc[S_BETANEW] = sc[S_ZR];
  sc[S_BETAOLD] = sc[S_BETA];
  sc[S_BETA]    = sc[S_ZR];
}

int ew_grid(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)std::max(1, sm_count()) * 16)); }

template <int OP>
int reduce_launch(const double *x, const double *y, int64_t n, double *d_out, int post, cudaStream_t st)
{
  if (!d_out) return set_error(B200_ERR_ARG, "reduction needs a device output pointer");
  B200_TRY(ensure_device());
  double   *partials;
  unsigned *counter;
  B200_TRY(red_scratch(&partials, &counter));
  B200_LAUNCH((k_reduce<OP>), red_grid(n), RED_THREADS, 0, st, x, y, (long long)n, partials, counter, d_out, post);
  return B200_OK;
}

}  // namespace

extern "C" int b200_vec_set(double *d_x, double a, int64_t n, void *stream)
{
  if (n <= 0) return B200_OK;
  B200_TRY(ensure_device());
  B200_LAUNCH(k_set, ew_grid(n), 256, 0, (cudaStream_t)stream, d_x, a, (long long)n);
  return B200_OK;
}
extern "C" int b200_vec_copy(double *d_y, const double *d_x, int64_t n, void *stream)
{
  if (n <= 0) return B200_OK;
  B200_TRY(ensure_device());
  B200_LAUNCH(k_vcopy, ew_grid(n), 256, 0, (cudaStream_t)stream, d_y, d_x, (long long)n);
  return B200_OK;
}
extern "C" int b200_vec_axpy(double *d_y, double a, const double *d_x, int64_t n, void *stream)
{
  if (n <= 0) return B200_OK;
  B200_TRY(ensure_device());
  B200_LAUNCH(k_axpy, ew_grid(n), 256, 0, (cudaStream_t)stream, d_y, a, d_x, (long long)n);
  return B200_OK;
}
extern "C" int b200_vec_aypx(double *d_y, double a, const double *d_x, int64_t n, void *stream)
{
  if (n <= 0) return B200_OK;
  B200_TRY(ensure_device());
  B200_LAUNCH(k_aypx, ew_grid(n), 256, 0, (cudaStream_t)stream, d_y, a, d_x, (long long)n);
  return B200_OK;
}
extern "C" int b200_vec_pointwise_mult(double *d_w, const double *d_x, const double *d_y, int64_t n, void *stream)
{
  if (n <= 0) return B200_OK;
  B200_TRY(ensure_device());
  B200_LAUNCH(k_pmult, ew_grid(n), 256, 0, (cudaStream_t)stream, d_w, d_x, d_y, (long long)n);
  return B200_OK;
}
extern "C" int b200_vec_dot(const double *d_x, const double *d_y, int64_t n, double *d_out, void *stream)
{
  return reduce_launch<OP_DOT>(d_x, d_y, n, d_out, 0, (cudaStream_t)stream);
}
extern "C" int b200_vec_norm2(const double *d_x, int64_t n, double *d_out, void *stream)
{
  return reduce_launch<OP_DOT>(d_x, d_x, n, d_out, 1, (cudaStream_t)stream);
}
extern "C" int b200_vec_norm_inf(const double *d_x, int64_t n, double *d_out, void *stream)
{
  return reduce_launch<OP_MAXABS>(d_x, nullptr, n, d_out, 0, (cudaStream_t)stream);
}
extern "C" int b200_vec_sum(const double *d_x, int64_t n, double *d_out, void *stream)
{
  return reduce_launch<OP_SUM>(d_x, nullptr, n, d_out, 0, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// KSPSolve_CG [P376] (left preconditioning, preconditioned-residual norm, zero initial guess,
// KSPConvergedDefault: rnorm < max(rtol*rnorm0, atol)), PCJACOBI.  One body for the single-GPU and
// the row-partitioned solve: `mult` is MatMult, `allreduce` sums device scalars over the ranks
// (absent on one GPU).  Per iteration: p-update, MatMult, dot, one fused x/r/z/norm/(z,r) kernel,
// one scalar kernel, one 64-byte read-back.
// ---------------------------------------------------------------------------------------------
namespace b200 {
int cg_jacobi_run(const CgOps &ops, const double *d_b, double *d_x, double rtol, double atol,
                  int32_t max_it, b200_cg_result_t *res, cudaStream_t st)
{
  B200_TRY(ensure_device());
  const int       m = ops.m;
  const long long n = m;
  double *buf = nullptr, *sc = nullptr, *h_sc = nullptr;
  B200_CUDA_TRY(cudaMalloc((void **)&buf, sizeof(double) * 5 * (size_t)std::max(m, 1)));
  B200_CUDA_TRY(cudaMalloc((void **)&sc, sizeof(double) * S_COUNT));
  B200_CUDA_TRY(cudaHostAlloc((void **)&h_sc, sizeof(double) * S_COUNT, cudaHostAllocDefault));
  double *r = buf, *z = buf + n, *p = buf + 2 * n, *w = buf + 3 * n, *dinv = buf + 4 * n;
  cudaEvent_t e0, e1;
  B200_CUDA_TRY(cudaEventCreate(&e0));
  B200_CUDA_TRY(cudaEventCreate(&e1));
  const uint64_t l0 = b200_launch_count();

  auto body = [&]() -> int {
    double   *partials;
    unsigned *counter;
    B200_CUDA_TRY(cudaEventRecord(e0, st));
    B200_CUDA_TRY(cudaMemsetAsync(sc, 0, sizeof(double) * S_COUNT, st));
    if (m) B200_LAUNCH(k_diag_inv, (m + 127) / 128, 128, 0, st, m, ops.ai, ops.aj, ops.aa, dinv);
    B200_TRY(b200_vec_set(d_x, 0.0, n, st));
    B200_TRY(b200_vec_copy(r, d_b, n, st));
    B200_TRY(red_scratch(&partials, &counter));
    B200_LAUNCH(k_cg_init, red_grid(n), RED_THREADS, 0, st, n, r, dinv, z, partials, counter, sc);
    if (ops.allreduce) B200_TRY(ops.allreduce(sc + S_ZZ, 2, st));
    B200_LAUNCH(k_cg_post_init, 1, 1, 0, st, sc);
    B200_CUDA_TRY(cudaMemcpyAsync(h_sc, sc, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, st));
    B200_CUDA_TRY(cudaStreamSynchronize(st));
    double dp = h_sc[S_DP];
    res->rnorm0 = dp;
    const double ttol = std::max(rtol * dp, atol);
    int it = 0;
    res->reason = 0;
    if (!(dp == dp)) res->reason = -9;  // KSP_DIVERGED_NANORINF
    else if (dp < ttol) res->reason = (dp < atol) ? 3 : 2;
    while (!res->reason && it < max_it) {
      B200_LAUNCH(k_cg_update_p, ew_grid(n), 256, 0, st, n, z, p, sc, it == 0);
      B200_TRY(ops.mult(p, w, st));
      B200_TRY(b200_vec_dot(p, w, n, sc + S_DPI, st));
      if (ops.allreduce) B200_TRY(ops.allreduce(sc + S_DPI, 1, st));
      B200_TRY(red_scratch(&partials, &counter));
      B200_LAUNCH(k_cg_step, red_grid(n), RED_THREADS, 0, st, n, d_x, r, z, p, w, dinv, partials, counter, sc);
      if (ops.allreduce) B200_TRY(ops.allreduce(sc + S_ZZ, 2, st));
      B200_LAUNCH(k_cg_rotate, 1, 1, 0, st, sc);
      B200_CUDA_TRY(cudaMemcpyAsync(h_sc, sc, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, st));
      B200_CUDA_TRY(cudaStreamSynchronize(st));
      dp = h_sc[S_DP];
      ++it;
      if (!(dp == dp)) res->reason = -9;
      else if (dp < ttol) res->reason = (dp < atol) ? 3 : 2;
    }
    if (!res->reason) res->reason = -3;  // KSP_DIVERGED_ITS
    res->its   = it;
    res->rnorm = dp;
    B200_CUDA_TRY(cudaEventRecord(e1, st));
    B200_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    B200_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    res->solve_ms = ms;
    res->launches = b200_launch_count() - l0;
    return B200_OK;
  };
  int rc = body();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  cudaFree(sc);
  cudaFreeHost(h_sc);
  return rc;
}
}  // namespace b200

extern "C" int b200_cg_jacobi(b200_csr_t A, const double *d_b, double *d_x, double rtol,
                              double atol, int32_t max_it, int mode, b200_cg_result_t *res,
                              void *stream)
{
  NvtxRange nvtx_("b200_cg_jacobi");
  if (!A || !d_b || !d_x || !res) return set_error(B200_ERR_ARG, "b200_cg_jacobi: null argument");
  B200_TRY(ensure_device());
  b200_csr_info_t info;
  B200_TRY(b200_csr_get_info(A, &info));
  if (info.m != info.n) return set_error(B200_ERR_ARG, "b200_cg_jacobi: matrix must be square");
  CgOps ops;
  ops.m = info.m;
  B200_TRY(b200_csr_device_arrays(A, &ops.ai, &ops.aj, &ops.aa));
  ops.mult = [A, mode](const double *p, double *w, cudaStream_t s) { return b200_spmv(A, p, w, mode, s); };
  return cg_jacobi_run(ops, d_b, d_x, rtol, atol, max_it, res, (cudaStream_t)stream);
}
