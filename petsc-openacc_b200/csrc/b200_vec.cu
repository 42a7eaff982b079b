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
#include <map>
#include <mutex>
#include <vector>

#include "b200_common.h"

using namespace b200;

namespace {

constexpr int RED_THREADS  = 256;
constexpr int RED_MAX_GRID = 1024;  // partial sums per reduction

// Reduction scratch: one slot (partial sums + arrival counter) per (device, stream).  Reductions
// queued on one stream run in order and the last block resets the counter, so a stream can reuse its
// slot; two streams, or two devices, never share one.
struct SlotKey {
  int          dev;
  cudaStream_t st;
  bool operator<(const SlotKey &o) const { return dev != o.dev ? dev < o.dev : st < o.st; }
};
std::map<SlotKey, RedSlot> g_slots;
std::mutex                 g_slots_mutex;

int alloc_slot(RedSlot *s)
{
  B200_CUDA_TRY(cudaMalloc((void **)&s->partials, sizeof(double) * 2 * RED_MAX_GRID));
  B200_CUDA_TRY(cudaMalloc((void **)&s->counter, sizeof(unsigned)));
  B200_CUDA_TRY(cudaMemset(s->counter, 0, sizeof(unsigned)));
  return B200_OK;
}

int red_scratch(cudaStream_t st, double **partials, unsigned **counter)
{
  RedSlot s;
  B200_TRY(b200::red_slot_for(st, &s));
  *partials = s.partials;
  *counter  = s.counter;
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
                                             int post_a, int cg_post = 0, double *sc = nullptr, int *st = nullptr)
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
    if (cg_post) cg_scalar_post(cg_post, sc, st);   // the CG scalar step of a one-GPU solve rides on the reduction
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
// KSPSolve_CG [P376] + PCJACOBI in three passes per iteration, all scalars and the convergence test
// in device memory (indices CG_* / CGI_* of b200_common.h):
//   k_cg_p     x += a_prev p (the previous iteration's update, deferred so that p is read once);
//              p = z + (beta / betaold) p with z = dinv .* r formed on the fly (no z vector)
//   MatMult    w = A p with (p, w) folded into its epilogue (k_stream EPI_DOT)
//   k_cg_r     r -= (beta / (p,w)) w;  (z,z), (z,r);  the last block rotates the scalars and tests
//              convergence -- once st[CGI_DONE] is set every later launch returns at once, so the host
//              queues iterations in chunks and only polls the state word.
// 10 vector reads/writes per iteration besides the MatMult (it was 13 + a separate dot pass).
// Every kernel here is launched with programmatic stream serialization: it waits for its
// predecessors FIRST and only then lets its successor start (the stream kernel reads the state word
// before its own wait and relies on this order).

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

__device__ __forceinline__ double2 ld2(const double *p, long long i2) { return reinterpret_cast<const double2 *>(p)[i2]; }
__device__ __forceinline__ void st2(double *p, long long i2, double2 v) { reinterpret_cast<double2 *>(p)[i2] = v; }

// x = 0; r = b; (z,z), (z,r) with z = dinv .* r     (start of the solve; zero initial guess)
__global__ void __launch_bounds__(RED_THREADS)
    k_cg_init(long long n, const double *__restrict__ b, const double *__restrict__ dinv, double *__restrict__ x,
              double *__restrict__ r, double *partials, unsigned *counter, double *sc, int *st, int post)
{
  pdl_wait();
  pdl_launch_dependents();
  double    zz = 0.0, zr = 0.0;
  long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x, s = (long long)gridDim.x * RED_THREADS;
  for (; i < n; i += s) {
    const double ri = b[i], zi = dinv[i] * ri;
    x[i] = 0.0;
    r[i] = ri;
    zz   = __fma_rn(zi, zi, zz);
    zr   = __fma_rn(zi, ri, zr);
  }
  block_reduce2<false>(zz, zr);
  grid_finish2<false>(zz, zr, partials, counter, sc + CG_ZZ, sc + CG_ZR, 0, post, sc, st);
}

// VEC: all pointers 16-byte aligned -> two elements per load/store
template <bool VEC>
__global__ void __launch_bounds__(256)
    k_cg_p(long long n, double *__restrict__ x, double *__restrict__ p, const double *__restrict__ r,
           const double *__restrict__ dinv, const double *sc, const int *st)
{
  pdl_wait();
  pdl_launch_dependents();
  if (st[CGI_DONE]) return;
  const bool   first = st[CGI_ITS] == 0;
  const double a = sc[CG_A], b = first ? 0.0 : sc[CG_BETA] / sc[CG_BETAOLD];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x, s = (long long)gridDim.x * blockDim.x;
  if (VEC) {
    const long long n2 = n >> 1;
    for (long long i = t; i < n2; i += s) {
      const double2 rv = ld2(r, i), dv = ld2(dinv, i);
      const double  z0 = dv.x * rv.x, z1 = dv.y * rv.y;
      if (first) st2(p, i, make_double2(z0, z1));
      else {
        const double2 pv = ld2(p, i), xv = ld2(x, i);
        st2(x, i, make_double2(__fma_rn(a, pv.x, xv.x), __fma_rn(a, pv.y, xv.y)));
        st2(p, i, make_double2(__fma_rn(b, pv.x, z0), __fma_rn(b, pv.y, z1)));
      }
    }
    if (t == 0 && (n & 1)) {
      const long long i = n - 1;
      const double    z = dinv[i] * r[i];
      if (first) p[i] = z;
      else { const double pv = p[i]; x[i] = __fma_rn(a, pv, x[i]); p[i] = __fma_rn(b, pv, z); }
    }
  } else {
    for (long long i = t; i < n; i += s) {
      const double z = dinv[i] * r[i];
      if (first) p[i] = z;
      else { const double pv = p[i]; x[i] = __fma_rn(a, pv, x[i]); p[i] = __fma_rn(b, pv, z); }
    }
  }
}

// r -= a w with a = beta / (p,w); (z,z), (z,r).  r, w, dinv are workspace vectors (16-byte aligned).
__global__ void __launch_bounds__(RED_THREADS)
    k_cg_r(long long n, double *__restrict__ r, const double *__restrict__ w, const double *__restrict__ dinv,
           double *partials, unsigned *counter, double *sc, int *st, int post)
{
  pdl_wait();
  pdl_launch_dependents();
  if (st[CGI_DONE]) return;
  const double a  = sc[CG_BETA] / sc[CG_DPI];
  double       zz = 0.0, zr = 0.0;
  const long long n2 = n >> 1;
  const long long t = (long long)blockIdx.x * RED_THREADS + threadIdx.x, s = (long long)gridDim.x * RED_THREADS;
  for (long long i = t; i < n2; i += s) {
    const double2 rv = ld2(r, i), wv = ld2(w, i), dv = ld2(dinv, i);
    const double  r0 = __fma_rn(-a, wv.x, rv.x), r1 = __fma_rn(-a, wv.y, rv.y);
    st2(r, i, make_double2(r0, r1));
    const double z0 = dv.x * r0, z1 = dv.y * r1;
    zz = __fma_rn(z0, z0, zz); zr = __fma_rn(z0, r0, zr);
    zz = __fma_rn(z1, z1, zz); zr = __fma_rn(z1, r1, zr);
  }
  if (t == 0 && (n & 1)) {
    const long long i  = n - 1;
    const double    ri = __fma_rn(-a, w[i], r[i]), zi = dinv[i] * ri;
    r[i] = ri;
    zz = __fma_rn(zi, zi, zz); zr = __fma_rn(zi, ri, zr);
  }
  block_reduce2<false>(zz, zr);
  grid_finish2<false>(zz, zr, partials, counter, sc + CG_ZZ, sc + CG_ZR, 0, post, sc, st);
}

// the deferred update of the last iteration
template <bool VEC>
__global__ void __launch_bounds__(256)
    k_cg_finish(long long n, double *__restrict__ x, const double *__restrict__ p, const double *sc, const int *st)
{
  pdl_wait();
  pdl_launch_dependents();
  if (st[CGI_ITS] == 0) return;
  const double a = sc[CG_A];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x, s = (long long)gridDim.x * blockDim.x;
  if (VEC) {
    const long long n2 = n >> 1;
    for (long long i = t; i < n2; i += s) {
      const double2 pv = ld2(p, i), xv = ld2(x, i);
      st2(x, i, make_double2(__fma_rn(a, pv.x, xv.x), __fma_rn(a, pv.y, xv.y)));
    }
    if (t == 0 && (n & 1)) x[n - 1] = __fma_rn(a, p[n - 1], x[n - 1]);
  } else {
    for (long long i = t; i < n; i += s) x[i] = __fma_rn(a, p[i], x[i]);
  }
}

int ew_grid(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)std::max(1, sm_count()) * 16)); }

template <int OP>
int reduce_launch(const double *x, const double *y, int64_t n, double *d_out, int post, cudaStream_t st)
{
  if (!d_out) return set_error(B200_ERR_ARG, "reduction needs a device output pointer");
  B200_TRY(ensure_device());
  double   *partials;
  unsigned *counter;
  B200_TRY(red_scratch(st, &partials, &counter));
  B200_LAUNCH((k_reduce<OP>), red_grid(n), RED_THREADS, 0, st, x, y, (long long)n, partials, counter, d_out, post);
  return B200_OK;
}

// ---- CG workspace: vectors, scalars, reduction scratch and the pinned state words of a solve, kept
// per host thread and device between solves (the in-process multi-rank mode drives one solve per
// thread on one device; a process may also drive several devices).
struct CgWorkspace {
  size_t       len = 0;          // doubles per vector (even)
  double      *buf = nullptr;    // r, p, w, dinv
  double      *sc = nullptr, *partials = nullptr, *dot_partials = nullptr;
  int         *st = nullptr;
  unsigned    *counters = nullptr;   // [0] k_cg_init / k_cg_r, [1] the MatMult's dot
  double      *h_sc = nullptr;
  int         *h_st = nullptr;       // 2 slots of CGI_NINT
  cudaEvent_t  ev[2] = {nullptr, nullptr}, e0 = nullptr, e1 = nullptr;
  bool         ready = false;
};
constexpr int DOT_PARTIALS = 4096;
struct CgWorkspaces {
  CgWorkspace w[64];
  ~CgWorkspaces()
  {
    for (auto &c : w) {
      if (!c.ready && !c.buf) continue;
      cudaFree(c.buf); cudaFree(c.sc); cudaFree(c.partials); cudaFree(c.dot_partials); cudaFree(c.st); cudaFree(c.counters);
      if (c.h_sc) cudaFreeHost(c.h_sc);
      if (c.h_st) cudaFreeHost(c.h_st);
      for (auto &e : c.ev) if (e) cudaEventDestroy(e);
      if (c.e0) cudaEventDestroy(c.e0);
      if (c.e1) cudaEventDestroy(c.e1);
    }
  }
};
thread_local CgWorkspaces g_cgws;

int cg_workspace(size_t n, CgWorkspace **out)
{
  DeviceState *d = device_state();
  if (!d) return B200_ERR_NO_DEVICE;
  CgWorkspace &c = g_cgws.w[d->ordinal];
  if (!c.ready) {
    B200_CUDA_TRY(cudaMalloc((void **)&c.sc, sizeof(double) * CG_NSCAL));
    B200_CUDA_TRY(cudaMalloc((void **)&c.st, sizeof(int) * CGI_NINT));
    B200_CUDA_TRY(cudaMalloc((void **)&c.partials, sizeof(double) * 2 * RED_MAX_GRID));
    B200_CUDA_TRY(cudaMalloc((void **)&c.dot_partials, sizeof(double) * DOT_PARTIALS));
    B200_CUDA_TRY(cudaMalloc((void **)&c.counters, sizeof(unsigned) * 2));
    B200_CUDA_TRY(cudaMemset(c.counters, 0, sizeof(unsigned) * 2));
    B200_CUDA_TRY(cudaHostAlloc((void **)&c.h_sc, sizeof(double) * CG_NSCAL, cudaHostAllocDefault));
    B200_CUDA_TRY(cudaHostAlloc((void **)&c.h_st, sizeof(int) * 2 * CGI_NINT, cudaHostAllocDefault));
    for (auto &e : c.ev) B200_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    B200_CUDA_TRY(cudaEventCreate(&c.e0));
    B200_CUDA_TRY(cudaEventCreate(&c.e1));
    c.ready = true;
  }
  const size_t need = (std::max<size_t>(n, 1) + 1) & ~(size_t)1;
  if (c.len < need) {
    cudaFree(c.buf); c.buf = nullptr; c.len = 0;
    B200_CUDA_TRY(cudaMalloc((void **)&c.buf, sizeof(double) * 4 * need));
    c.len = need;
  }
  *out = &c;
  return B200_OK;
}

}  // namespace

namespace b200 {
int red_slot_for(cudaStream_t st, RedSlot *out)
{
  DeviceState *d = device_state();
  if (!d) return B200_ERR_NO_DEVICE;
  std::lock_guard<std::mutex> lock(g_slots_mutex);
  RedSlot &s = g_slots[SlotKey{d->ordinal, st}];
  if (!s.partials) B200_TRY(alloc_slot(&s));
  *out = s;
  return B200_OK;
}
}  // namespace b200

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
// the row-partitioned solve: `mult_dot` is MatMult with (p, A p) folded in, `allreduce` sums device
// scalars over the ranks and runs the scalar step (absent on one GPU, where the scalar step rides
// on the last block of the reduction).  Per iteration: k_cg_p, MatMult, k_cg_r -- no host round trip:
// iterations are queued in chunks, the state word of chunk c is copied out behind it and read while
// chunk c + 1 runs; after convergence the rest of the queue drains as no-ops.
// ---------------------------------------------------------------------------------------------
namespace b200 {
int cg_jacobi_run(const CgOps &ops, const double *d_b, double *d_x, double rtol, double atol,
                  int32_t max_it, b200_cg_result_t *res, cudaStream_t st)
{
  B200_TRY(ensure_device());
  const int       m = ops.m;
  const long long n = m;
  CgWorkspace *ws = nullptr;
  B200_TRY(cg_workspace((size_t)m, &ws));
  if (ops.dot_partials > DOT_PARTIALS) return set_error(B200_ERR_STATE, "stream grid of %d CTAs exceeds the dot scratch", ops.dot_partials);
  double *r = ws->buf, *p = ws->buf + ws->len, *w = ws->buf + 2 * ws->len, *dinv = ws->buf + 3 * ws->len;
  double *sc = ws->sc;
  int    *sti = ws->st;
  const bool multi = (bool)ops.allreduce;
  const bool xvec  = (reinterpret_cast<uintptr_t>(d_x) & 15) == 0;
  const uint64_t l0 = b200_launch_count();
  const int gr = red_grid(n), ge = (int)std::max<int64_t>(1, std::min<int64_t>((n / 2 + 255) / 256, (int64_t)std::max(1, sm_count()) * 8));
  DotArgs dot;
  dot.partials = ops.dot_partials > 0 ? ws->dot_partials : nullptr;
  dot.counter  = ws->counters + 1;
  dot.out      = sc + CG_DPI;
  dot.skip     = sti + CGI_DONE;

  // scalars of the solve: tolerances and the iteration limit go down once
  for (int k = 0; k < CG_NSCAL; ++k) ws->h_sc[k] = 0.0;
  ws->h_sc[CG_RTOL] = rtol; ws->h_sc[CG_ATOL] = atol; ws->h_sc[CG_BETAOLD] = 1.0;
  int *h_init = ws->h_st;
  h_init[CGI_ITS] = 0; h_init[CGI_REASON] = 0; h_init[CGI_DONE] = 0; h_init[CGI_MAXIT] = max_it;
  B200_CUDA_TRY(cudaEventRecord(ws->e0, st));
  B200_CUDA_TRY(cudaMemcpyAsync(sc, ws->h_sc, sizeof(double) * CG_NSCAL, cudaMemcpyHostToDevice, st));
  B200_CUDA_TRY(cudaMemcpyAsync(sti, h_init, sizeof(int) * CGI_NINT, cudaMemcpyHostToDevice, st));
  if (m) B200_LAUNCH(k_diag_inv, (m + 127) / 128, 128, 0, st, m, ops.ai, ops.aj, ops.aa, dinv);
  B200_LAUNCH_PDL(k_cg_init, gr, RED_THREADS, 0, st, n, d_b, (const double *)dinv, d_x, r, ws->partials, ws->counters, sc, sti,
                  multi ? (int)CG_POST_NONE : (int)CG_POST_BEGIN);
  if (multi) B200_TRY(ops.allreduce(sc + CG_ZZ, 2, CG_POST_BEGIN, sc, sti, st));
  // the host copy of the initial state must not be overwritten before the copy above has run
  B200_CUDA_TRY(cudaStreamSynchronize(st));

  auto iteration = [&]() -> int {
    if (xvec) B200_LAUNCH_PDL((k_cg_p<true>), ge, 256, 0, st, n, d_x, p, (const double *)r, (const double *)dinv, (const double *)sc, (const int *)sti);
    else B200_LAUNCH_PDL((k_cg_p<false>), ge, 256, 0, st, n, d_x, p, (const double *)r, (const double *)dinv, (const double *)sc, (const int *)sti);
    B200_TRY(ops.mult_dot(p, w, dot, st));
    if (multi) B200_TRY(ops.allreduce(sc + CG_DPI, 1, CG_POST_NONE, sc, sti, st));
    B200_LAUNCH_PDL(k_cg_r, gr, RED_THREADS, 0, st, n, r, (const double *)w, (const double *)dinv, ws->partials, ws->counters, sc, sti,
                    multi ? (int)CG_POST_NONE : (int)CG_POST_ROTATE);
    if (multi) B200_TRY(ops.allreduce(sc + CG_ZZ, 2, CG_POST_ROTATE, sc, sti, st));
    return B200_OK;
  };

  // chunked queue: the state of chunk c is polled while chunk c + 1 runs
  int  queued = 0, slot = 0, pending = -1;
  bool done = false;
  const int chunk_max = std::max(1, env_int("B200_CG_CHUNK", 16));
  int  chunk = std::min(4, chunk_max);
  while (!done) {
    const int todo = std::min(chunk, max_it - queued);
    for (int k = 0; k < todo; ++k) B200_TRY(iteration());
    queued += todo;
    B200_CUDA_TRY(cudaMemcpyAsync(ws->h_st + slot * CGI_NINT, sti, sizeof(int) * CGI_NINT, cudaMemcpyDeviceToHost, st));
    B200_CUDA_TRY(cudaEventRecord(ws->ev[slot], st));
    if (pending >= 0) {
      B200_CUDA_TRY(cudaEventSynchronize(ws->ev[pending]));
      if (ws->h_st[pending * CGI_NINT + CGI_DONE]) done = true;
    }
    pending = slot;
    slot ^= 1;
    chunk = chunk_max;
    if (!done && queued >= max_it) {   // nothing left to queue: the device sets DONE at max_it at the latest
      B200_CUDA_TRY(cudaEventSynchronize(ws->ev[pending]));
      done = true;
    }
  }
  if (xvec) B200_LAUNCH_PDL((k_cg_finish<true>), ge, 256, 0, st, n, d_x, (const double *)p, (const double *)sc, (const int *)sti);
  else B200_LAUNCH_PDL((k_cg_finish<false>), ge, 256, 0, st, n, d_x, (const double *)p, (const double *)sc, (const int *)sti);
  B200_CUDA_TRY(cudaMemcpyAsync(ws->h_sc, sc, sizeof(double) * CG_NSCAL, cudaMemcpyDeviceToHost, st));
  B200_CUDA_TRY(cudaMemcpyAsync(ws->h_st, sti, sizeof(int) * CGI_NINT, cudaMemcpyDeviceToHost, st));
  B200_CUDA_TRY(cudaEventRecord(ws->e1, st));
  B200_CUDA_TRY(cudaEventSynchronize(ws->e1));
  float ms = 0.f;
  B200_CUDA_TRY(cudaEventElapsedTime(&ms, ws->e0, ws->e1));
  res->its      = ws->h_st[CGI_ITS];
  res->reason   = ws->h_st[CGI_REASON] ? ws->h_st[CGI_REASON] : -3;
  res->rnorm    = ws->h_sc[CG_DP];
  res->rnorm0   = ws->h_sc[CG_RNORM0];
  res->solve_ms = ms;
  res->launches = b200_launch_count() - l0;
  return B200_OK;
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
  if (mode == B200_MODE_FAST) mode = B200_MODE_EXACT_FMA;
  ops.dot_partials = stream_grid_of(A);
  ops.mult_dot = [A, mode](const double *p, double *w, const DotArgs &dot, cudaStream_t s) { return spmv_dot(A, p, w, mode, dot, s); };
  return cg_jacobi_run(ops, d_b, d_x, rtol, atol, max_it, res, (cudaStream_t)stream);
}
