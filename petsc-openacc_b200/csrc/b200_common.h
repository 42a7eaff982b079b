// b200_common.h -- error plumbing, launch accounting and PTX helpers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <functional>
#include <string>

#include <nvtx3/nvToolsExt.h>

#include "../../include/b200_seqaij.h"

namespace b200 {

// NVTX ranges around the C-ABI calls (B200_NVTX=1): the counterpart of the reference's Score-P /
// nvprof instrumented builds (Makefile:137-150, runs/single-node-nvprof.pbs) for nsys / ncu.
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char *name);
  ~NvtxRange() { if (on) nvtxRangePop(); }
};

extern thread_local std::string g_last_error;
extern std::atomic<uint64_t>    g_launches;

int  set_error(int code, const char *fmt, ...);
int  ensure_device();        // B200_OK when an sm_100 device is current
int  sm_count();             // of the current device
int  env_int(const char *name, int dflt);

// State that belongs to one device (the process may drive several): capability check, SM count, the
// "opt-in shared memory limit raised" latch of the stream kernels (cudaFuncSetAttribute is per
// device), and the reduction scratch (one slot per stream: reductions on one stream are ordered,
// reductions on different streams never share partial sums).
struct RedSlot {
  double   *partials = nullptr;   // 2 * RED_MAX_GRID
  unsigned *counter  = nullptr;
};
struct DeviceState {
  int  ordinal = -1, sm_count = 0;
  bool checked = false, ok = false, stream_attrs_set = false;
};
DeviceState *device_state();                        // of the current device; nullptr on failure (error set)
int red_slot_for(cudaStream_t st, RedSlot *out);    // b200_vec.cu

#define B200_CUDA_TRY(expr)                                                                    \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return b200::set_error(B200_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,       \
                             cudaGetErrorString(_e));                                          \
  } while (0)

#define B200_TRY(expr)                                                                         \
  do {                                                                                         \
    int _r = (expr);                                                                           \
    if (_r) return _r;                                                                         \
  } while (0)

// every kernel launch of the library goes through this so b200_launch_count() is honest
#define B200_LAUNCH(kernel, grid, block, smem, stream, ...)                                    \
  do {                                                                                         \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                \
    b200::g_launches.fetch_add(1, std::memory_order_relaxed);                                  \
    B200_CUDA_TRY(cudaGetLastError());                                                         \
  } while (0)

// Programmatic dependent launch: the kernel may become resident while the previous kernel of the
// stream is still draining; everything it does before `griddepcontrol.wait` (shared-memory set-up,
// mbarrier init, the first bulk copies of the CONSTANT matrix arrays) overlaps that tail.  Only for
// kernels that execute pdl_wait() before touching anything an earlier kernel wrote.
// B200_PDL=0 launches them the ordinary way (A/B measurements).
#define B200_LAUNCH_PDL(kernel, grid_, block_, smem_, stream_, ...)                            \
  do {                                                                                         \
    cudaLaunchConfig_t    cfg_{};                                                              \
    cudaLaunchAttribute   att_[1];                                                             \
    cfg_.gridDim = dim3(grid_); cfg_.blockDim = dim3(block_);                                  \
    cfg_.dynamicSmemBytes = (smem_); cfg_.stream = (stream_);                                  \
    att_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                           \
    att_[0].val.programmaticStreamSerializationAllowed = 1;                                    \
    cfg_.attrs = att_; cfg_.numAttrs = b200::pdl_enabled() ? 1 : 0;                            \
    cudaError_t le_ = cudaLaunchKernelEx(&cfg_, kernel, __VA_ARGS__);                          \
    b200::g_launches.fetch_add(1, std::memory_order_relaxed);                                  \
    if (le_ != cudaSuccess)                                                                    \
      return b200::set_error(B200_ERR_CUDA, "%s:%d launch of %s -> %s", __FILE__, __LINE__,    \
                             #kernel, cudaGetErrorString(le_));                                \
  } while (0)
bool pdl_enabled();

// ---------------------------------------------------------------------------------------------
// Halo pieces shared by b200_halo.cu (owner) and the fused stream kernel in b200_spmv.cu.
// ---------------------------------------------------------------------------------------------
struct PushBlock {  // one CTA of the push role
  int32_t peer_slot, start, count, pad;
};
struct PushPeer {   // one destination rank
  double             *dst[2];  // peer lvec buffers (biased so that the flat element number indexes)
  unsigned long long *flag;    // &peer_window.flags[my rank]
  int32_t             nblocks, pad;
};
struct HaloArgs {
  // send side (VecScatterBegin)
  const PushBlock *blocks;
  const PushPeer  *peers;
  const int       *send_idx;
  unsigned        *done;
  int              npush;
  // tile schedule of this launch: the CTA-major copy of the tile table and, per CTA, its [first, last)
  // range in it (tiles dealt out by cost, ghost rows included, so that the CTAs that close many
  // ghost rows stream fewer tiles); nullptr = tile t belongs to CTA t % grid
  const int4      *sched_tiles;
  const int       *sched_first;
  // receive side (VecScatterEnd + MatMultAdd of the off-diagonal block, compressed row)
  const int       *cta_ptr;    // per stream CTA: [first, last) in cta_rows
  const int       *cta_rows;   // compressed-row positions of B grouped by the CTA that owns the row's tile
  const int       *cpi, *ridx, *bj;
  const double    *ba;
  const double    *lvec;
  const unsigned long long *flags;
  const int       *srcs;
  int              nsrc;
  unsigned long long seq;
  unsigned long long *err;
  unsigned long long  timeout_ns;
};
// MatMult with CG's (p, A p) folded into its epilogue (k_stream, EPI_DOT): every CTA leaves one
// partial sum, the last CTA to finish adds them in index order -> *out (deterministic for a plan).
// skip != nullptr and *skip != 0: the launch does nothing (the solve has converged on the device
// and the remaining launches of the chunk drain).
struct DotArgs {
  double    *partials = nullptr;  // one per CTA of the stream grid
  unsigned  *counter  = nullptr;
  double    *out      = nullptr;
  const int *skip     = nullptr;
};
// CG scalars and state words in device memory (b200_vec.cu, also finished by k_allreduce)
enum { CG_BETA = 0, CG_BETAOLD = 1, CG_DPI = 2, CG_DP = 3, CG_A = 4, CG_ZZ = 5, CG_ZR = 6, CG_TTOL = 7,
       CG_RNORM0 = 8, CG_RTOL = 9, CG_ATOL = 10, CG_NSCAL = 16 };
enum { CGI_ITS = 0, CGI_REASON = 1, CGI_DONE = 2, CGI_MAXIT = 3, CGI_NINT = 4 };
enum { CG_POST_NONE = 0, CG_POST_BEGIN = 1, CG_POST_ROTATE = 2 };
// the CG body shared by b200_cg_jacobi and b200_mpiaij_cg_jacobi (b200_vec.cu)
struct CgOps {
  int            m = 0;                       // local rows
  const int32_t *ai = nullptr, *aj = nullptr; // diagonal block (for PCJACOBI)
  const double  *aa = nullptr;
  int            dot_partials = 0;            // partial sums mult_dot needs (0: it reduces by itself)
  // w = A p and dot.out = (p, w) over the local rows
  std::function<int(const double *, double *, const DotArgs &, cudaStream_t)> mult_dot;
  // in place sum over the ranks of n device scalars, then the scalar step `post` on sc/st; empty on one GPU
  std::function<int(double *, int, int post, double *sc, int *st, cudaStream_t)> allreduce;
};
int cg_jacobi_run(const CgOps &ops, const double *d_b, double *d_x, double rtol, double atol,
                  int32_t max_it, b200_cg_result_t *res, cudaStream_t st);

// internal (not part of the C ABI): the stream plan of a matrix and the fused launch
int stream_plan_tiles(b200_csr_t A, int4 **d_tiles, int *ntiles, int *grid, int *threads);   // 0 tiles = not applicable
int launch_stream_halo(b200_csr_t A, const double *x, double *y, int mode, const HaloArgs &h,
                       cudaStream_t st, const DotArgs *dot = nullptr);
int stream_grid_of(b200_csr_t A);   // CTAs of the stream launch (0: the plan is not the stream kernel)
// the row blocks of the host-vector pipeline (tile aligned, ~B200_HOST_BLOCK_ROWS rows): block b covers
// rows [row0, row1) and reads x up to the end of block `need`; launch = A x for those rows only
int host_block_count(b200_csr_t A, int mode);
int host_block_info(b200_csr_t A, int b, int *row0, int *row1, int *need);
int launch_host_block(b200_csr_t A, int b, const double *x, double *y, int mode, cudaStream_t st);
// w = A x with (x, w) folded in when the plan is the stream kernel, else MatMult + a reduction
int spmv_dot(b200_csr_t A, const double *x, double *y, int mode, const DotArgs &dot, cudaStream_t st);

#ifdef __CUDACC__
// The scalar steps of KSPSolve_CG [P376] (one thread): after the first (z,z), (z,r) ...
__device__ __forceinline__ void cg_scalar_converged(double dp, double *sc, int *st)
{
  if (!(dp == dp)) { st[CGI_REASON] = -9; st[CGI_DONE] = 1; }               // KSP_DIVERGED_NANORINF
  else if (dp < sc[CG_TTOL]) { st[CGI_REASON] = (dp < sc[CG_ATOL]) ? 3 : 2; st[CGI_DONE] = 1; }
  else if (st[CGI_ITS] >= st[CGI_MAXIT]) { st[CGI_REASON] = -3; st[CGI_DONE] = 1; }   // KSP_DIVERGED_ITS
}
__device__ __forceinline__ void cg_scalar_begin(double *sc, int *st)
{
  const double dp = sqrt(sc[CG_ZZ]);
  sc[CG_DP] = dp; sc[CG_RNORM0] = dp; sc[CG_BETA] = sc[CG_ZR];
  sc[CG_TTOL] = fmax(sc[CG_RTOL] * dp, sc[CG_ATOL]);
  st[CGI_ITS] = 0;
  cg_scalar_converged(dp, sc, st);
}
// ... and after each iteration's (z,z), (z,r): dp = ||z||, betaold <- beta <- (z,r), a kept for the
// deferred x update, iteration count, convergence test (KSPConvergedDefault)
__device__ __forceinline__ void cg_scalar_rotate(double *sc, int *st)
{
  const double dp = sqrt(sc[CG_ZZ]);
  sc[CG_A]       = sc[CG_BETA] / sc[CG_DPI];
  sc[CG_DP]      = dp;
  sc[CG_BETAOLD] = sc[CG_BETA];
  sc[CG_BETA]    = sc[CG_ZR];
  st[CGI_ITS] += 1;
  cg_scalar_converged(dp, sc, st);
}
__device__ __forceinline__ void cg_scalar_post(int post, double *sc, int *st)
{
  if (post == CG_POST_BEGIN) cg_scalar_begin(sc, st);
  else if (post == CG_POST_ROTATE) cg_scalar_rotate(sc, st);
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// programmatic dependent launch (see B200_LAUNCH_PDL): let the next kernel of the stream start its
// prologue; wait until everything earlier kernels wrote is visible
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// the push role: gather x[send_idx] into the peer's lvec over NVLink; the last CTA of a peer
// releases that peer's flag
// NT > 0: only the first NT threads of the CTA take part (the consumer threads of k_stream; they
// meet on named barrier 1); NT == 0: the whole CTA.
// Two steps, so that the fused kernel can put a tile of useful work between them:
//   stores  every thread writes its share of the boundary values into the peer's lvec (posted NVLink
//           writes) and carries on;
//   signal  the CTA's threads meet on a barrier (their stores happen-before thread 0 from here on),
//           thread 0 alone runs the system-scope fence -- cumulative: it orders everything that
//           happened before it, the pattern of a cooperative grid barrier -- and counts the block in;
//           the last block of a peer releases that peer's flag.  By then the write acknowledgements
//           have long arrived, so the fence does not stall the way it does right behind the stores.
template <int NT>
__device__ __forceinline__ void halo_push_stores(const HaloArgs &h, const double *__restrict__ x, int block)
{
  const PushBlock b   = h.blocks[block];
  const PushPeer  p   = h.peers[b.peer_slot];
  double         *dst = (h.seq & 1) ? p.dst[1] : p.dst[0];
  const int       nt  = NT > 0 ? NT : (int)blockDim.x;
  // eight elements per thread and round: the index loads, then the gathers, then the stores of a round
  // are in flight together (a push block of the fused launch carries a few thousand elements)
  constexpr int R = 8;
#pragma unroll 1
  for (int t0 = threadIdx.x; t0 < b.count; t0 += R * nt) {
    int    idx[R];
    double v[R];
#pragma unroll
    for (int j = 0; j < R; ++j) idx[j] = (t0 + j * nt < b.count) ? __ldg(h.send_idx + b.start + t0 + j * nt) : 0;
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = __ldg(x + idx[j]);
#pragma unroll
    for (int j = 0; j < R; ++j)
      if (t0 + j * nt < b.count) dst[b.start + t0 + j * nt] = v[j];
  }
}
template <int NT>
__device__ __forceinline__ void halo_push_signal(const HaloArgs &h, int block)
{
  if (NT > 0) asm volatile("bar.sync 1, %0;" ::"n"(NT > 0 ? NT : 32) : "memory");
  else __syncthreads();
  if (threadIdx.x == 0) {
    const int peer_slot = h.blocks[block].peer_slot;
    // release / acquire fences at system scope, not __threadfence_system(): that one is fence.sc and
    // ptxas adds an invalidation of the whole L1 (CCTL.IVALL) to it, which the x gathers of the five
    // CTAs on this SM then pay for
    fence_acq_rel_sys();
    const unsigned prev = atomicAdd(h.done + peer_slot, 1u);
    if (prev == (unsigned)h.peers[peer_slot].nblocks - 1) {
      h.done[peer_slot] = 0;
      fence_acq_rel_sys();
      st_release_sys(h.peers[peer_slot].flag, h.seq);
    }
  }
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p)
{
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Wait until every source rank's flag is >= seq, bounded by the spin budget.  RELAXED system-scope
// loads and NO system fence: the ghost values are read with ld.global.cg (L2 only, never L1), the
// peer's stores reach this GPU's L2 in release order (data before flag), so once the flag is seen
// in L2 the data is there too.  An acquire at system scope would make ptxas emit an L1 invalidation
// (CCTL.IVALL) and, measured on B200, a system-scope fence executed by every resident CTA costs
// tens of microseconds per MatMult; keep such fences to the release side, one per SM.
__device__ __forceinline__ void halo_wait_flags(const HaloArgs &h)
{
  const unsigned long long t0 = globaltimer_ns();
  for (int s = 0; s < h.nsrc; ++s) {
    const unsigned long long *f = h.flags + h.srcs[s];
    while (ld_relaxed_sys(f) < h.seq) {
      if (globaltimer_ns() - t0 > h.timeout_ns) { atomicExch(h.err, 1ull); break; }
      __nanosleep(100);
    }
  }
}
#endif

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// mbarrier + bulk-copy (TMA engine, 1-D) helpers.  SASS: SYNCS.* and UBLKCP.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  uint32_t addr = smem_u32(bar), done, spins = 0;
  unsigned long long t0 = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    // a barrier that never completes is a bug in the byte accounting: trap (the launch fails with
    // an error the host reports) instead of spinning until someone resets the GPU
    if (!done && (++spins & 0x3FFFu) == 0) {
      const unsigned long long t = globaltimer_ns();
      if (!t0) t0 = t;
      else if (t - t0 > 4000000000ull) __trap();
    }
  } while (!done);
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar, uint64_t policy)
{
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
// scalar loads with an explicit L2 eviction policy: the matrix stream is read once (evict first),
// the gathered vector is reused (evict last) -- keeps x resident in the 126 MB L2 under the stream
__device__ __forceinline__ double ldg_f64_policy(const double *p, uint64_t pol)
{
  double v;
  asm("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ double ldg_f64_stream_policy(const double *p, uint64_t pol)
{
  double v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ int ldg_s32_stream_policy(const int *p, uint64_t pol)
{
  int v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
// streaming (read-once) vector loads that do not allocate in L1
__device__ __forceinline__ double2 ldg_stream_f64x2(const double *p)
{
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
               : "=d"(v.x), "=d"(v.y)
               : "l"(p));
  return v;
}
__device__ __forceinline__ int4 ldg_stream_s32x4(const int *p)
{
  int4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
#endif  // __CUDACC__

}  // namespace b200
