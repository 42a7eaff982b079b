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
int  sm_count();
int  env_int(const char *name, int dflt);

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
// the CG body shared by b200_cg_jacobi and b200_mpiaij_cg_jacobi (b200_vec.cu)
struct CgOps {
  int            m = 0;                       // local rows
  const int32_t *ai = nullptr, *aj = nullptr; // diagonal block (for PCJACOBI)
  const double  *aa = nullptr;
  std::function<int(const double *, double *, cudaStream_t)> mult;       // w = A p
  std::function<int(double *, int, cudaStream_t)>            allreduce;  // in place, device scalars; empty on one GPU
};
int cg_jacobi_run(const CgOps &ops, const double *d_b, double *d_x, double rtol, double atol,
                  int32_t max_it, b200_cg_result_t *res, cudaStream_t st);

// internal (not part of the C ABI): the stream plan of a matrix and the fused launch
int stream_plan_tiles(b200_csr_t A, int4 **d_tiles, int *ntiles, int *grid, int *threads);   // 0 tiles = not applicable
int launch_stream_halo(b200_csr_t A, const double *x, double *y, int mode, const HaloArgs &h,
                       cudaStream_t st);

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
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
// the push role: gather x[send_idx] into the peer's lvec over NVLink; the last CTA of a peer
// releases that peer's flag
__device__ __forceinline__ void halo_push_block(const HaloArgs &h, const double *__restrict__ x, int block)
{
  const PushBlock b   = h.blocks[block];
  const PushPeer  p   = h.peers[b.peer_slot];
  double         *dst = p.dst[h.seq & 1];
  for (int t = threadIdx.x; t < b.count; t += blockDim.x) {
    const int e = b.start + t;
    dst[e]      = __ldg(x + h.send_idx[e]);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(h.done + b.peer_slot, 1u);
    if (prev == (unsigned)p.nblocks - 1) {
      h.done[b.peer_slot] = 0;
      __threadfence_system();
      st_release_sys(p.flag, h.seq);
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
