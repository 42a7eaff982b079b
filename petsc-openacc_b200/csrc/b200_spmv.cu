// b200_spmv.cu -- SeqAIJ sparse mat-vec kernel family for sm_100a and the C ABI around it.
//
// What it replaces: the single OpenACC compute region of the reference,
//   src/openacc-step1/MatMult_SeqAIJ.patch:19-32  (K1)  ... step4 :55-88 (K4a/K4b),
// i.e. PETSc 3.7.6 MatMult_SeqAIJ's row loop, plus MatMultAdd / MatMultTranspose[Add] which the
// reference leaves on the CPU.  Not a port: the matrix streams through shared memory with the
// TMA engine's bulk copies (cp.async.bulk + mbarrier ring) in a persistent, warp-specialised
// kernel, each row is then summed by one thread strictly left to right out of shared memory,
// which makes the fast path of short-row matrices bit-reproducible against the CPU loop.
//
// Kernels (all fp64 values, int32 indices):
//   k_stream  bulk-copy staged CSR tiles, thread per row -- default for short, regular rows; template
//             flags: EPI (MatMult / MatMultAdd / residual / Jacobi sweep / MatMult + CG's (p, A p)), HALO
//             (MatMult_MPIAIJ in one launch: NVLink push + A x + ghost rows), IDX8 (1-byte diagonal codes
//             instead of int32 columns); launched with programmatic stream serialization: the set-up
//             and the first bulk copies of the constant matrix run under the previous kernel's tail
//   k_wmerge  warp-granular chunks of <= 128 non-zeros, products staged coalesced, rows summed in
//             CSR order, persistent warps drawing work from a counter -- skewed row lengths (power law),
//             EXACT and FAST
//   k_merge   nnz-balanced tiles + deterministic carry fix-up, lanes share a row -- B200_MERGE_SPLIT=1 only
//   k_row     thread per row from global memory          -- the reference's launch shape; EXACT fallback
//   k_vector  LANES lanes per row + __shfl_xor reduction -- override / comparison
//   k_cprow   compressed-row (only non-empty rows)       -- matrices with >= 60 % empty rows
//   k_tr_atomic  A^T x by fp64 atomics                   -- transpose without a second copy
//   k_sell    SELL-32-sigma copy (optional, b200_csr_build_sell), column-major chunks of 32 rows,
//             padding skipped -- override only: measured 0.456 ms against k_stream's 0.332 ms on the
//             7-point 300^3 matrix and 0.431 against 0.356 ms on the 27-point 200^3 matrix
//             (profiles/r02_sweep_*.log)
// Why the SELL-C-sigma / sliced-ELL copy is not the default: staging the CSR tile in shared memory
// already gives what those formats buy on a GPU -- the matrix is read in fully coalesced
// 16-byte-aligned bulk copies and lane = consecutive row makes the x gathers of a stencil
// contiguous -- without padding and without a second copy of the values.
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "b200_common.h"

namespace b200 {

thread_local std::string g_last_error;
static thread_local int  g_last_code = 0;
std::atomic<uint64_t>    g_launches{0};

int set_error(int code, const char *fmt, ...)
{
  char    buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  g_last_code  = code;
  return code;
}

int env_int(const char *name, int dflt)
{
  const char *s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

NvtxRange::NvtxRange(const char *name)
{
  static const int enabled = env_int("B200_NVTX", 0);
  on = enabled != 0;
  if (on) nvtxRangePushA(name);
}

static DeviceState g_devs[64];
static std::mutex  g_dev_mutex;

DeviceState *device_state()
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error(B200_ERR_NO_DEVICE, "no CUDA device visible: the b200 library has no CPU fallback");
    return nullptr;
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    set_error(B200_ERR_CUDA, "cudaGetDevice failed or device ordinal out of range");
    return nullptr;
  }
  DeviceState *d = &g_devs[dev];
  if (!d->checked) {
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    if (!d->checked) {
      cudaDeviceProp p;
      if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
        set_error(B200_ERR_CUDA, "cudaGetDeviceProperties(%d) failed", dev);
        return nullptr;
      }
      d->ordinal  = dev;
      d->sm_count = p.multiProcessorCount;
      d->ok       = (p.major == 10);
      if (!d->ok) d->sm_count = p.major * 10 + p.minor;   // kept for the message below
      d->checked  = true;
    }
  }
  if (!d->ok) {
    set_error(B200_ERR_NO_DEVICE, "device %d is sm_%d; this library is built for sm_100a only", dev, d->sm_count);
    return nullptr;
  }
  return d;
}

int ensure_device() { return device_state() ? B200_OK : g_last_code; }
int sm_count()
{
  DeviceState *d = device_state();
  return d ? d->sm_count : 0;
}
bool pdl_enabled()
{
  static const bool on = env_int("B200_PDL", 1) != 0;
  return on;
}

}  // namespace b200

using namespace b200;

// =============================================================================================
// device code
// =============================================================================================
template <int MODE>
__device__ __forceinline__ double acc(double s, double a, double x)
{
  // EXACT: product and sum rounded separately, like the oracle built with -ffp-contract=off.
  if (MODE == B200_MODE_EXACT) return __dadd_rn(s, __dmul_rn(a, x));
  return __fma_rn(a, x, s);
}

// ---------------------------------------------------------------------------------------------
// k_row: one thread per row, left to right, straight from global memory.
// Same launch shape as the reference's `gang vector(32)` loop
// (src/openacc-step1/MatMult_SeqAIJ.patch:19-32); kept as the simplest deterministic kernel.
// ---------------------------------------------------------------------------------------------
template <int MODE, bool ADD>
__global__ void __launch_bounds__(128) k_row(int m, const int *__restrict__ ii,
                                             const int *__restrict__ aj,
                                             const double *__restrict__ aa,
                                             const double *__restrict__ x, const double *yin,
                                             double *y)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  int    lo = ii[i], hi = ii[i + 1];
  double sum = ADD ? yin[i] : 0.0;
  for (int k = lo; k < hi; ++k) sum = acc<MODE>(sum, aa[k], __ldg(x + aj[k]));
  y[i] = sum;
}

// epilogue over a finished product t = A x held in y (plans without a fused epilogue)
template <int EPI>
__global__ void k_epilogue(int m, const double *__restrict__ x, const double *__restrict__ b,
                           const double *__restrict__ dinv, double *y)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double r = __dsub_rn(b[i], y[i]);
  y[i] = (EPI == 2) ? r : __dadd_rn(x[i], __dmul_rn(dinv[i], r));
}

// ---------------------------------------------------------------------------------------------
// k_vector: LANES lanes per row, strided partial sums, butterfly reduction.
// ---------------------------------------------------------------------------------------------
template <int LANES, bool ADD>
__global__ void __launch_bounds__(256) k_vector(int m, const int *__restrict__ ii,
                                                const int *__restrict__ aj,
                                                const double *__restrict__ aa,
                                                const double *__restrict__ x, const double *yin,
                                                double *y)
{
  long long t    = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int       row  = (int)(t / LANES);
  int       lane = (int)(t % LANES);
  double    sum  = 0.0;
  int       lo = 0, hi = 0;
  if (row < m) { lo = __ldg(ii + row); hi = __ldg(ii + row + 1); }
  for (int k = lo + lane; k < hi; k += LANES) sum = __fma_rn(aa[k], __ldg(x + aj[k]), sum);
#pragma unroll
  for (int off = LANES / 2; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off, LANES);
  if (row < m && lane == 0) y[row] = ADD ? yin[row] + sum : sum;
}

// ---------------------------------------------------------------------------------------------
// k_stream: persistent, warp-specialised.  One producer lane drives the TMA engine: for each tile
// it bulk-copies the tile's slice of aa, aj and ii into a ring of shared-memory stages
// (cp.async.bulk ... mbarrier::complete_tx, evict-first in L2: the matrix is read exactly once).
// THREADS consumer threads own one row each and sum it left to right out of shared memory;
// x is gathered through the read-only path (L1/L2 resident window for banded matrices, and for
// stencil matrices lane = consecutive rows makes every gather a contiguous 256-byte request).
// tiles[t] = {first row, last row + 1, ii[first], ii[last + 1]}.
// ---------------------------------------------------------------------------------------------
#define STREAM_PAD 16  // alignment slack (<= 3 + 3) + predicated read-ahead (<= 7)
#define STREAM_U 8

// IDX8 (compressed column indices): when a matrix has at most 256 distinct diagonals (col - row),
// as every stencil matrix on a structured grid has, the kernel streams one byte per non-zero (a
// code into a table of offsets held in shared memory) instead of the 4-byte column index.  Same
// values, same order, same bits in y; 9 instead of 12 bytes of DRAM traffic per non-zero.
#define STREAM_PAD8 48  // byte codes: 16-byte alignment slack on both ends + read-ahead
struct Idx8Args {
  const unsigned char *aj8;   // nz (+ pad) codes
  const int           *offs;  // 256 entries: col = row + offs[code]
  // RL8 (compressed row pointers, rows of at most 255 entries): one byte of row LENGTH per row and,
  // per tile and warp, the offset of the warp's first row; the lanes rebuild their row's range with a
  // warp scan.  1.125 instead of 4 bytes per row of DRAM traffic; same values, same order, same bits.
  const unsigned char *rl8;   // m (+ pad) row lengths
  const int           *wbase; // ntiles x (THREADS / 32): ai[first row of the warp]
};

__host__ __device__ inline size_t stream_aj_bytes(int cap, bool idx8)
{
  return idx8 ? (size_t)((cap + STREAM_PAD8 + 15) & ~15) : (size_t)(cap + STREAM_PAD) * 4;
}
__host__ __device__ inline size_t stream_stage_bytes(int threads, int cap, bool idx8 = false)
{
  return (size_t)(cap + STREAM_PAD) * 8 + stream_aj_bytes(cap, idx8) + (size_t)(threads + 8) * 4;
}
__host__ __device__ inline size_t stream_header_bytes(bool idx8) { return idx8 ? 128 + 1024 : 128; }

// EPI: what happens to the row sum t_i = sum_k a_ik x_k before it is stored --
//   0  y_i = t_i                          MatMult
//   1  y_i = yin_i + ... (the accumulator STARTS at yin_i)   MatMultAdd
//   2  y_i = b_i - t_i                    the residual of KSPRichardson / KSPInitialResidual
//   3  y_i = x_i + dinv_i * (b_i - t_i)   one Richardson(1) + PCJACOBI sweep of the GAMG levels
//      (configs/PETSc_SolverOptions_GAMG.info:15-21), i.e. MatMult, VecAYPX, VecPointwiseMult and
//      VecAXPY in one pass; every step is rounded separately, as the four PETSc calls round.
//   4  y_i = t_i and the CTA accumulates x_i * t_i: CG's (p, A p) without a second pass over p and w
enum { EPI_NONE = 0, EPI_ADD = 1, EPI_RESIDUAL = 2, EPI_JACOBI = 3, EPI_DOT = 4 };

template <int MODE, int EPI, int THREADS, bool HALO, bool IDX8, bool RL8 = false>
__global__ void __launch_bounds__(THREADS + 32, THREADS == 256 ? 5 : 10)
    k_stream(const int4 *__restrict__ tiles_in, int ntiles, const int *__restrict__ ii,
             const int *__restrict__ aj, const double *__restrict__ aa,
             const double *__restrict__ x, const double *yin, double *y, int cap, int stages,
             const HaloArgs h, const Idx8Args ix, const double *__restrict__ aux, const DotArgs dot)
{
  // HALO = MatMult_MPIAIJ in one launch:
  //  * VecScatterBegin: the consumers of the first h.npush CTAs (at most one per SM) start by storing a
  //    block of this rank's boundary values into a peer's lvec over NVLink and, once a peer's last block
  //    is out, releasing that peer's flag; then they stream tiles like every other CTA;
  //  * VecScatterEnd + MatMultAdd(B, lvec, y, y): after its last tile a CTA waits for the flags of
  //    this rank's sources (the ghosts arrived long ago: the wait is off the critical path) and
  //    continues the rows of its own tiles that touch a ghost -- A terms first, then B terms, the
  //    order of MatMult followed by MatMultAdd.
  // Measured alternatives (profiles/r01_fused_halo_notes.md): folding the ghost terms into the tile
  // loop behind a per-tile bit mask, a separate communication warp, spreading the push over every
  // CTA -- all slower than this arrangement.
  //
  // Programmatic dependent launch: the matrix arrays (tiles, ii, aj/codes, aa, the diagonal table) are
  // constant for the life of the handle, so the shared-memory set-up and the producer's first bulk
  // copies run BEFORE griddepcontrol.wait, i.e. under the tail of the previous kernel of the stream;
  // only the consumers (x, yin, y, the halo) wait for it.
  pdl_launch_dependents();
  const int bid = (int)blockIdx.x;
  const int nb  = (int)gridDim.x;
  const int4 *__restrict__ tiles = tiles_in;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full  = reinterpret_cast<uint64_t *>(smem);
  uint64_t *empty = full + stages;
  int           *soffs  = reinterpret_cast<int *>(smem + 128);  // IDX8: the diagonal table
  unsigned char *stage0 = smem + stream_header_bytes(IDX8);     // barriers live in the first 128 bytes
  const size_t   sbytes = stream_stage_bytes(THREADS, cap, IDX8);
  const size_t   aabytes = (size_t)(cap + STREAM_PAD) * 8, ajbytes = stream_aj_bytes(cap, IDX8);

  const int tid = threadIdx.x;
  // this CTA's tiles: every grid-th one, or (HALO, always) its range of the launch's own schedule --
  // decided at compile time: the HALO instantiation sits at the register cap of 5 CTAs per SM, and
  // two more live values in the tile loop made ptxas issue the eight gathers of a row in two groups
  // of four (3 % slower)
  if (HALO) tiles = h.sched_tiles;
  const int tile0 = HALO ? __ldg(h.sched_first + bid) : bid;
  const int tile1 = HALO ? __ldg(h.sched_first + bid + 1) : ntiles;
  const int tstep = HALO ? 1 : nb;
  if (IDX8) {
    // all 256 table entries, also when the CTA has fewer than 256 threads (THREADS = 128 -> 160)
    for (int t = tid; t < 256; t += THREADS + 32) soffs[t] = __ldg(ix.offs + t);
  }
  int *sskip = reinterpret_cast<int *>(smem + 120);             // EPI_DOT: "this launch is a no-op"
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], THREADS / 32);
    }
    mbar_fence_init();
    // The skip word is read BEFORE griddepcontrol.wait, once per CTA.  Every CTA of the launch reads
    // the same value because the word is only written by a scalar step at least two kernels back in
    // the stream and the kernel in between triggers its dependents after its own wait (see
    // cg_jacobi_run): that scalar step has completed before any CTA of this launch starts.
    if (EPI == EPI_DOT) *sskip = (dot.skip != nullptr) ? *reinterpret_cast<const volatile int *>(dot.skip) : 0;
  }
  __syncthreads();
  if (EPI == EPI_DOT && *sskip) return;

  if (tid >= THREADS) {
    // ------------------------------- producer warp -------------------------------------------
    if (tid == THREADS) {
      const uint64_t pol = l2_policy_evict_first();
      int it = 0;
      for (int tile = tile0; tile < tile1; tile += tstep, ++it) {
        const int s = it % stages;
        if (it >= stages) mbar_wait(&empty[s], ((it / stages) - 1) & 1);
        const int4 d   = __ldg(tiles + tile);
        const int  s4  = d.z & ~3;
        const int  nn  = ((d.w + 3) & ~3) - s4;
        const int  r0a = d.x & ~3;
        const int  nr  = ((d.y + 1 + 3) & ~3) - r0a;
        unsigned char *st  = stage0 + (size_t)s * sbytes;
        double        *saa = reinterpret_cast<double *>(st);
        unsigned char *saj = st + aabytes;
        int           *sii = reinterpret_cast<int *>(st + aabytes + ajbytes);
        const int      s16 = d.z & ~15;
        const int      nn8 = ((d.w + 15) & ~15) - s16;   // IDX8: bytes of codes
        // expect exactly the bytes issued below: a tile of empty rows (nn == 0) copies no codes even
        // when the 16-byte rounding of the code range (nn8) is not empty
        const int      r16 = d.x & ~15;
        const int      nrb = ((d.y + 15) & ~15) - r16;   // RL8: bytes of row lengths
        mbar_expect_tx(&full[s], (uint32_t)((nn > 0 ? nn * 8 + (IDX8 ? nn8 : nn * 4) : 0) + (RL8 ? nrb + (THREADS / 32) * 4 : nr * 4)));
        if (nn > 0) {
          bulk_g2s(saa, aa + s4, (uint32_t)nn * 8, &full[s], pol);
          if (IDX8) bulk_g2s(saj, ix.aj8 + s16, (uint32_t)nn8, &full[s], pol);
          else bulk_g2s(saj, aj + s4, (uint32_t)nn * 4, &full[s], pol);
        }
        if (RL8) {
          // [0, 32): the warps' first offsets; [32, ...): the row lengths from the 16-byte boundary below d.x
          bulk_g2s(sii, ix.wbase + (size_t)tile * (THREADS / 32), (uint32_t)(THREADS / 32) * 4, &full[s], pol);
          bulk_g2s(reinterpret_cast<unsigned char *>(sii) + 32, ix.rl8 + r16, (uint32_t)nrb, &full[s], pol);
        } else {
          bulk_g2s(sii, ii + r0a, (uint32_t)nr * 4, &full[s], pol);
        }
      }
    } else if (HALO) {
      // The other 31 lanes of the producer warp have nothing to do: they walk this CTA's ghost-row
      // list once and pull what the closing phase will read (row list, B's row pointers, row numbers,
      // column positions, values -- all constant) into L2, so that the dependent loads at the very end
      // of the CTA, which sit on the critical path of the launch, are L2 hits instead of DRAM round trips.
      const int q0 = __ldg(h.cta_ptr + bid), q1 = __ldg(h.cta_ptr + bid + 1);
#pragma unroll 1
      for (int q = q0 + (tid - THREADS - 1); q < q1; q += 31) {
        const int c  = __ldg(h.cta_rows + q);
        const int lo = __ldg(h.cpi + c);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(h.ridx + c));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(h.bj + lo));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(h.ba + lo));
      }
    }
    return;
  }

  // --------------------------------- consumers -------------------------------------------------
  pdl_wait();
  // VecScatterBegin: the stores go out now; the fence + flag release follow a tile or two later, when
  // the write acknowledgements are back and the fence is cheap (the peers need the flag at THEIR end)
  if (HALO && bid < h.npush) halo_push_stores<THREADS>(h, x, bid);
  // iteration after which a push CTA signals (-1: this CTA pushes nothing)
  const int signal_at = (HALO && bid < h.npush) ? ((tile1 - tile0 > tstep) ? 1 : 0) : -1;
  double dacc = 0.0;
  int it = 0;
  for (int tile = tile0; tile < tile1; tile += tstep, ++it) {
    const int  s = it % stages;
    const int4 d = __ldg(tiles + tile);
    unsigned char *st  = stage0 + (size_t)s * sbytes;
    const double  *saa = reinterpret_cast<const double *>(st);
    const int     *saj = reinterpret_cast<const int *>(st + aabytes);
    const unsigned char *saj8 = st + aabytes;
    const int     *sii = reinterpret_cast<const int *>(st + aabytes + ajbytes) + (d.x & 3);
    const int      r   = d.x + tid;
    double sum = 0.0;
    double bi = 0.0;
    if (EPI == EPI_ADD) { if (r < d.y) sum = yin[r]; }
    if (EPI == EPI_RESIDUAL || EPI == EPI_JACOBI) { if (r < d.y) bi = yin[r]; }
    mbar_wait(&full[s], (it / stages) & 1);
    int lo, hi;
    if (RL8) {
      const unsigned char *srl = reinterpret_cast<const unsigned char *>(sii - (d.x & 3)) + 32 + (d.x & 15);
      const int len = (r < d.y) ? (int)srl[tid] : 0;
      int incl = len;   // inclusive scan of the row lengths over the warp
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if ((tid & 31) >= off) incl += t;
      }
      hi = (sii - (d.x & 3))[tid >> 5] + incl;
      lo = hi - len;
    } else {
      lo = sii[tid];
      hi = sii[tid + 1];
    }
    if (r < d.y) {
      const int p  = lo - (d.z & ~3);
      const int p8 = lo - (d.z & ~15);
      const int n  = hi - lo;
      for (int k = 0; k < n; k += STREAM_U) {
        double av[STREAM_U], xv[STREAM_U];
#pragma unroll
        for (int j = 0; j < STREAM_U; ++j) {
          const bool ok = (k + j) < n;
          av[j] = saa[p + k + j];
          const int c = IDX8 ? r + soffs[saj8[p8 + k + j]] : saj[p + k + j];
          xv[j] = __ldg(x + (ok ? c : 0));   // unconditional load (x[0] past the row end): keeps the 8 gathers batched
        }
#pragma unroll
        for (int j = 0; j < STREAM_U; ++j)
          if ((k + j) < n) sum = acc<MODE>(sum, av[j], xv[j]);
      }
      if (EPI == EPI_RESIDUAL) sum = __dsub_rn(bi, sum);
      if (EPI == EPI_JACOBI) sum = __dadd_rn(__ldg(x + r), __dmul_rn(__ldg(aux + r), __dsub_rn(bi, sum)));
      if (EPI == EPI_DOT) dacc = __fma_rn(__ldg(x + r), sum, dacc);
      y[r] = sum;
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    if (HALO && it == signal_at) halo_push_signal<THREADS>(h, bid);
  }
  if (HALO) {
    // one lane per source rank waits for that rank's flag; the consumer-only barrier publishes the
    // flags and this CTA's y stores; cta_rows lists the ghost-touching rows of this CTA's tiles
    if (tid < h.nsrc) {
      const unsigned long long t0 = globaltimer_ns();
      const unsigned long long *f = h.flags + h.srcs[tid];
#pragma unroll 1
      while (ld_relaxed_sys(f) < h.seq) {
        if (globaltimer_ns() - t0 > h.timeout_ns) { atomicExch(h.err, 1ull); break; }
        __nanosleep(32);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
    // (the closing loops are kept rolled: the kernel's code must stay well inside the instruction cache)
#pragma unroll 1
    for (int q = h.cta_ptr[bid] + tid; q < h.cta_ptr[bid + 1]; q += THREADS) {
      const int c  = h.cta_rows[q];
      const int lo = h.cpi[c], hi = h.cpi[c + 1], i = h.ridx[c];
      const double s0 = y[i];
      double    sb = s0;
#pragma unroll 1
      for (int k = lo; k < hi; ++k) sb = acc<MODE>(sb, h.ba[k], __ldcg(h.lvec + h.bj[k]));
      y[i] = sb;
      if (EPI == EPI_DOT) dacc = __fma_rn(__ldg(x + i), __dsub_rn(sb, s0), dacc);   // the ghost terms' share of (x, y)
    }
  }
  if (EPI == EPI_DOT) {
    // CTA sum in a fixed order (lanes by butterfly, warps in index order), one partial per CTA; the
    // last CTA to arrive adds the partials in index order: the same bits from run to run
    double *red = reinterpret_cast<double *>(stage0);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, off);
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");   // every consumer is done with the stages
    if ((tid & 31) == 0) red[tid >> 5] = dacc;
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
    if (tid < 32) {
      int last = 0;
      if (tid == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) t += red[w];
        dot.partials[bid] = t;
        __threadfence();
        last = (atomicAdd(dot.counter, 1u) == (unsigned)nb - 1);
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last) {
        __threadfence();
        double t = 0.0;
#pragma unroll 1
        for (int q = tid; q < nb; q += 32) t += __ldcg(dot.partials + q);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        if (tid == 0) { *dot.out = t; *dot.counter = 0u; }
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------
// k_merge: nnz-balanced tiles for skewed row lengths (power-law matrices).  A tile is a run of at
// most MERGE_CAP consecutive non-zeros and at most MERGE_RCAP rows; long rows are split across
// tiles.  Phase 1: the CTA forms all products of its run, coalesced over non-zeros, into shared
// memory.  Phase 2: a segmented reduction with L = 1..32 lanes per row (L chosen per tile from its
// row count, so a tile holding one 2048-entry row and a tile holding 600 short rows are both
// busy).  Rows completed inside the tile are stored; the head / tail partial of a split row goes
// to head[tile] / tail[tile] and k_merge_fixup adds them in tile order -- no atomics, so the
// result is the same from run to run.
// tiles[t] = {first row, last row + 1, first nnz, last nnz + 1}.
// ---------------------------------------------------------------------------------------------
#define MERGE_CAP 2048
#define MERGE_RCAP 1024
#define MERGE_THREADS 256

template <bool ADD>
__global__ void __launch_bounds__(MERGE_THREADS)
    k_merge(const int4 *__restrict__ tiles, const int *__restrict__ ii, const int *__restrict__ aj,
            const double *__restrict__ aa, const double *__restrict__ x, const double *yin,
            double *y, double *head, double *tail)
{
  __shared__ double prod[MERGE_CAP];
  __shared__ int    rp[MERGE_RCAP + 1];
  const int4 d   = __ldg(tiles + blockIdx.x);
  const int  tid = threadIdx.x, nr = d.y - d.x, s = d.z, e = d.w;
  for (int j = tid; j <= nr; j += MERGE_THREADS) rp[j] = min(max(__ldg(ii + d.x + j), s), e) - s;
  const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
  {
    // 8 non-zeros per thread, all index loads first, then all gathers: 8 independent requests in
    // flight per thread (the gather is the latency-bound part)
    constexpr int PER = MERGE_CAP / MERGE_THREADS;
    const int     n   = e - s;
    int           c[PER];
    double        a[PER], xv[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int k = tid + i * MERGE_THREADS;
      c[i] = (k < n) ? ldg_s32_stream_policy(aj + s + k, pol_stream) : 0;
      a[i] = (k < n) ? ldg_f64_stream_policy(aa + s + k, pol_stream) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) xv[i] = (tid + i * MERGE_THREADS < n) ? ldg_f64_policy(x + c[i], pol_keep) : 0.0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int k = tid + i * MERGE_THREADS;
      if (k < n) prod[k] = a[i] * xv[i];
    }
  }
  __syncthreads();
  int L = 1;
  while (L < 32 && L * nr * 2 <= MERGE_THREADS) L <<= 1;   // lanes per row, uniform over the CTA
  const int lane = tid & (L - 1), slot = tid / L, slots = MERGE_THREADS / L;
  for (int jb = 0; jb < nr; jb += slots) {
    const int j   = jb + slot;
    double    sum = 0.0;
    if (j < nr)
      for (int k = rp[j] + lane; k < rp[j + 1]; k += L) sum += prod[k];
    for (int off = L >> 1; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if (j < nr && lane == 0) {
      const int  r         = d.x + j;
      const bool head_part = (j == 0) && (__ldg(ii + r) < s);
      const bool tail_part = (j == nr - 1) && (__ldg(ii + r + 1) > e);
      if (!head_part && !tail_part) y[r] = ADD ? yin[r] + sum : sum;
      else if (head_part) head[blockIdx.x] = sum;   // also the "whole tile inside one row" case
      else tail[blockIdx.x] = sum;
    }
  }
}

// split[q] = {row, first tile, last tile}: y[row] = tail[first] + head[first+1..last]
template <bool ADD>
__global__ void k_merge_fixup(int nsplit, const int4 *__restrict__ split, const double *__restrict__ head,
                              const double *__restrict__ tail, const double *yin, double *y)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nsplit) return;
  const int4 sp  = split[q];
  double     sum = tail[sp.y];
  for (int t = sp.y + 1; t <= sp.z; ++t) sum += head[t];
  y[sp.x] = ADD ? yin[sp.x] + sum : sum;
}

// ---------------------------------------------------------------------------------------------
// k_wmerge: the exact summation order on skewed matrices, warp granular.  The matrix is cut into
// CHUNKS of at most WM_CAP consecutive non-zeros: either a run of whole rows (each at most WM_CAP
// long) or one piece of a longer row.  A warp stages a chunk -- column indices and values coalesced
// over the non-zeros, four per lane, all gathers of x in flight together -- as rounded products (or
// a and x side by side for the fused chain) in its own slice of shared memory, then its lanes sum
// the chunk's rows left to right, one row per lane; a long row is summed by lane 0 chunk after
// chunk with the partial sum carried in a register.  No block barrier, no tile is ever cut short by
// a long row next to it (what left the whole-row CTA tiles of round 1's k_mergex a quarter full on the
// power-law matrix: 1.01 ms there, 0.856 ms here), and persistent warps draw BLOCKS of consecutive
// chunks from an atomic counter, so a 10,000-entry row delays one warp, not a CTA.  At 115 G
// non-zeros/s it runs at the rate this GPU gives to independent 8-byte gathers from an 80 MB vector:
// the stream kernel on the (regular) transpose of the same matrix reaches the same 115, a bare
// torch.index_select over the same index stream 148 (profiles/r02_powerlaw.md).
// chunks[c] = {first row, rows (>= 0) | -1 middle / -2 first / -3 last piece of a long row, k0, k1}.
// ---------------------------------------------------------------------------------------------
#define WM_CAP 128
#define WM_WARPS 8
#define WM_PER (WM_CAP / 32)
template <int MODE, bool ADD>
__global__ void __launch_bounds__(WM_WARPS * 32)
    k_wmerge(const int4 *__restrict__ chunks, const int *__restrict__ blk, int nblk, unsigned *counters,
             const int *__restrict__ ii, const int *__restrict__ aj, const double *__restrict__ aa,
             const double *__restrict__ x, const double *yin, double *y)
{
  constexpr bool PROD = (MODE == B200_MODE_EXACT);
  __shared__ double sp[WM_WARPS][WM_CAP];
  __shared__ double sx[WM_WARPS][PROD ? 1 : WM_CAP];
  __shared__ int    srp[WM_WARPS][WM_CAP + 1];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
  double carry = 0.0;   // lane 0: the running sum of a long row
  for (;;) {
    int b = 0;
    if (lane == 0) b = (int)atomicAdd(counters, 1u);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= nblk) break;
    const int c0 = __ldg(blk + b), c1 = __ldg(blk + b + 1);
    for (int c = c0; c < c1; ++c) {
      const int4 d = __ldg(chunks + c);
      const int  n = d.w - d.z;
      int    cj[WM_PER];
      double av[WM_PER], xv[WM_PER];
#pragma unroll
      for (int j = 0; j < WM_PER; ++j) {
        const int k = lane + 32 * j;
        cj[j] = (k < n) ? ldg_s32_stream_policy(aj + d.z + k, pol_stream) : 0;
        av[j] = (k < n) ? ldg_f64_stream_policy(aa + d.z + k, pol_stream) : 0.0;
      }
      if (d.y >= 0)
        for (int j = lane; j <= d.y; j += 32) srp[w][j] = __ldg(ii + d.x + j) - d.z;
#pragma unroll
      for (int j = 0; j < WM_PER; ++j) xv[j] = (lane + 32 * j < n) ? ldg_f64_policy(x + cj[j], pol_keep) : 0.0;
#pragma unroll
      for (int j = 0; j < WM_PER; ++j) {
        const int k = lane + 32 * j;
        if (k < n) {
          if (PROD) sp[w][k] = __dmul_rn(av[j], xv[j]);
          else { sp[w][k] = av[j]; sx[w][k] = xv[j]; }
        }
      }
      __syncwarp();
      if (d.y >= 0) {
        for (int r = lane; r < d.y; r += 32) {
          const int row = d.x + r;
          double    sum = ADD ? yin[row] : 0.0;
          const int lo = srp[w][r], hi = srp[w][r + 1];
          if (PROD) for (int k = lo; k < hi; ++k) sum = __dadd_rn(sum, sp[w][k]);
          else for (int k = lo; k < hi; ++k) sum = __fma_rn(sp[w][k], sx[w][k], sum);
          y[row] = sum;
        }
      } else if (lane == 0) {
        if (d.y == -2) carry = ADD ? yin[d.x] : 0.0;
        if (n == WM_CAP) {
#pragma unroll 16
          for (int l = 0; l < WM_CAP; ++l) carry = PROD ? __dadd_rn(carry, sp[w][l]) : __fma_rn(sp[w][l], sx[w][l], carry);
        } else {
          for (int l = 0; l < n; ++l) carry = PROD ? __dadd_rn(carry, sp[w][l]) : __fma_rn(sp[w][l], sx[w][l], carry);
        }
        if (d.y == -3) y[d.x] = carry;
      }
      __syncwarp();
    }
  }
  // the last warp out re-arms the work counter for the next launch
  if (lane == 0) {
    __threadfence();
    const unsigned total = gridDim.x * WM_WARPS;
    if (atomicAdd(counters + 1, 1u) == total - 1) { counters[0] = 0u; counters[1] = 0u; }
  }
}

// ---------------------------------------------------------------------------------------------
// k_cprow: compressed-row MatMultAdd / MatMult [P376 MatMult_SeqAIJ compressed branch]: only the
// cprow.nrows non-empty rows are touched (the off-diagonal block B has ~2% non-empty rows).
// For MatMult (ADD = false) y has been zeroed by the caller.
// ---------------------------------------------------------------------------------------------
template <int MODE, bool ADD>
__global__ void __launch_bounds__(128) k_cprow(int nrows, const int *__restrict__ cpi,
                                               const int *__restrict__ ridx,
                                               const int *__restrict__ aj,
                                               const double *__restrict__ aa,
                                               const double *__restrict__ x, const double *yin,
                                               double *y)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nrows) return;
  int    lo = cpi[t], hi = cpi[t + 1], i = ridx[t];
  double sum = ADD ? yin[i] : 0.0;
  for (int k = lo; k < hi; ++k) sum = acc<MODE>(sum, aa[k], __ldg(x + aj[k]));
  y[i] = sum;
}

// ---------------------------------------------------------------------------------------------
// k_tr_atomic: y[aj[k]] += x[i]*aa[k] with fp64 atomics (RED.E.ADD.F64); y pre-initialised.
// Run-to-run order is not fixed, so this serves MODE_FAST only.
// ---------------------------------------------------------------------------------------------
template <int LANES>
__global__ void __launch_bounds__(256) k_tr_atomic(int m, const int *__restrict__ ii,
                                                   const int *__restrict__ aj,
                                                   const double *__restrict__ aa,
                                                   const double *__restrict__ x, double *y)
{
  long long t    = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int       row  = (int)(t / LANES);
  int       lane = (int)(t % LANES);
  if (row >= m) return;
  const double alpha = x[row];
  const int    lo = ii[row], hi = ii[row + 1];
  for (int k = lo + lane; k < hi; k += LANES) atomicAdd(y + aj[k], alpha * aa[k]);
}

// ---------------------------------------------------------------------------------------------
// k_sell: SELL-32-sigma (sliced ELLPACK, slice height = one warp).  Rows are grouped in chunks of
// 32 (inside windows of sigma rows they are first ordered by decreasing length); a chunk stores
// max-row-length x 32 entries column-major, so lane l's k-th entry sits at base + 32 k + l: every
// load of values and indices is one contiguous 256- / 128- / 32-byte request without any staging.
// Padding is marked (column -1, or byte code 255) and SKIPPED, never multiplied, so a row is still
// summed left to right over exactly its own entries: the bits of the CPU loop.  No row pointers
// are read (4 bytes per chunk instead of per row).
// The optional copy the north star lists for stencil matrices; built by b200_csr_build_sell and
// used only under the B200_KERNEL_SELL override (the stream kernel stays the measured default).
// ---------------------------------------------------------------------------------------------
#define SELL_C 32
#define SELL_U 8
#define SELL_PAD_CODE 255
// `asm volatile` loads: the compiler keeps them in program order (8 values, 8 indices, then 8
// gathers issued back to back) instead of sinking each load next to its use
__device__ __forceinline__ double sell_ld_f64_stream(const double *p, uint64_t pol)
{
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ int sell_ld_s32_stream(const int *p, uint64_t pol)
{
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ int sell_ld_u8_stream(const unsigned char *p, uint64_t pol)
{
  unsigned int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return (int)v;
}
__device__ __forceinline__ double sell_ld_f64(const double *p)
{
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
template <int MODE, bool ADD, bool IDX8, bool PERM>
__global__ void __launch_bounds__(256)
    k_sell(int m, int nchunks, const unsigned int *__restrict__ cs, const double *__restrict__ val,
           const int *__restrict__ col, const unsigned char *__restrict__ code, const int *__restrict__ offs,
           const int *__restrict__ perm, const double *__restrict__ x, const double *yin, double *y)
{
  __shared__ int soffs[256];
  if (IDX8) {
    for (int t = threadIdx.x; t < 256; t += blockDim.x) soffs[t] = __ldg(offs + t);
    __syncthreads();
  }
  const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = (int)(slot >> 5), lane = (int)(slot & 31);
  if (chunk >= nchunks) return;
  const int row = PERM ? __ldg(perm + slot) : (slot < m ? (int)slot : -1);   // -1: padding slot
  const unsigned int base = __ldg(cs + chunk);
  const int len = (int)((__ldg(cs + chunk + 1) - base) >> 5);
  double sum = (ADD && row >= 0) ? yin[row] : 0.0;
  const uint64_t pol = l2_policy_evict_first();
  for (int k = 0; k < len; k += SELL_U) {
    // three passes so that the 8 value loads, the 8 index loads and then the 8 gathers are all in
    // flight together: every load is unconditional (the arrays carry SELL_U * 32 entries of tail
    // padding, a gather of an absent entry reads x[0]) and only the adds are predicated
    double a[SELL_U], xv[SELL_U];
    int    c[SELL_U];
#pragma unroll
    for (int j = 0; j < SELL_U; ++j) {
      const size_t at = (size_t)base + (size_t)(k + j) * SELL_C + lane;
      a[j] = sell_ld_f64_stream(val + at, pol);
      c[j] = IDX8 ? sell_ld_u8_stream(code + at, pol) : sell_ld_s32_stream(col + at, pol);
    }
#pragma unroll
    for (int j = 0; j < SELL_U; ++j) {
      const bool ok = (k + j) < len;
      if (IDX8) c[j] = (ok && c[j] != SELL_PAD_CODE) ? row + soffs[c[j]] : -1;
      else if (!ok) c[j] = -1;
      xv[j] = sell_ld_f64(x + (c[j] >= 0 ? c[j] : 0));
    }
#pragma unroll
    for (int j = 0; j < SELL_U; ++j)
      if (c[j] >= 0) sum = acc<MODE>(sum, a[j], xv[j]);
  }
  if (row >= 0) y[row] = sum;
}

// ---- device-side transpose build (setup, once per matrix) ------------------------------------
// row id of every non-zero (thread per row), column histogram, and the gather through the sorted
// permutation; the stable radix sort by column keeps ascending row order inside each column.
__global__ void k_tr_expand(int m, const int *__restrict__ ii, int *__restrict__ rowid, const int *__restrict__ aj, int *cnt)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  for (int k = ii[i]; k < ii[i + 1]; ++k) { rowid[k] = i; atomicAdd(cnt + aj[k], 1); }
}
__global__ void k_tr_gather(int nz, const int *__restrict__ perm, const int *__restrict__ rowid, const double *__restrict__ aa,
                            int *__restrict__ tj, double *__restrict__ ta)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nz) return;
  const int k = perm[p];
  tj[p] = rowid[k];
  ta[p] = aa[k];
}
__global__ void k_iota(int n, int *a)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}

__global__ void k_copy(double *__restrict__ dst, const double *__restrict__ src, long long n)
{
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long s = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += s) dst[i] = src[i];
}

// =============================================================================================
// host side: the handle, the plan, the C ABI
// =============================================================================================
struct b200_csr_s {
  int32_t m = 0, n = 0, nz = 0;
  int    *d_ai = nullptr, *d_aj = nullptr;
  double *d_aa = nullptr;
  // statistics (MatAssemblyEnd_SeqAIJ bookkeeping)
  int32_t nonzerorowcnt = 0, rmax = 0;
  int32_t hist[16]      = {0};
  // compressed row (MatCheckCompressedRow, ratio 0.6)
  bool    cprow_use = false;
  int32_t cprow_nrows = 0;
  int    *d_cpi = nullptr, *d_ridx = nullptr;
  // stream plan
  int4   *d_tiles = nullptr;
  // compressed column indices (see IDX8)
  bool           idx8 = false;
  unsigned char *d_aj8 = nullptr;
  int           *d_offs = nullptr;
  int32_t        noffs = 0;
  // compressed row pointers (see RL8)
  bool           rl8 = false;
  unsigned char *d_rl8 = nullptr;
  int           *d_wbase = nullptr;
  std::vector<int4> h_tiles;
  int32_t ntiles = 0, stream_threads = 256, stream_cap = 0, stream_stages = 0, stream_grid = 0;
  size_t  stream_smem = 0;
  // merge plan
  int4   *d_mtiles = nullptr, *d_msplit = nullptr;
  double *d_mhead = nullptr, *d_mtail = nullptr;
  int32_t nmtiles = 0, nmsplit = 0;
  // warp-granular exact-order plan (k_wmerge): chunks, blocks of chunks, work counters
  int4     *d_wchunks = nullptr;
  int      *d_wblk = nullptr;
  unsigned *d_wcounters = nullptr;
  int32_t   nwchunks = 0, nwblk = 0;
  // optional SELL-32-sigma copy (b200_csr_build_sell)
  double        *d_sval = nullptr;
  int           *d_scol = nullptr, *d_sperm = nullptr;
  unsigned char *d_scode = nullptr;
  unsigned int  *d_scs = nullptr;
  int32_t        sell_chunks = 0, sell_sigma = 0;
  uint64_t       sell_padded = 0;
  // column blocks of the k_wmerge plan when x does not fit L2 (see build_colblocks): A = [A_0 | A_1 | ...]
  std::vector<b200_csr_s *> colblocks;
  bool is_colblock = false;
  // vector plan
  int32_t vector_lanes = 8;
  int32_t kernel_fast = B200_KERNEL_ROW, kernel_exact = B200_KERNEL_ROW, kernel_override = 0;
  // explicit transpose (n x m), built lazily
  b200_csr_s *T = nullptr;
  // host-vector path scratch
  double      *d_hx = nullptr, *d_hy = nullptr;
  size_t       hx_len = 0, hy_len = 0;
  cudaStream_t hs[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t  hev[2] = {nullptr, nullptr};
  // row-blocked pipeline of the host-vector path (tile-aligned blocks of ~1M rows)
  std::vector<int>         pb_tile;   // block b = tiles [pb_tile[b], pb_tile[b+1])
  std::vector<int>         pb_need;   // last x chunk block b reads (chunks = the same row blocks)
  std::vector<cudaEvent_t> pb_evx, pb_evk;
  uint64_t     device_bytes = 0;
};

static int hist_bucket(int len)
{
  if (len <= 2) return len;  // 0,1,2
  int b = 3, hi = 4;         // [3-4],[5-8],[9-16],...
  while (len > hi && b < 15) { hi <<= 1; ++b; }
  return b;
}

template <typename T>
static int dev_alloc(T **p, size_t count, b200_csr_s *A)
{
  size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  B200_CUDA_TRY(cudaMalloc((void **)p, bytes));
  if (A) A->device_bytes += bytes;
  return B200_OK;
}

template <int THREADS>
static int stream_occupancy(size_t smem, int *ctas)
{
  auto kern = k_stream<B200_MODE_EXACT_FMA, EPI_NONE, THREADS, false, false>;  // the footprint every instantiation is bounded to
  B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, kern, THREADS + 32, smem));
  return B200_OK;
}

template <int MODE, int THREADS>
static int stream_set_attr(size_t smem)
{
#define B200_SET(EPI_, HALO_, I8_)                                                                  \
  B200_CUDA_TRY(cudaFuncSetAttribute(k_stream<MODE, EPI_, THREADS, HALO_, I8_>,                      \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
  B200_SET(EPI_NONE, false, false);     B200_SET(EPI_NONE, false, true);
  B200_SET(EPI_ADD, false, false);      B200_SET(EPI_ADD, false, true);
  B200_SET(EPI_RESIDUAL, false, false); B200_SET(EPI_RESIDUAL, false, true);
  B200_SET(EPI_JACOBI, false, false);   B200_SET(EPI_JACOBI, false, true);
  B200_SET(EPI_DOT, false, false);      B200_SET(EPI_DOT, false, true);
  B200_SET(EPI_NONE, true, false);      B200_SET(EPI_NONE, true, true);
  B200_SET(EPI_DOT, true, false);       B200_SET(EPI_DOT, true, true);
#undef B200_SET
#define B200_SET_RL8(EPI_)                                                                          \
  B200_CUDA_TRY(cudaFuncSetAttribute(k_stream<MODE, EPI_, THREADS, false, true, true>,               \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
  B200_SET_RL8(EPI_NONE); B200_SET_RL8(EPI_ADD); B200_SET_RL8(EPI_RESIDUAL); B200_SET_RL8(EPI_JACOBI); B200_SET_RL8(EPI_DOT);
#undef B200_SET_RL8
  return B200_OK;
}

// The opt-in limit is a per-function, PER-DEVICE attribute: raise it once on each device the process
// uses, to the architectural maximum, so that matrices with different stage sizes can coexist.
static int stream_set_all_attrs(int /*threads*/, size_t /*smem*/)
{
  DeviceState *d = device_state();
  if (!d) return B200_ERR_NO_DEVICE;
  if (d->stream_attrs_set) return B200_OK;
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  if (d->stream_attrs_set) return B200_OK;
  const size_t maxsmem = 227 * 1024;
  B200_TRY((stream_set_attr<B200_MODE_EXACT, 256>(maxsmem)));
  B200_TRY((stream_set_attr<B200_MODE_EXACT_FMA, 256>(maxsmem)));
  B200_TRY((stream_set_attr<B200_MODE_EXACT, 128>(maxsmem)));
  B200_TRY((stream_set_attr<B200_MODE_EXACT_FMA, 128>(maxsmem)));
  d->stream_attrs_set = true;
  return B200_OK;
}

// Tiles of the merge kernel: greedy over rows; a row is split only when it cannot fit a fresh
// tile or the current tile is less than half full.
static int build_merge_plan(b200_csr_s *A, const int32_t *ai)
{
  const int m = A->m;
  std::vector<int4> tiles, split;
  int r = 0, pos = 0;
  int open_row = -1, open_first = -1;  // a row whose tail is still being emitted
  while (r < m) {
    const int s = pos, r0 = r, cap_end = s + MERGE_CAP;
    int rr = r;
    while (rr < m && ai[rr + 1] <= cap_end && (rr - r0) < MERGE_RCAP) ++rr;
    int e, r1, split_row = -1;
    const bool more  = rr < m && (rr - r0) < MERGE_RCAP;
    const int  used  = (rr > r0 ? ai[rr] : s) - s;
    const int  start = more ? std::max(ai[rr], s) : 0;  // where row rr would begin inside this tile
    if (more && start < cap_end && ai[rr + 1] > cap_end && (used * 2 < MERGE_CAP || ai[rr + 1] - start > MERGE_CAP)) {
      e = cap_end; r1 = rr + 1; split_row = rr;   // row rr continues in the next tile
    } else {
      if (rr == r0) return set_error(B200_ERR_STATE, "merge plan made no progress at row %d", r0);
      e = ai[rr]; r1 = rr;
    }
    r = rr; pos = e;
    const int t = (int)tiles.size();
    tiles.push_back(make_int4(r0, r1, s, e));
    // 1. a row split earlier ends here when it is this tile's first row and is complete in it
    if (open_row >= 0 && open_row == r0 && open_first < t && ai[r0 + 1] <= e) {
      split.push_back(make_int4(open_row, open_first, t, 0));
      open_row = -1;
    }
    // 2. a row whose first piece is in this tile
    if (split_row >= 0 && open_row != split_row) { open_row = split_row; open_first = t; }
  }
  A->nmtiles = (int)tiles.size();
  A->nmsplit = (int)split.size();
  if (!A->nmtiles) return B200_OK;
  B200_TRY(dev_alloc(&A->d_mtiles, tiles.size(), A));
  B200_TRY(dev_alloc(&A->d_msplit, split.size(), A));
  B200_TRY(dev_alloc(&A->d_mhead, tiles.size(), A));
  B200_TRY(dev_alloc(&A->d_mtail, tiles.size(), A));
  B200_CUDA_TRY(cudaMemcpy(A->d_mtiles, tiles.data(), tiles.size() * sizeof(int4), cudaMemcpyHostToDevice));
  if (!split.empty()) B200_CUDA_TRY(cudaMemcpy(A->d_msplit, split.data(), split.size() * sizeof(int4), cudaMemcpyHostToDevice));
  B200_CUDA_TRY(cudaMemset(A->d_mhead, 0, tiles.size() * sizeof(double)));
  B200_CUDA_TRY(cudaMemset(A->d_mtail, 0, tiles.size() * sizeof(double)));
  return B200_OK;
}

// Compressed column indices: find the distinct diagonals (col - row); with at most 256 of them
// every column index becomes a one-byte code (see IDX8 at k_stream).  Host work, once per matrix,
// row blocks in parallel.
static int build_idx8(b200_csr_s *A, const int32_t *ai, const int32_t *aj)
{
  A->idx8 = false;
  if (!aj || A->nz == 0 || !env_int("B200_INDEX8", 1)) return B200_OK;
  const int m  = A->m;
  const int nt = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  auto rows_of = [&](int t) { return std::make_pair((int)((long long)m * t / nt), (int)((long long)m * (t + 1) / nt)); };
  std::vector<std::vector<int>> local(nt);
  std::atomic<bool> too_many{false};
  auto scan = [&](int t) {
    std::vector<int> &set = local[t];
    int last = INT32_MIN;
    const auto [r0, r1] = rows_of(t);
    for (int r = r0; r < r1 && !too_many.load(std::memory_order_relaxed); ++r)
      for (int k = ai[r]; k < ai[r + 1]; ++k) {
        const int d = aj[k] - r;
        if (d == last) continue;
        last = d;
        auto it = std::lower_bound(set.begin(), set.end(), d);
        if (it == set.end() || *it != d) {
          set.insert(it, d);
          if (set.size() > 256) { too_many = true; return; }
        }
      }
  };
  {
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(scan, t);
    scan(0);
    for (auto &x : th) x.join();
  }
  if (too_many) return B200_OK;
  std::vector<int> offs;
  for (auto &v : local) offs.insert(offs.end(), v.begin(), v.end());
  std::sort(offs.begin(), offs.end());
  offs.erase(std::unique(offs.begin(), offs.end()), offs.end());
  if (offs.size() > 256) return B200_OK;
  std::vector<unsigned char> codes((size_t)A->nz + 64, 0);
  auto encode = [&](int t) {
    const auto [r0, r1] = rows_of(t);
    for (int r = r0; r < r1; ++r)
      for (int k = ai[r]; k < ai[r + 1]; ++k)
        codes[k] = (unsigned char)(std::lower_bound(offs.begin(), offs.end(), aj[k] - r) - offs.begin());
  };
  {
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(encode, t);
    encode(0);
    for (auto &x : th) x.join();
  }
  A->noffs = (int)offs.size();
  offs.resize(256, 0);
  B200_TRY(dev_alloc(&A->d_aj8, codes.size(), A));
  B200_TRY(dev_alloc(&A->d_offs, offs.size(), A));
  B200_CUDA_TRY(cudaMemcpy(A->d_aj8, codes.data(), codes.size(), cudaMemcpyHostToDevice));
  B200_CUDA_TRY(cudaMemcpy(A->d_offs, offs.data(), offs.size() * sizeof(int), cudaMemcpyHostToDevice));
  A->idx8 = true;
  return B200_OK;
}

// Chunks and work blocks of k_wmerge.  A chunk holds whole rows while they fit WM_CAP non-zeros (and
// at most WM_CAP rows: runs of empty rows); a row longer than WM_CAP becomes pieces of WM_CAP.  A work
// block is ~WM_BLOCK consecutive chunks and never separates the pieces of one row.
#define WM_BLOCK 16
static void wmerge_plan(int m, const int32_t *ai, std::vector<int4> &chunks, std::vector<int> &blk)
{
  chunks.clear(); blk.clear();
  chunks.reserve((size_t)ai[m] / (WM_CAP / 2) + 16);
  int r = 0;
  while (r < m) {
    const int len = ai[r + 1] - ai[r];
    if (len > WM_CAP) {
      for (int k = ai[r]; k < ai[r + 1]; k += WM_CAP) {
        const int k1 = std::min(k + WM_CAP, ai[r + 1]);
        chunks.push_back(make_int4(r, k == ai[r] ? -2 : (k1 == ai[r + 1] ? -3 : -1), k, k1));
      }
      ++r;
      continue;
    }
    int rr = r;
    while (rr < m && (rr - r) < WM_CAP && ai[rr + 1] - ai[rr] <= WM_CAP && ai[rr + 1] - ai[r] <= WM_CAP) ++rr;
    chunks.push_back(make_int4(r, rr - r, ai[r], ai[rr]));
    r = rr;
  }
  blk.push_back(0);
  for (size_t c = 0; c < chunks.size();) {
    size_t e = std::min(chunks.size(), c + WM_BLOCK);
    while (e < chunks.size() && (chunks[e].y == -1 || chunks[e].y == -3)) ++e;   // keep a long row in one block
    blk.push_back((int)e);
    c = e;
  }
}

// the plan itself, host only (no device needed): for the CPU tests, which walk it the way the kernel does
extern "C" int b200_wmerge_plan_size(int32_t m, const int32_t *h_ai, int32_t *nchunks, int32_t *nblocks)
{
  if (m < 0 || !h_ai || !nchunks || !nblocks) return set_error(B200_ERR_ARG, "b200_wmerge_plan_size: bad argument");
  std::vector<int4> chunks;
  std::vector<int>  blk;
  wmerge_plan(m, h_ai, chunks, blk);
  *nchunks = (int32_t)chunks.size();
  *nblocks = (int32_t)blk.size() - 1;
  return B200_OK;
}
extern "C" int b200_wmerge_plan(int32_t m, const int32_t *h_ai, int32_t *chunks4, int32_t *blk)
{
  if (m < 0 || !h_ai || !chunks4 || !blk) return set_error(B200_ERR_ARG, "b200_wmerge_plan: bad argument");
  std::vector<int4> chunks;
  std::vector<int>  b;
  wmerge_plan(m, h_ai, chunks, b);
  for (size_t c = 0; c < chunks.size(); ++c) { chunks4[4 * c] = chunks[c].x; chunks4[4 * c + 1] = chunks[c].y; chunks4[4 * c + 2] = chunks[c].z; chunks4[4 * c + 3] = chunks[c].w; }
  std::copy(b.begin(), b.end(), blk);
  return B200_OK;
}

static int build_wmerge_plan(b200_csr_s *A, const int32_t *ai)
{
  std::vector<int4> chunks;
  std::vector<int>  blk;
  wmerge_plan(A->m, ai, chunks, blk);
  A->nwchunks = (int)chunks.size();
  A->nwblk    = (int)blk.size() - 1;
  B200_TRY(dev_alloc(&A->d_wchunks, chunks.size(), A));
  B200_TRY(dev_alloc(&A->d_wblk, blk.size(), A));
  B200_TRY(dev_alloc(&A->d_wcounters, 2, A));
  if (!chunks.empty()) B200_CUDA_TRY(cudaMemcpy(A->d_wchunks, chunks.data(), chunks.size() * sizeof(int4), cudaMemcpyHostToDevice));
  B200_CUDA_TRY(cudaMemcpy(A->d_wblk, blk.data(), blk.size() * sizeof(int), cudaMemcpyHostToDevice));
  B200_CUDA_TRY(cudaMemset(A->d_wcounters, 0, 2 * sizeof(unsigned)));
  return B200_OK;
}

// Build the plan from the host row-pointer array.
static int build_plan(b200_csr_s *A, const int32_t *ai, const int32_t *aj)
{
  const int m = A->m;
  // --- statistics ---------------------------------------------------------------------------
  A->nonzerorowcnt = 0;
  A->rmax          = 0;
  memset(A->hist, 0, sizeof A->hist);
  for (int i = 0; i < m; ++i) {
    int len = ai[i + 1] - ai[i];
    A->nonzerorowcnt += (len > 0);
    A->rmax = std::max(A->rmax, len);
    A->hist[hist_bucket(len)]++;
  }
  // --- compressed row: zero rows >= 0.6 m ([P376] MatCheckCompressedRow, ratio at
  //     src/openacc-step2/MatAssemblyEnd_SeqAIJ.patch:35 context) -----------------------------
  const int nzero = m - A->nonzerorowcnt;
  A->cprow_use    = !(nzero < 0.6 * m) && m > 0;
  if (A->cprow_use) {
    std::vector<int> cpi(A->nonzerorowcnt + 1), ridx(std::max(A->nonzerorowcnt, 1));
    int row = 0;
    cpi[0]  = 0;
    for (int i = 0; i < m; ++i) {
      if (ai[i + 1] == ai[i]) continue;
      cpi[row + 1] = ai[i + 1];
      ridx[row++]  = i;
    }
    A->cprow_nrows = row;
    B200_TRY(dev_alloc(&A->d_cpi, cpi.size(), A));
    B200_TRY(dev_alloc(&A->d_ridx, ridx.size(), A));
    B200_CUDA_TRY(cudaMemcpy(A->d_cpi, cpi.data(), cpi.size() * sizeof(int), cudaMemcpyHostToDevice));
    B200_CUDA_TRY(cudaMemcpy(A->d_ridx, ridx.data(), ridx.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  // --- stream tiles ---------------------------------------------------------------------------
  const double mean = A->nonzerorowcnt ? (double)A->nz / A->nonzerorowcnt : 0.0;
  // Stream configuration.  Measured on B200 (profiles/r01_sweep_*.log): the x gather is latency
  // bound, so throughput follows the number of RESIDENT consumer threads per SM until ~1024, where
  // HBM saturates.  Pick (threads per CTA, ring depth) that maximises resident consumers; prefer
  // the 2-deep ring once 1024 are reached (7-point: 256 x 2 stages x 4 CTAs; 27-point: 128 x 1
  // stage x 5 CTAs).  B200_STREAM_* environment variables override for sweeps.
  B200_TRY(stream_set_all_attrs(0, 0));  // opt in to > 48 KB before asking for occupancy
  if (A->rmax <= 8192) B200_TRY(build_idx8(A, ai, aj));
  const bool i8 = A->idx8;
  int threads = env_int("B200_STREAM_THREADS", 0), stages = env_int("B200_STREAM_STAGES", 0);
  int cap = env_int("B200_STREAM_CAP", 0);
  auto cap_for = [&](int T) { return std::max(((std::min((int)(T * std::max(mean, 1.0) * 1.05) + 32, 8192) + 3) & ~3), 64); };
  if ((threads != 128 && threads != 256) || stages <= 0) {
    long best = -1;
    int  bt = 256, bs = 2;
    for (int T : {256, 128}) {
      if (threads == 128 || threads == 256) { if (T != threads) continue; }
      for (int S : {2, 1}) {
        if (stages > 0 && S != stages) continue;
        const size_t smem = stream_header_bytes(i8) + (size_t)S * stream_stage_bytes(T, cap > 0 ? ((cap + 3) & ~3) : cap_for(T), i8);
        if (smem > 227 * 1024) continue;
        int ctas = 0;
        if (T == 256) B200_TRY(stream_occupancy<256>(smem, &ctas)); else B200_TRY(stream_occupancy<128>(smem, &ctas));
        // ties: 2-deep ring first; then 256-thread CTAs for short rows, 128-thread CTAs (more, smaller
        // tiles in flight) for long rows -- 27-point: 6.5 vs 5.8 TB/s in r01_sweep_stencil27_200.log
        const long thr = (long)ctas * T, score = std::min(thr, 1024L) * 4 + (S == 2 ? 2 : 0) + ((T == 256) == (mean <= 12.0) ? 1 : 0);
        if (score > best) { best = score; bt = T; bs = S; }
      }
    }
    if (threads != 128 && threads != 256) threads = bt;
    if (stages <= 0) stages = bs;
  }
  if (cap <= 0) cap = cap_for(threads);
  cap = std::max((cap + 3) & ~3, 64);
  A->ntiles  = 0;
  bool stream_ok = (m > 0 && A->nz > 0 && A->rmax <= cap);
  if (stream_ok) {
    std::vector<int4> tiles;
    tiles.reserve((size_t)m / threads + (size_t)A->nz / cap + 2);
    int r = 0;
    while (r < m) {
      int r1 = std::min(m, r + threads);
      if (ai[r1] - ai[r] > cap) {
        const int32_t *ub = std::upper_bound(ai + r, ai + r1 + 1, ai[r] + cap);
        r1 = (int)(ub - ai) - 1;
      }
      if (r1 <= r) { stream_ok = false; break; }
      tiles.push_back(make_int4(r, r1, ai[r], ai[r1]));
      r = r1;
    }
    if (stream_ok) {
      const size_t sbytes = stream_stage_bytes(threads, cap, i8);
      stages = std::min(std::max(stages, 1), 7);  // 2 x 7 mbarriers + the halo arrival word fit the 128-byte header
      A->stream_smem = stream_header_bytes(i8) + (size_t)stages * sbytes;
      if (A->stream_smem > 227 * 1024) { stages = 1; A->stream_smem = stream_header_bytes(i8) + sbytes; }
      if (A->stream_smem > 227 * 1024) stream_ok = false;
    }
    if (stream_ok) {
      A->ntiles = (int)tiles.size();
      A->h_tiles = tiles;
      A->stream_threads = threads;
      A->stream_cap     = cap;
      A->stream_stages  = stages;
      B200_TRY(dev_alloc(&A->d_tiles, tiles.size(), A));
      B200_CUDA_TRY(cudaMemcpy(A->d_tiles, tiles.data(), tiles.size() * sizeof(int4), cudaMemcpyHostToDevice));
      if (i8 && A->rmax <= 255 && env_int("B200_ROWLEN8", 1)) {
        // one-byte row lengths + the first offset of every warp of every tile (see RL8 at k_stream)
        std::vector<unsigned char> rl((size_t)m + 64, 0);
        for (int i = 0; i < m; ++i) rl[i] = (unsigned char)(ai[i + 1] - ai[i]);
        const int wpt = threads / 32;
        std::vector<int> wb(tiles.size() * (size_t)wpt);
        for (size_t t = 0; t < tiles.size(); ++t)
          for (int w = 0; w < wpt; ++w) wb[t * wpt + w] = ai[std::min(tiles[t].x + 32 * w, tiles[t].y)];
        B200_TRY(dev_alloc(&A->d_rl8, rl.size(), A));
        B200_TRY(dev_alloc(&A->d_wbase, wb.size(), A));
        B200_CUDA_TRY(cudaMemcpy(A->d_rl8, rl.data(), rl.size(), cudaMemcpyHostToDevice));
        B200_CUDA_TRY(cudaMemcpy(A->d_wbase, wb.data(), wb.size() * sizeof(int), cudaMemcpyHostToDevice));
        A->rl8 = true;
      }
      B200_TRY(stream_set_all_attrs(threads, A->stream_smem));
      int ctas = 0;
      if (threads == 256) B200_TRY(stream_occupancy<256>(A->stream_smem, &ctas));
      else B200_TRY(stream_occupancy<128>(A->stream_smem, &ctas));
      if (ctas < 1) stream_ok = false;
      int want = env_int("B200_STREAM_CTAS_PER_SM", ctas);
      ctas     = std::max(1, std::min(ctas, want));
      A->stream_grid = std::min(A->ntiles, sm_count() * ctas);
    }
  }
  if (!stream_ok) A->ntiles = 0;
  // --- vector lanes: smallest power of two >= mean row length, 2..32 ---------------------------
  int lanes = 2;
  while (lanes < 32 && lanes < mean) lanes <<= 1;
  A->vector_lanes = lanes;
  // --- choice -----------------------------------------------------------------------------------
  // EXACT needs one accumulator per row: stream when every row fits a stage, else thread per row.
  A->kernel_exact = A->cprow_use ? B200_KERNEL_CPROW : (A->ntiles ? B200_KERNEL_STREAM : B200_KERNEL_ROW);
  // FAST: rows of similar, short length -> stream (SIMT lanes stay balanced);
  //       otherwise sub-warp vector kernel.
  // B200_MERGE_MEAN_ABOVE=k (experiment, off by default): uniformly LONG rows (mean > k; the coarse
  // multigrid operators have 100-200 entries per row) also take the exact-order merge kernels --
  // a stream tile holds only cap/mean rows, so most of its consumer threads would idle.
  const int  merge_mean_above = env_int("B200_MERGE_MEAN_ABOVE", 0);
  const bool regular = A->rmax <= std::max(32.0, 4.0 * mean) && (merge_mean_above <= 0 || mean <= merge_mean_above);
  if (A->cprow_use) A->kernel_fast = B200_KERNEL_CPROW;
  else if (A->ntiles && regular) A->kernel_fast = B200_KERNEL_STREAM;
  else if (A->nz > 0) {
    // skewed row lengths: nnz-balanced merge tiles
    B200_TRY(build_merge_plan(A, ai));
    A->kernel_fast = A->nmtiles ? B200_KERNEL_MERGE : B200_KERNEL_VECTOR;
    // the same matrices in EXACT mode: whole-row tiles summed in CSR order (+ a warp per very long row)
    if (!A->cprow_use && !(A->ntiles && regular)) {
      B200_TRY(build_wmerge_plan(A, ai));
      if (A->nwblk) A->kernel_exact = B200_KERNEL_MERGE;
    }
  } else A->kernel_fast = B200_KERNEL_ROW;
  return B200_OK;
}

// Row blocks for the pipelined host-vector path: tile-aligned, about B200_HOST_BLOCK_ROWS rows
// each (default 2^20; the reference's step 4 uses 983,040, src/openacc-step4/MatMult_SeqAIJ.patch:51).
// need[b] = the x chunk that holds the largest column any row of block b touches; columns ascend
// inside a row, so that is the last entry of each row.
static void build_host_blocks(b200_csr_s *A, const int32_t *ai, const int32_t *aj, const std::vector<int4> &tiles)
{
  A->pb_tile.clear(); A->pb_need.clear();
  if (A->ntiles < 2 || A->m != A->n || !aj) return;
  const int target = std::max(1024, env_int("B200_HOST_BLOCK_ROWS", 1 << 20));
  // Graded sizes: the blocks grow from target/8 at the start and shrink to target/8 at the end, so that
  // the first kernel does not wait for a megarow of x and the last download is short -- the fill and
  // the drain of the pipeline are what separates it from the duplex rate of the link (measured on
  // this pool: 4.76 ms for 216 MB each way at once, 5.35 ms for the MatMult with uniform 1 M-row
  // blocks; scripts/probe_e2e.py).  B200_HOST_BLOCK_GRADED=0: uniform blocks.
  const bool graded = env_int("B200_HOST_BLOCK_GRADED", 1) != 0;
  auto want = [&](int row) {   // block size wanted for a block that starts at `row`
    if (!graded) return target;
    int sz = target;
    for (int k = 0, lim = target / 8, pos = 0; k < 3; ++k, lim *= 2) {   // ramp up: target/8, /4, /2
      if (row < pos + lim) { sz = std::min(sz, lim); break; }
      pos += lim;
    }
    for (int k = 0, lim = target / 8, pos = A->m; k < 3; ++k, lim *= 2) {   // ramp down towards the end
      if (row >= pos - lim) { sz = std::min(sz, lim); break; }
      pos -= lim;
    }
    return std::max(sz, 1024);
  };
  A->pb_tile.push_back(0);
  int rows = 0, need_rows = want(0);
  for (int t = 0; t < A->ntiles; ++t) {
    rows += tiles[t].y - tiles[t].x;
    if (rows >= need_rows && t + 1 < A->ntiles) {
      A->pb_tile.push_back(t + 1);
      rows = 0;
      need_rows = want(tiles[t + 1].x);
    }
  }
  A->pb_tile.push_back(A->ntiles);
  const int nblk = (int)A->pb_tile.size() - 1;
  if (nblk < 2) { A->pb_tile.clear(); return; }
  std::vector<int> rowend(nblk);
  for (int b = 0; b < nblk; ++b) rowend[b] = tiles[A->pb_tile[b + 1] - 1].y;
  A->pb_need.resize(nblk);
  for (int b = 0; b < nblk; ++b) {
    const int r0 = tiles[A->pb_tile[b]].x, r1 = rowend[b];
    int cmax = -1;
    for (int r = r0; r < r1; ++r) if (ai[r + 1] > ai[r]) cmax = std::max(cmax, aj[ai[r + 1] - 1]);
    int j = 0;
    while (j < nblk - 1 && rowend[j] <= cmax) ++j;
    A->pb_need[b] = j;
  }
}

// Column blocks for skewed matrices whose x does not stay in L2.  The gathers of k_wmerge are what
// bounds it, and they run 20 % faster when x is L2 resident (10 M rows, 98 M non-zeros: 0.686 ms with
// a 40 MB x, 0.860 ms with 80 MB; profiles/r02_powerlaw.md) -- the 126 MB L2 keeps ~60 MB of a randomly
// gathered vector under the matrix stream.  So A is cut into column blocks [A_0 | A_1 | ...] of at
// most B200_COLBLOCK_MB (default 40) of x each, every block a CSR of its own with its own k_wmerge
// plan, and y = A x becomes y = A_0 x, y += A_1 x, ... -- MatMultAdd continues the row sum from y, and
// because columns ascend inside a row that is exactly the reference's left-to-right order: same bits.
// An extra copy of aj/aa (like the explicit transpose); dropped by b200_csr_update_values.
// number of column blocks for an x of n doubles (0: no blocking) and the first column past block b
static int colblock_count(int64_t n, size_t limit)
{
  const size_t xbytes = (size_t)n * sizeof(double);
  if (xbytes <= limit + limit / 2) return 0;                        // up to 1.5 blocks the one-pass kernel is as fast
  return (int)((xbytes + limit - 1) / limit);
}
static int32_t colblock_end(int32_t n, int nb, int b) { return (b + 1 == nb) ? n : (int32_t)((long long)n * (b + 1) / nb); }

// host only (for the CPU tests): how many blocks, and for every block the end of its part of every row --
// split[b * m + r] = the first entry of row r that belongs to a later block (so block b holds
// [split[(b-1) * m + r], split[b * m + r]) of row r, block 0 from ai[r])
extern "C" int b200_colblock_split(int32_t m, int32_t n, const int32_t *h_ai, const int32_t *h_aj, int64_t limit_bytes,
                                   int32_t *nblocks, int32_t *split)
{
  if (m < 0 || n < 0 || !h_ai || limit_bytes < 1 || !nblocks) return set_error(B200_ERR_ARG, "b200_colblock_split: bad argument");
  const int nb = colblock_count(n, (size_t)limit_bytes);
  *nblocks = nb;
  if (!split || nb == 0) return B200_OK;
  std::vector<int32_t> from(h_ai, h_ai + m);
  for (int b = 0; b < nb; ++b) {
    const int32_t cend = colblock_end(n, nb, b);
    for (int r = 0; r < m; ++r) {
      int32_t k = from[r];
      while (k < h_ai[r + 1] && h_aj[k] < cend) ++k;
      from[r] = k;
      split[(size_t)b * m + r] = k;
    }
  }
  return B200_OK;
}

static int alloc_mirrors(b200_csr_s *A);
static int fill_ai_tail(b200_csr_s *A);
static int build_colblocks(b200_csr_s *A, const int32_t *ai, const int32_t *aj, const double *aa)
{
  if (A->is_colblock || !aa || !aj || !env_int("B200_COLBLOCK", 1)) return B200_OK;
  // (B200_COLBLOCK_KB: the same limit in KB, so that the tests can block small matrices)
  const int    kb    = env_int("B200_COLBLOCK_KB", 0);
  const size_t limit = kb > 0 ? (size_t)kb << 10 : (size_t)std::max(1, env_int("B200_COLBLOCK_MB", 40)) << 20;
  const int nb = colblock_count(A->n, limit);                       // 0 up to 60 MB of x: the one-pass kernel is as fast
  if (nb == 0) return B200_OK;
  const int m = A->m;
  std::vector<int32_t> bi((size_t)m + 1), bj;
  std::vector<double>  ba;
  std::vector<int32_t> from(ai, ai + m);                            // next unread entry of every row
  for (int b = 0; b < nb; ++b) {
    const int32_t cend = colblock_end(A->n, nb, b);
    bj.clear(); ba.clear();
    bi[0] = 0;
    for (int r = 0; r < m; ++r) {
      int32_t k = from[r];
      while (k < ai[r + 1] && aj[k] < cend) { bj.push_back(aj[k]); ba.push_back(aa[k]); ++k; }
      from[r] = k;
      bi[r + 1] = (int32_t)bj.size();
    }
    b200_csr_s *S = new (std::nothrow) b200_csr_s;
    if (!S) return set_error(B200_ERR_MEM, "out of host memory");
    S->m = m; S->n = A->n; S->nz = bi[m]; S->is_colblock = true;
    A->colblocks.push_back(S);                                      // owned from here on (destroyed with A)
    B200_TRY(alloc_mirrors(S));
    B200_CUDA_TRY(cudaMemcpy(S->d_ai, bi.data(), ((size_t)m + 1) * sizeof(int), cudaMemcpyHostToDevice));
    B200_TRY(fill_ai_tail(S));
    if (S->nz) {
      B200_CUDA_TRY(cudaMemcpy(S->d_aj, bj.data(), (size_t)S->nz * sizeof(int), cudaMemcpyHostToDevice));
      B200_CUDA_TRY(cudaMemcpy(S->d_aa, ba.data(), (size_t)S->nz * sizeof(double), cudaMemcpyHostToDevice));
    }
    // the block runs k_wmerge whatever its own histogram says (half of its rows may be empty)
    S->rmax = A->rmax;
    B200_TRY(build_wmerge_plan(S, bi.data()));
    S->kernel_fast = S->kernel_exact = B200_KERNEL_MERGE;
    A->device_bytes += S->device_bytes;
  }
  return B200_OK;
}

static int create_common(b200_csr_s *A, const int32_t *h_ai, const int32_t *h_aj, const double *h_aa = nullptr)
{
  B200_TRY(build_plan(A, h_ai, h_aj));
  if (A->kernel_exact == B200_KERNEL_MERGE && A->nwblk) B200_TRY(build_colblocks(A, h_ai, h_aj, h_aa));
  return B200_OK;
}

extern "C" int b200_init(int device)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return set_error(B200_ERR_NO_DEVICE, "no CUDA device visible: the b200 library has no CPU fallback");
  }
  if (device < 0 || device >= n) return set_error(B200_ERR_ARG, "device %d out of range [0,%d)", device, n);
  B200_CUDA_TRY(cudaSetDevice(device));
  return ensure_device();
}

extern "C" const char *b200_last_error(void) { return g_last_error.c_str(); }
extern "C" uint64_t    b200_launch_count(void) { return g_launches.load(); }
extern "C" int         b200_device_sm_count(void) { return ensure_device() ? 0 : sm_count(); }
extern "C" const char *b200_version(void) { return "b200-seqaij 0.2 (sm_100a)"; }

static int alloc_mirrors(b200_csr_s *A)
{
  // +8 elements of slack: the bulk copies round tile ends up to 16-byte multiples
  B200_TRY(dev_alloc(&A->d_ai, (size_t)A->m + 1 + 8, A));
  B200_TRY(dev_alloc(&A->d_aj, (size_t)A->nz + 8, A));
  B200_TRY(dev_alloc(&A->d_aa, (size_t)A->nz + 8, A));
  B200_CUDA_TRY(cudaMemset(A->d_aj + A->nz, 0, 8 * sizeof(int)));
  B200_CUDA_TRY(cudaMemset(A->d_aa + A->nz, 0, 8 * sizeof(double)));
  return B200_OK;
}

static int fill_ai_tail(b200_csr_s *A)
{
  int tail[8];
  for (int k = 0; k < 8; ++k) tail[k] = A->nz;
  B200_CUDA_TRY(cudaMemcpy(A->d_ai + A->m + 1, tail, sizeof tail, cudaMemcpyHostToDevice));
  return B200_OK;
}

extern "C" int b200_csr_create(b200_csr_t *out, int32_t m, int32_t n, const int32_t *h_ai,
                               const int32_t *h_aj, const double *h_aa)
{
  NvtxRange nvtx_("b200_csr_create");
  if (!out || m < 0 || n < 0 || !h_ai) return set_error(B200_ERR_ARG, "b200_csr_create: bad argument");
  B200_TRY(ensure_device());
  const int32_t nz = h_ai[m];
  if (nz < 0 || h_ai[0] != 0) return set_error(B200_ERR_ARG, "b200_csr_create: ai[0] must be 0 and ai[m] >= 0");
  if (nz > 0 && (!h_aj || !h_aa)) return set_error(B200_ERR_ARG, "b200_csr_create: null aj/aa with nz > 0");
  b200_csr_s *A = new (std::nothrow) b200_csr_s;
  if (!A) return set_error(B200_ERR_MEM, "out of host memory");
  A->m = m; A->n = n; A->nz = nz;
  int rc = alloc_mirrors(A);
  if (!rc) rc = [&]() -> int {
    B200_CUDA_TRY(cudaMemcpy(A->d_ai, h_ai, ((size_t)m + 1) * sizeof(int), cudaMemcpyHostToDevice));
    B200_TRY(fill_ai_tail(A));
    if (nz) {
      B200_CUDA_TRY(cudaMemcpy(A->d_aj, h_aj, (size_t)nz * sizeof(int), cudaMemcpyHostToDevice));
      B200_CUDA_TRY(cudaMemcpy(A->d_aa, h_aa, (size_t)nz * sizeof(double), cudaMemcpyHostToDevice));
    }
    B200_TRY(create_common(A, h_ai, h_aj, h_aa));
    if (A->ntiles) build_host_blocks(A, h_ai, h_aj, A->h_tiles);
    return B200_OK;
  }();
  if (rc) { b200_csr_destroy(A); return rc; }
  *out = A;
  return B200_OK;
}

extern "C" int b200_csr_create_from_device(b200_csr_t *out, int32_t m, int32_t n,
                                           const int32_t *d_ai, const int32_t *d_aj,
                                           const double *d_aa)
{
  if (!out || m < 0 || n < 0 || !d_ai) return set_error(B200_ERR_ARG, "b200_csr_create_from_device: bad argument");
  B200_TRY(ensure_device());
  std::vector<int32_t> h_ai((size_t)m + 1);
  B200_CUDA_TRY(cudaMemcpy(h_ai.data(), d_ai, h_ai.size() * sizeof(int), cudaMemcpyDeviceToHost));
  const int32_t nz = h_ai[m];
  b200_csr_s *A = new (std::nothrow) b200_csr_s;
  if (!A) return set_error(B200_ERR_MEM, "out of host memory");
  A->m = m; A->n = n; A->nz = nz;
  int rc = alloc_mirrors(A);
  if (!rc) rc = [&]() -> int {
    B200_CUDA_TRY(cudaMemcpy(A->d_ai, d_ai, ((size_t)m + 1) * sizeof(int), cudaMemcpyDeviceToDevice));
    B200_TRY(fill_ai_tail(A));
    if (nz) {
      B200_CUDA_TRY(cudaMemcpy(A->d_aj, d_aj, (size_t)nz * sizeof(int), cudaMemcpyDeviceToDevice));
      B200_CUDA_TRY(cudaMemcpy(A->d_aa, d_aa, (size_t)nz * sizeof(double), cudaMemcpyDeviceToDevice));
    }
    // the column indices come to the host once: the diagonal-code detection and the host-vector
    // pipeline blocks are host work (a transpose built on the device gets the same plan as a matrix
    // created from host arrays)
    std::vector<int32_t> h_aj((size_t)nz);
    if (nz) B200_CUDA_TRY(cudaMemcpy(h_aj.data(), d_aj, (size_t)nz * sizeof(int), cudaMemcpyDeviceToHost));
    B200_TRY(create_common(A, h_ai.data(), nz ? h_aj.data() : nullptr));
    if (A->ntiles && nz) build_host_blocks(A, h_ai.data(), h_aj.data(), A->h_tiles);
    return B200_OK;
  }();
  if (rc) { b200_csr_destroy(A); return rc; }
  *out = A;
  return B200_OK;
}

static void sell_drop(b200_csr_s *A)
{
  cudaFree(A->d_sval); cudaFree(A->d_scol); cudaFree(A->d_sperm); cudaFree(A->d_scode); cudaFree(A->d_scs);
  A->d_sval = nullptr; A->d_scol = nullptr; A->d_sperm = nullptr; A->d_scode = nullptr; A->d_scs = nullptr;
  A->sell_chunks = 0; A->sell_sigma = 0; A->sell_padded = 0;
}

extern "C" int b200_csr_update_values(b200_csr_t A, const double *h_aa)
{
  if (!A || (!h_aa && A->nz)) return set_error(B200_ERR_ARG, "b200_csr_update_values: bad argument");
  if (A->nz) B200_CUDA_TRY(cudaMemcpy(A->d_aa, h_aa, (size_t)A->nz * sizeof(double), cudaMemcpyHostToDevice));
  if (A->T) { b200_csr_destroy(A->T); A->T = nullptr; }  // stale transpose values
  for (auto *b : A->colblocks) b200_csr_destroy(b);       // stale column-block copies: back to the one-pass kernel
  A->colblocks.clear();
  if (A->sell_chunks) {                                  // stale SELL values: back to the plan's kernel
    sell_drop(A);
    if (A->kernel_override == B200_KERNEL_SELL) A->kernel_override = 0;
  }
  return B200_OK;
}

extern "C" int b200_csr_destroy(b200_csr_t A)
{
  if (!A) return B200_OK;
  if (A->T) b200_csr_destroy(A->T);
  for (auto *b : A->colblocks) b200_csr_destroy(b);
  cudaFree(A->d_ai); cudaFree(A->d_aj); cudaFree(A->d_aa);
  cudaFree(A->d_cpi); cudaFree(A->d_ridx); cudaFree(A->d_tiles);
  cudaFree(A->d_aj8); cudaFree(A->d_offs); cudaFree(A->d_rl8); cudaFree(A->d_wbase);
  cudaFree(A->d_mtiles); cudaFree(A->d_msplit); cudaFree(A->d_mhead); cudaFree(A->d_mtail);
  cudaFree(A->d_wchunks); cudaFree(A->d_wblk); cudaFree(A->d_wcounters);
  sell_drop(A);
  cudaFree(A->d_hx); cudaFree(A->d_hy);
  for (auto &s : A->hs) if (s) cudaStreamDestroy(s);
  for (auto &e : A->hev) if (e) cudaEventDestroy(e);
  for (auto &e : A->pb_evx) if (e) cudaEventDestroy(e);
  for (auto &e : A->pb_evk) if (e) cudaEventDestroy(e);
  delete A;
  return B200_OK;
}

extern "C" int b200_csr_get_info(b200_csr_t A, b200_csr_info_t *info)
{
  if (!A || !info) return set_error(B200_ERR_ARG, "b200_csr_get_info: bad argument");
  memset(info, 0, sizeof *info);
  info->m = A->m; info->n = A->n; info->nz = A->nz;
  info->nonzerorowcnt = A->nonzerorowcnt; info->rmax = A->rmax;
  info->compressedrow_use = A->cprow_use; info->cprow_nrows = A->cprow_nrows;
  info->kernel_fast = A->kernel_fast; info->kernel_exact = A->kernel_exact;
  info->vector_lanes = A->vector_lanes; info->stream_tiles = A->ntiles;
  info->merge_tiles = A->nmtiles; info->has_transpose = A->T != nullptr;
  info->index8_diagonals = A->idx8 ? A->noffs : 0;
  memcpy(info->hist, A->hist, sizeof A->hist);
  info->device_bytes = A->device_bytes + (A->T ? A->T->device_bytes : 0);
  info->sell_chunks = A->sell_chunks; info->sell_sigma = A->sell_sigma; info->sell_padded_nnz = A->sell_padded;
  info->rowlen8 = A->rl8 ? 1 : 0;
  return B200_OK;
}

extern "C" int b200_csr_set_kernel(b200_csr_t A, int kernel)
{
  if (!A || kernel < 0 || kernel > B200_KERNEL_SELL) return set_error(B200_ERR_ARG, "b200_csr_set_kernel: bad argument");
  if (kernel == B200_KERNEL_SELL && !A->sell_chunks) return set_error(B200_ERR_STATE, "no SELL copy: call b200_csr_build_sell first");
  if (kernel == B200_KERNEL_STREAM && !A->ntiles) return set_error(B200_ERR_STATE, "stream kernel not applicable: a row exceeds the stage capacity");
  if (kernel == B200_KERNEL_CPROW && !A->cprow_use) return set_error(B200_ERR_STATE, "matrix has no compressed-row index");
  if (kernel == B200_KERNEL_MERGE && !A->nmtiles) {
    std::vector<int32_t> ai((size_t)A->m + 1);
    B200_CUDA_TRY(cudaMemcpy(ai.data(), A->d_ai, ai.size() * sizeof(int), cudaMemcpyDeviceToHost));
    B200_TRY(build_merge_plan(A, ai.data()));
    if (!A->nmtiles) return set_error(B200_ERR_STATE, "merge kernel not applicable (empty matrix)");
    if (!A->nwblk) B200_TRY(build_wmerge_plan(A, ai.data()));   // the exact-order form of the same override
  }
  A->kernel_override = kernel;
  return B200_OK;
}

extern "C" int b200_csr_device_arrays(b200_csr_t A, const int32_t **d_ai, const int32_t **d_aj,
                                      const double **d_aa)
{
  if (!A) return set_error(B200_ERR_ARG, "null handle");
  if (d_ai) *d_ai = A->d_ai;
  if (d_aj) *d_aj = A->d_aj;
  if (d_aa) *d_aa = A->d_aa;
  return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------------------------
// one place that instantiates the stream kernel for (threads per CTA, halo, index width)
template <int MODE, int EPI, bool HALO>
static int launch_stream_any(b200_csr_s *A, int grid, const int4 *tiles, int ntiles, const double *x,
                             const double *yin, double *y, const HaloArgs &h, cudaStream_t st,
                             const double *aux = nullptr, const DotArgs &dot = DotArgs{})
{
  const Idx8Args ix{A->d_aj8, A->d_offs, A->d_rl8, A->d_wbase};
  B200_TRY(stream_set_all_attrs(0, 0));   // first use on this device (a handle made on another one)
#define B200_STREAM_GO(T, I8)                                                                       \
  B200_LAUNCH_PDL((k_stream<MODE, EPI, T, HALO, I8>), grid, T + 32, A->stream_smem, st, tiles, ntiles, \
                  (const int *)A->d_ai, (const int *)A->d_aj, (const double *)A->d_aa, x, yin, y,     \
                  (int)A->stream_cap, (int)A->stream_stages, h, ix, aux, dot)
  // one-byte row lengths: whole-matrix launches of the byte-code plan (the tile index addresses wbase)
  if (!HALO && A->rl8 && tiles == A->d_tiles) {
#define B200_STREAM_GO_RL8(T)                                                                       \
  B200_LAUNCH_PDL((k_stream<MODE, EPI, T, false, true, true>), grid, T + 32, A->stream_smem, st, tiles, ntiles, \
                  (const int *)A->d_ai, (const int *)A->d_aj, (const double *)A->d_aa, x, yin, y,     \
                  (int)A->stream_cap, (int)A->stream_stages, h, ix, aux, dot)
    if (A->stream_threads == 256) B200_STREAM_GO_RL8(256); else B200_STREAM_GO_RL8(128);
#undef B200_STREAM_GO_RL8
    return B200_OK;
  }
  if (A->idx8) { if (A->stream_threads == 256) B200_STREAM_GO(256, true); else B200_STREAM_GO(128, true); }
  else { if (A->stream_threads == 256) B200_STREAM_GO(256, false); else B200_STREAM_GO(128, false); }
#undef B200_STREAM_GO
  return B200_OK;
}

template <int MODE, bool ADD>
static int launch_stream(b200_csr_s *A, const double *x, const double *yin, double *y, cudaStream_t st)
{
  return launch_stream_any<MODE, ADD ? EPI_ADD : EPI_NONE, false>(A, A->stream_grid, A->d_tiles, A->ntiles, x, yin, y, HaloArgs{}, st);
}

// tiles [t0, t0 + nt) only: the row-blocked host pipeline
template <int MODE, bool ADD>
static int launch_stream_range(b200_csr_s *A, int t0, int nt, const double *x, const double *yin, double *y, cudaStream_t st)
{
  return launch_stream_any<MODE, ADD ? EPI_ADD : EPI_NONE, false>(A, std::min(nt, A->stream_grid), A->d_tiles + t0, nt, x, yin, y, HaloArgs{}, st);
}

// MatMult_MPIAIJ in one launch (see k_stream, HALO): push prologue on the first CTAs, then the
// persistent stream loop, then the ghost rows.
template <int MODE>
static int launch_stream_halo_mode(b200_csr_s *A, const double *x, double *y, const HaloArgs &h, cudaStream_t st, const DotArgs *dot)
{
  const int grid = A->stream_grid;  // cta_ptr / cta_rows were built for exactly this grid
  if (h.npush > grid) return set_error(B200_ERR_STATE, "more push blocks (%d) than CTAs (%d)", h.npush, grid);
  if (dot) return launch_stream_any<MODE, EPI_DOT, true>(A, grid, A->d_tiles, A->ntiles, x, nullptr, y, h, st, nullptr, *dot);
  return launch_stream_any<MODE, EPI_NONE, true>(A, grid, A->d_tiles, A->ntiles, x, nullptr, y, h, st);
}

// (x, y) over m rows into *out: the reduction of the plans that cannot fold it into the MatMult
__global__ void __launch_bounds__(256) k_dot_simple(int m, const double *__restrict__ x, const double *__restrict__ y, double *out)
{
  __shared__ double sw[8];
  double t = 0.0;
  for (int i = threadIdx.x; i < m; i += 256) t = __fma_rn(x[i], y[i], t);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += sw[w];
    *out = s;
  }
}

namespace b200 {
int stream_plan_tiles(b200_csr_t A, int4 **d_tiles, int *ntiles, int *grid, int *threads)
{
  *d_tiles = A->d_tiles;
  *grid    = A->stream_grid;
  *threads = A->stream_threads;
  *ntiles  = A->kernel_override && A->kernel_override != B200_KERNEL_STREAM ? 0 : A->ntiles;
  return B200_OK;
}
int launch_stream_halo(b200_csr_t A, const double *x, double *y, int mode, const HaloArgs &h, cudaStream_t st, const DotArgs *dot)
{
  if (!A->ntiles) return set_error(B200_ERR_STATE, "fused halo launch needs the stream plan");
  if (mode == B200_MODE_EXACT) return launch_stream_halo_mode<B200_MODE_EXACT>(A, x, y, h, st, dot);
  return launch_stream_halo_mode<B200_MODE_EXACT_FMA>(A, x, y, h, st, dot);
}
int stream_grid_of(b200_csr_t A)
{
  const int k = A->kernel_override ? A->kernel_override : A->kernel_exact;
  return (k == B200_KERNEL_STREAM && A->ntiles) ? A->stream_grid : 0;
}
}  // namespace b200

template <bool ADD>
static int launch_vector(b200_csr_s *A, const double *x, const double *yin, double *y, cudaStream_t st)
{
  const int       L      = A->vector_lanes;
  const long long thr    = (long long)A->m * L;
  const unsigned  blocks = (unsigned)((thr + 255) / 256);
  switch (L) {
    case 2: B200_LAUNCH((k_vector<2, ADD>), blocks, 256, 0, st, A->m, A->d_ai, A->d_aj, A->d_aa, x, yin, y); break;
    case 4: B200_LAUNCH((k_vector<4, ADD>), blocks, 256, 0, st, A->m, A->d_ai, A->d_aj, A->d_aa, x, yin, y); break;
    case 8: B200_LAUNCH((k_vector<8, ADD>), blocks, 256, 0, st, A->m, A->d_ai, A->d_aj, A->d_aa, x, yin, y); break;
    case 16: B200_LAUNCH((k_vector<16, ADD>), blocks, 256, 0, st, A->m, A->d_ai, A->d_aj, A->d_aa, x, yin, y); break;
    default: B200_LAUNCH((k_vector<32, ADD>), blocks, 256, 0, st, A->m, A->d_ai, A->d_aj, A->d_aa, x, yin, y); break;
  }
  return B200_OK;
}

template <int MODE, bool ADD>
static int launch_mode(b200_csr_s *A, int kernel, const double *x, const double *yin, double *y, cudaStream_t st)
{
  switch (kernel) {
    case B200_KERNEL_STREAM: return launch_stream<MODE, ADD>(A, x, yin, y, st);
    case B200_KERNEL_CPROW:
      if (!ADD) B200_CUDA_TRY(cudaMemsetAsync(y, 0, (size_t)A->m * sizeof(double), st));
      else if (yin != y) {
        B200_LAUNCH(k_copy, std::max(1, std::min(sm_count() * 8, (A->m + 255) / 256)), 256, 0, st, y, yin, (long long)A->m);
        yin = y;
      }
      if (A->cprow_nrows)
        B200_LAUNCH((k_cprow<MODE, ADD>), (A->cprow_nrows + 127) / 128, 128, 0, st, A->cprow_nrows,
                    A->d_cpi, A->d_ridx, A->d_aj, A->d_aa, x, yin, y);
      return B200_OK;
    case B200_KERNEL_ROW:
    default:
      B200_LAUNCH((k_row<MODE, ADD>), (A->m + 127) / 128, 128, 0, st, A->m, A->d_ai, A->d_aj, A->d_aa, x, yin, y);
      return B200_OK;
  }
}

// ---------------------------------------------------------------------------------------------
// SELL-32-sigma copy: host packing (pure host code, also exported for the CPU tests), upload, launch
// ---------------------------------------------------------------------------------------------
// Slot order: inside every window of `sigma` rows, rows by decreasing length (stable, so sigma = 1
// or equal lengths keep the natural order); slots past the last row are padding (-1).
static void sell_order(int32_t m, const int32_t *ai, int32_t sigma, std::vector<int32_t> &order)
{
  const size_t nslots = ((size_t)m + SELL_C - 1) / SELL_C * SELL_C;
  order.assign(nslots, -1);
  for (int32_t r = 0; r < m; ++r) order[r] = r;
  if (sigma > 1)
    for (int64_t w = 0; w < m; w += sigma)
      std::stable_sort(order.begin() + w, order.begin() + std::min<int64_t>(m, w + sigma),
                       [&](int32_t a, int32_t b) { return ai[a + 1] - ai[a] > ai[b + 1] - ai[b]; });
}

extern "C" int b200_sell_pack_size(int32_t m, const int32_t *h_ai, int32_t sigma, int32_t *nchunks, uint64_t *padded)
{
  if (m < 0 || (m && !h_ai) || sigma < 1 || !nchunks || !padded) return set_error(B200_ERR_ARG, "b200_sell_pack_size: bad argument");
  std::vector<int32_t> order;
  sell_order(m, h_ai, sigma, order);
  const int32_t nc = (int32_t)(order.size() / SELL_C);
  uint64_t tot = 0;
  for (int32_t c = 0; c < nc; ++c) {
    int32_t len = 0;
    for (int l = 0; l < SELL_C; ++l) { const int32_t r = order[(size_t)c * SELL_C + l]; if (r >= 0) len = std::max(len, h_ai[r + 1] - h_ai[r]); }
    tot += (uint64_t)len * SELL_C;
  }
  *nchunks = nc; *padded = tot;
  return B200_OK;
}

// cs[nchunks + 1], perm[nchunks * 32], val[padded], col[padded] (col -1 = padding)
extern "C" int b200_sell_pack(int32_t m, const int32_t *h_ai, const int32_t *h_aj, const double *h_aa, int32_t sigma,
                              uint32_t *cs, int32_t *perm, double *val, int32_t *col)
{
  if (m < 0 || sigma < 1 || !cs || (m && (!h_ai || !perm))) return set_error(B200_ERR_ARG, "b200_sell_pack: bad argument");
  std::vector<int32_t> order;
  sell_order(m, h_ai, sigma, order);
  const int32_t nc = (int32_t)(order.size() / SELL_C);
  uint64_t tot = 0;
  cs[0] = 0;
  for (int32_t c = 0; c < nc; ++c) {
    int32_t len = 0;
    for (int l = 0; l < SELL_C; ++l) { const int32_t r = order[(size_t)c * SELL_C + l]; if (r >= 0) len = std::max(len, h_ai[r + 1] - h_ai[r]); }
    tot += (uint64_t)len * SELL_C;
    if (tot > 0xFFFFFFFFull) return set_error(B200_ERR_ARG, "SELL copy exceeds 2^32 entries");
    cs[c + 1] = (uint32_t)tot;
  }
  std::copy(order.begin(), order.end(), perm);
  const int nt = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  auto fill = [&](int t) {
    for (int32_t c = (int32_t)((int64_t)nc * t / nt); c < (int32_t)((int64_t)nc * (t + 1) / nt); ++c) {
      const int32_t len = (int32_t)((cs[c + 1] - cs[c]) / SELL_C);
      for (int l = 0; l < SELL_C; ++l) {
        const int32_t r = order[(size_t)c * SELL_C + l];
        const int32_t n = (r >= 0) ? h_ai[r + 1] - h_ai[r] : 0;
        for (int32_t k = 0; k < len; ++k) {
          const size_t at = (size_t)cs[c] + (size_t)k * SELL_C + l;
          val[at] = (k < n) ? h_aa[h_ai[r] + k] : 0.0;
          col[at] = (k < n) ? h_aj[h_ai[r] + k] : -1;
        }
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(fill, t);
  fill(0);
  for (auto &x : th) x.join();
  return B200_OK;
}

extern "C" int b200_csr_build_sell(b200_csr_t A, int32_t sigma)
{
  NvtxRange nvtx_("b200_csr_build_sell");
  if (!A || sigma < 1) return set_error(B200_ERR_ARG, "b200_csr_build_sell: bad argument");
  B200_TRY(ensure_device());
  sell_drop(A);
  if (A->kernel_override == B200_KERNEL_SELL) A->kernel_override = 0;
  if (A->m == 0) return B200_OK;
  std::vector<int32_t> ai((size_t)A->m + 1), aj((size_t)std::max(A->nz, 1));
  std::vector<double>  aa((size_t)std::max(A->nz, 1));
  B200_CUDA_TRY(cudaMemcpy(ai.data(), A->d_ai, ai.size() * sizeof(int), cudaMemcpyDeviceToHost));
  if (A->nz) {
    B200_CUDA_TRY(cudaMemcpy(aj.data(), A->d_aj, (size_t)A->nz * sizeof(int), cudaMemcpyDeviceToHost));
    B200_CUDA_TRY(cudaMemcpy(aa.data(), A->d_aa, (size_t)A->nz * sizeof(double), cudaMemcpyDeviceToHost));
  }
  int32_t  nc = 0;
  uint64_t padded = 0;
  B200_TRY(b200_sell_pack_size(A->m, ai.data(), sigma, &nc, &padded));
  std::vector<uint32_t> cs((size_t)nc + 1);
  std::vector<int32_t>  perm((size_t)nc * SELL_C), col((size_t)std::max<uint64_t>(padded, 1));
  std::vector<double>   val((size_t)std::max<uint64_t>(padded, 1));
  B200_TRY(b200_sell_pack(A->m, ai.data(), aj.data(), aa.data(), sigma, cs.data(), perm.data(), val.data(), col.data()));
  // the kernel reads up to SELL_U - 1 chunk columns past a chunk's end without a bounds test
  const size_t tail = (size_t)SELL_U * SELL_C;
  val.resize(val.size() + tail, 0.0);
  col.resize(col.size() + tail, -1);
  B200_TRY(dev_alloc(&A->d_scs, cs.size(), A));
  B200_TRY(dev_alloc(&A->d_sval, val.size(), A));
  B200_CUDA_TRY(cudaMemcpy(A->d_scs, cs.data(), cs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  B200_CUDA_TRY(cudaMemcpy(A->d_sval, val.data(), val.size() * sizeof(double), cudaMemcpyHostToDevice));
  if (sigma > 1) {
    B200_TRY(dev_alloc(&A->d_sperm, perm.size(), A));
    B200_CUDA_TRY(cudaMemcpy(A->d_sperm, perm.data(), perm.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  if (A->idx8 && A->noffs < SELL_PAD_CODE) {
    // one byte per entry: the code of its diagonal (col - row), 255 = padding
    std::vector<int> offs(256);
    B200_CUDA_TRY(cudaMemcpy(offs.data(), A->d_offs, 256 * sizeof(int), cudaMemcpyDeviceToHost));
    offs.resize((size_t)A->noffs);
    std::vector<unsigned char> code(col.size(), SELL_PAD_CODE);
    for (int32_t c = 0; c < nc; ++c)
      for (size_t at = cs[c]; at < cs[c + 1]; ++at) {
        if (col[at] < 0) continue;
        const int32_t r = perm[(size_t)c * SELL_C + (at - cs[c]) % SELL_C];
        code[at] = (unsigned char)(std::lower_bound(offs.begin(), offs.end(), col[at] - r) - offs.begin());
      }
    B200_TRY(dev_alloc(&A->d_scode, code.size(), A));
    B200_CUDA_TRY(cudaMemcpy(A->d_scode, code.data(), code.size(), cudaMemcpyHostToDevice));
  } else {
    B200_TRY(dev_alloc(&A->d_scol, col.size(), A));
    B200_CUDA_TRY(cudaMemcpy(A->d_scol, col.data(), col.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  A->sell_chunks = nc; A->sell_sigma = sigma; A->sell_padded = padded;
  return B200_OK;
}

template <bool ADD>
static int launch_sell(b200_csr_s *A, const double *x, const double *yin, double *y, int mode, cudaStream_t st)
{
  if (!A->sell_chunks) return set_error(B200_ERR_STATE, "no SELL copy: call b200_csr_build_sell first");
  const unsigned grid = (unsigned)(((size_t)A->sell_chunks * SELL_C + 255) / 256);
#define B200_SELL_GO(MODE_, I8, PM)                                                                                   \
  B200_LAUNCH((k_sell<MODE_, ADD, I8, PM>), grid, 256, 0, st, A->m, A->sell_chunks, A->d_scs, A->d_sval, A->d_scol, \
              A->d_scode, A->d_offs, A->d_sperm, x, yin, y)
#define B200_SELL_MODE(MODE_)                                                      \
  do {                                                                             \
    if (A->d_scode) { if (A->d_sperm) B200_SELL_GO(MODE_, true, true); else B200_SELL_GO(MODE_, true, false); }   \
    else { if (A->d_sperm) B200_SELL_GO(MODE_, false, true); else B200_SELL_GO(MODE_, false, false); }            \
  } while (0)
  if (mode == B200_MODE_EXACT) B200_SELL_MODE(B200_MODE_EXACT); else B200_SELL_MODE(B200_MODE_EXACT_FMA);
#undef B200_SELL_MODE
#undef B200_SELL_GO
  return B200_OK;
}

template <bool ADD>
static int spmv_dispatch(b200_csr_s *A, const double *x, const double *yin, double *y, int mode, cudaStream_t st)
{
  if (A->m == 0) return B200_OK;
  int kernel = A->kernel_override;
  if (!kernel) kernel = (mode == B200_MODE_FAST) ? A->kernel_fast : A->kernel_exact;
  if (kernel == B200_KERNEL_SELL) return launch_sell<ADD>(A, x, yin, y, mode, st);
  // Skewed matrices: the exact-order kernel is also the fastest measured (configs[4]: 0.86 ms vs
  // 1.23 ms for the split-row merge), so FAST uses it too; B200_MERGE_SPLIT=1 keeps the split-row
  // merge kernel reachable for comparison.
  const bool have_exact_plan = A->nwblk > 0;
  if (kernel == B200_KERNEL_MERGE && (mode != B200_MODE_FAST || (have_exact_plan && !env_int("B200_MERGE_SPLIT", 0)))) {
    if (!have_exact_plan) return set_error(B200_ERR_ARG, "no exact-order merge plan for this matrix");
    if (!A->colblocks.empty()) {
      // y = A_0 x, then y += A_b x block after block: the row sums continue left to right (build_colblocks)
      for (size_t b = 0; b < A->colblocks.size(); ++b) {
        if (b == 0) B200_TRY(spmv_dispatch<ADD>(A->colblocks[0], x, yin, y, mode, st));
        else B200_TRY(spmv_dispatch<true>(A->colblocks[b], x, y, y, mode, st));
      }
      return B200_OK;
    }
    // persistent warps: as many CTAs as fit (8 warps, 12-21 KB of shared memory each)
    const int grid = std::max(1, std::min(sm_count() * 6, (A->nwblk + WM_WARPS - 1) / WM_WARPS));
    if (mode == B200_MODE_EXACT_FMA)
      B200_LAUNCH((k_wmerge<B200_MODE_EXACT_FMA, ADD>), grid, WM_WARPS * 32, 0, st, A->d_wchunks, A->d_wblk, A->nwblk, A->d_wcounters, A->d_ai, A->d_aj, A->d_aa, x, yin, y);
    else
      B200_LAUNCH((k_wmerge<B200_MODE_EXACT, ADD>), grid, WM_WARPS * 32, 0, st, A->d_wchunks, A->d_wblk, A->nwblk, A->d_wcounters, A->d_ai, A->d_aj, A->d_aa, x, yin, y);
    return B200_OK;
  }
  if (mode != B200_MODE_FAST && kernel == B200_KERNEL_VECTOR)
    return set_error(B200_ERR_ARG, "kernel %d cannot honour an EXACT summation order", kernel);
  if (kernel == B200_KERNEL_VECTOR) return launch_vector<ADD>(A, x, yin, y, st);
  if (kernel == B200_KERNEL_MERGE) {
    B200_LAUNCH((k_merge<ADD>), A->nmtiles, MERGE_THREADS, 0, st, A->d_mtiles, A->d_ai, A->d_aj, A->d_aa, x, yin, y, A->d_mhead, A->d_mtail);
    if (A->nmsplit)
      B200_LAUNCH((k_merge_fixup<ADD>), (A->nmsplit + 127) / 128, 128, 0, st, A->nmsplit, A->d_msplit, A->d_mhead, A->d_mtail, yin, y);
    return B200_OK;
  }
  if (mode == B200_MODE_EXACT) return launch_mode<B200_MODE_EXACT, ADD>(A, kernel, x, yin, y, st);
  return launch_mode<B200_MODE_EXACT_FMA, ADD>(A, kernel, x, yin, y, st);
}

static int check_mode(int mode)
{
  if (mode < B200_MODE_FAST || mode > B200_MODE_EXACT_FMA) return set_error(B200_ERR_ARG, "unknown mode %d", mode);
  return B200_OK;
}

extern "C" int b200_spmv(b200_csr_t A, const double *d_x, double *d_y, int mode, void *stream)
{
  NvtxRange nvtx_("b200_spmv");
  if (!A || (!d_x && A->n) || (!d_y && A->m)) return set_error(B200_ERR_ARG, "b200_spmv: null argument");
  if (d_x == d_y && A->m) return set_error(B200_ERR_ARG, "b200_spmv: x and y must differ (MatMult contract)");
  B200_TRY(check_mode(mode));
  return spmv_dispatch<false>(A, d_x, nullptr, d_y, mode, (cudaStream_t)stream);
}

extern "C" int b200_spmv_add(b200_csr_t A, const double *d_x, const double *d_y, double *d_z,
                             int mode, void *stream)
{
  NvtxRange nvtx_("b200_spmv_add");
  if (!A || (!d_x && A->n) || ((!d_y || !d_z) && A->m)) return set_error(B200_ERR_ARG, "b200_spmv_add: null argument");
  if ((d_x == d_z) && A->m) return set_error(B200_ERR_ARG, "b200_spmv_add: x and z must differ");
  B200_TRY(check_mode(mode));
  return spmv_dispatch<true>(A, d_x, d_y, d_z, mode, (cudaStream_t)stream);
}

// r = b - A x and xnew = x + dinv .* (b - A x): fused into the stream kernel's epilogue when the plan
// is the stream kernel, otherwise MatMult followed by one element-wise kernel.
static int spmv_epilogue(b200_csr_s *A, int epi, const double *x, const double *b, const double *dinv, double *y, int mode, cudaStream_t st)
{
  if (A->m == 0) return B200_OK;
  int kernel = A->kernel_override;
  if (!kernel) kernel = (mode == B200_MODE_FAST) ? A->kernel_fast : A->kernel_exact;
  if (kernel == B200_KERNEL_STREAM) {
    const HaloArgs none{};
    if (mode == B200_MODE_EXACT) {
      if (epi == 2) return launch_stream_any<B200_MODE_EXACT, 2, false>(A, A->stream_grid, A->d_tiles, A->ntiles, x, b, y, none, st, dinv);
      return launch_stream_any<B200_MODE_EXACT, 3, false>(A, A->stream_grid, A->d_tiles, A->ntiles, x, b, y, none, st, dinv);
    }
    if (epi == 2) return launch_stream_any<B200_MODE_EXACT_FMA, 2, false>(A, A->stream_grid, A->d_tiles, A->ntiles, x, b, y, none, st, dinv);
    return launch_stream_any<B200_MODE_EXACT_FMA, 3, false>(A, A->stream_grid, A->d_tiles, A->ntiles, x, b, y, none, st, dinv);
  }
  B200_TRY(spmv_dispatch<false>(A, x, nullptr, y, mode, st));
  if (epi == 2) B200_LAUNCH((k_epilogue<2>), (A->m + 255) / 256, 256, 0, st, A->m, x, b, dinv, y);
  else B200_LAUNCH((k_epilogue<3>), (A->m + 255) / 256, 256, 0, st, A->m, x, b, dinv, y);
  return B200_OK;
}

// w = A x with (x, w) folded into the stream kernel's epilogue (EPI_DOT); other plans run their
// MatMult and then reduce in one CTA-sized pass (they serve small or skewed matrices).
namespace b200 {
int spmv_dot(b200_csr_t A, const double *x, double *y, int mode, const DotArgs &dot, cudaStream_t st)
{
  if (A->m == 0) { B200_CUDA_TRY(cudaMemsetAsync(dot.out, 0, sizeof(double), st)); return B200_OK; }
  int kernel = A->kernel_override;
  if (!kernel) kernel = (mode == B200_MODE_FAST) ? A->kernel_fast : A->kernel_exact;
  if (kernel == B200_KERNEL_STREAM && dot.partials) {
    const HaloArgs none{};
    if (mode == B200_MODE_EXACT) return launch_stream_any<B200_MODE_EXACT, EPI_DOT, false>(A, A->stream_grid, A->d_tiles, A->ntiles, x, nullptr, y, none, st, nullptr, dot);
    return launch_stream_any<B200_MODE_EXACT_FMA, EPI_DOT, false>(A, A->stream_grid, A->d_tiles, A->ntiles, x, nullptr, y, none, st, nullptr, dot);
  }
  B200_TRY(spmv_dispatch<false>(A, x, nullptr, y, mode, st));
  B200_LAUNCH(k_dot_simple, 1, 256, 0, st, A->m, x, (const double *)y, dot.out);
  return B200_OK;
}
}  // namespace b200

extern "C" int b200_spmv_residual(b200_csr_t A, const double *d_x, const double *d_b, double *d_r, int mode, void *stream)
{
  NvtxRange nvtx_("b200_spmv_residual");
  if (!A || ((!d_x && A->n) || ((!d_b || !d_r) && A->m))) return set_error(B200_ERR_ARG, "b200_spmv_residual: null argument");
  if (A->m && (d_x == d_r)) return set_error(B200_ERR_ARG, "b200_spmv_residual: x and r must differ");
  B200_TRY(check_mode(mode));
  return spmv_epilogue(A, 2, d_x, d_b, nullptr, d_r, mode, (cudaStream_t)stream);
}

extern "C" int b200_spmv_jacobi_sweep(b200_csr_t A, const double *d_x, const double *d_b, const double *d_dinv,
                                      double *d_xnew, int mode, void *stream)
{
  NvtxRange nvtx_("b200_spmv_jacobi_sweep");
  if (!A || A->m != A->n || (A->m && (!d_x || !d_b || !d_dinv || !d_xnew))) return set_error(B200_ERR_ARG, "b200_spmv_jacobi_sweep: bad argument (square matrix, non-null vectors)");
  if (A->m && d_x == d_xnew) return set_error(B200_ERR_ARG, "b200_spmv_jacobi_sweep: x and xnew must differ (other rows still read x)");
  B200_TRY(check_mode(mode));
  return spmv_epilogue(A, 3, d_x, d_b, d_dinv, d_xnew, mode, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// transpose: explicit copy (deterministic, ascending-row order per column = the reference's
// scatter order) or atomics.
// ---------------------------------------------------------------------------------------------
// Device build: histogram of the columns -> exclusive scan = row pointers of A^T; stable radix sort
// of the non-zero positions by column = the permutation; gather rows and values through it.
static int build_transpose_device(b200_csr_s *A, b200_csr_t *out)
{
  const int m = A->m, n = A->n, nz = A->nz;
  int    *rowid = nullptr, *cnt = nullptr, *ti = nullptr, *keys_out = nullptr, *perm_in = nullptr, *perm = nullptr, *tj = nullptr;
  double *ta = nullptr;
  void   *tmp = nullptr;
  auto cleanup = [&]() { cudaFree(rowid); cudaFree(cnt); cudaFree(ti); cudaFree(keys_out); cudaFree(perm_in); cudaFree(perm); cudaFree(tj); cudaFree(ta); cudaFree(tmp); };
  auto body = [&]() -> int {
    const size_t z = std::max(nz, 1);
    B200_CUDA_TRY(cudaMalloc((void **)&rowid, z * sizeof(int)));
    B200_CUDA_TRY(cudaMalloc((void **)&cnt, ((size_t)n + 1) * sizeof(int)));
    B200_CUDA_TRY(cudaMalloc((void **)&ti, ((size_t)n + 1) * sizeof(int)));
    B200_CUDA_TRY(cudaMalloc((void **)&keys_out, z * sizeof(int)));
    B200_CUDA_TRY(cudaMalloc((void **)&perm_in, z * sizeof(int)));
    B200_CUDA_TRY(cudaMalloc((void **)&perm, z * sizeof(int)));
    B200_CUDA_TRY(cudaMalloc((void **)&tj, z * sizeof(int)));
    B200_CUDA_TRY(cudaMalloc((void **)&ta, z * sizeof(double)));
    B200_CUDA_TRY(cudaMemset(cnt, 0, ((size_t)n + 1) * sizeof(int)));
    if (m) B200_LAUNCH(k_tr_expand, (m + 127) / 128, 128, 0, 0, m, A->d_ai, rowid, A->d_aj, cnt);
    size_t bytes = 0;
    B200_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt, ti, n + 1));
    B200_CUDA_TRY(cudaMalloc(&tmp, std::max<size_t>(bytes, 1)));
    B200_CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, bytes, cnt, ti, n + 1));
    g_launches.fetch_add(1);
    cudaFree(tmp); tmp = nullptr;
    if (nz) {
      B200_LAUNCH(k_iota, (nz + 255) / 256, 256, 0, 0, nz, perm_in);
      int bits = 1;
      while (bits < 31 && (1 << bits) < std::max(n, 2)) ++bits;
      B200_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, A->d_aj, keys_out, perm_in, perm, nz, 0, bits));
      B200_CUDA_TRY(cudaMalloc(&tmp, std::max<size_t>(bytes, 1)));
      B200_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, bytes, A->d_aj, keys_out, perm_in, perm, nz, 0, bits));
      g_launches.fetch_add(1);
      B200_LAUNCH(k_tr_gather, (nz + 255) / 256, 256, 0, 0, nz, perm, rowid, A->d_aa, tj, ta);
    }
    B200_CUDA_TRY(cudaDeviceSynchronize());
    return b200_csr_create_from_device(out, n, m, ti, tj, ta);
  };
  const int rc = body();
  cleanup();
  return rc;
}

extern "C" int b200_csr_build_transpose(b200_csr_t A)
{
  NvtxRange nvtx_("b200_csr_build_transpose");
  if (!A) return set_error(B200_ERR_ARG, "null handle");
  if (A->T) return B200_OK;
  if (!env_int("B200_TRANSPOSE_HOST", 0)) {
    b200_csr_t T = nullptr;
    B200_TRY(build_transpose_device(A, &T));
    A->T = T;
    return B200_OK;
  }
  const int m = A->m, n = A->n, nz = A->nz;
  std::vector<int>    ai((size_t)m + 1), aj((size_t)nz), ti((size_t)n + 1, 0), tj((size_t)nz);
  std::vector<double> aa((size_t)nz), ta((size_t)nz);
  B200_CUDA_TRY(cudaMemcpy(ai.data(), A->d_ai, ai.size() * sizeof(int), cudaMemcpyDeviceToHost));
  if (nz) {
    B200_CUDA_TRY(cudaMemcpy(aj.data(), A->d_aj, (size_t)nz * sizeof(int), cudaMemcpyDeviceToHost));
    B200_CUDA_TRY(cudaMemcpy(aa.data(), A->d_aa, (size_t)nz * sizeof(double), cudaMemcpyDeviceToHost));
  }
  // counting sort by column; rows are visited in ascending order so each column keeps it
  for (int k = 0; k < nz; ++k) {
    if (aj[k] < 0 || aj[k] >= n) return set_error(B200_ERR_ARG, "column index %d out of range [0,%d)", aj[k], n);
    ti[aj[k] + 1]++;
  }
  for (int c = 0; c < n; ++c) ti[c + 1] += ti[c];
  {
    std::vector<int> next(ti.begin(), ti.end() - 1);
    for (int i = 0; i < m; ++i)
      for (int k = ai[i]; k < ai[i + 1]; ++k) {
        int p = next[aj[k]]++;
        tj[p] = i;
        ta[p] = aa[k];
      }
  }
  b200_csr_t T = nullptr;
  B200_TRY(b200_csr_create(&T, n, m, ti.data(), tj.data(), ta.data()));
  A->T = T;
  return B200_OK;
}

static int transpose_common(b200_csr_s *A, const double *d_x, const double *d_z, double *d_y,
                            int mode, cudaStream_t st)
{
  B200_TRY(check_mode(mode));
  if (A->n == 0) return B200_OK;
  const bool atomic = (mode == B200_MODE_FAST) && !A->T && env_int("B200_TRANSPOSE_ATOMIC", 0);
  if (!atomic) {
    B200_TRY(b200_csr_build_transpose(A));
    if (d_z) return spmv_dispatch<true>(A->T, d_x, d_z, d_y, mode, st);
    return spmv_dispatch<false>(A->T, d_x, nullptr, d_y, mode, st);
  }
  if (d_z) {
    if (d_z != d_y)
      B200_LAUNCH(k_copy, std::max(1, std::min(sm_count() * 8, (A->n + 255) / 256)), 256, 0, st, d_y, d_z, (long long)A->n);
  } else {
    B200_CUDA_TRY(cudaMemsetAsync(d_y, 0, (size_t)A->n * sizeof(double), st));
  }
  if (A->m) {
    const long long thr = (long long)A->m * 4;
    B200_LAUNCH((k_tr_atomic<4>), (unsigned)((thr + 255) / 256), 256, 0, st, A->m, A->d_ai, A->d_aj, A->d_aa, d_x, d_y);
  }
  return B200_OK;
}

extern "C" int b200_spmv_transpose(b200_csr_t A, const double *d_x, double *d_y, int mode, void *stream)
{
  NvtxRange nvtx_("b200_spmv_transpose");
  if (!A || (!d_x && A->m) || (!d_y && A->n)) return set_error(B200_ERR_ARG, "b200_spmv_transpose: null argument");
  return transpose_common(A, d_x, nullptr, d_y, mode, (cudaStream_t)stream);
}

extern "C" int b200_spmv_transpose_add(b200_csr_t A, const double *d_x, const double *d_z,
                                       double *d_y, int mode, void *stream)
{
  NvtxRange nvtx_("b200_spmv_transpose_add");
  if (!A || (!d_x && A->m) || ((!d_y || !d_z) && A->n)) return set_error(B200_ERR_ARG, "b200_spmv_transpose_add: null argument");
  return transpose_common(A, d_x, d_z, d_y, mode, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// host-vector entry points: what MatMult_SeqAIJ(Mat,Vec,Vec) does with PETSc 3.7.6 host Vecs.
// ---------------------------------------------------------------------------------------------
static int host_scratch(b200_csr_s *A, size_t nx, size_t ny)
{
  if (A->hx_len < nx) {
    cudaFree(A->d_hx); A->d_hx = nullptr;
    B200_CUDA_TRY(cudaMalloc((void **)&A->d_hx, std::max<size_t>(nx, 1) * sizeof(double)));
    A->hx_len = nx;
  }
  if (A->hy_len < ny) {
    cudaFree(A->d_hy); A->d_hy = nullptr;
    B200_CUDA_TRY(cudaMalloc((void **)&A->d_hy, std::max<size_t>(ny, 1) * sizeof(double)));
    A->hy_len = ny;
  }
  for (auto &s : A->hs) if (!s) B200_CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  for (auto &e : A->hev) if (!e) B200_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return B200_OK;
}

// Row-blocked pipeline: x chunks go up on stream 0, block b's tiles run on stream 1 as soon as
// the last chunk it reads has landed, its y rows go down on stream 2 -- uploads, kernels and
// downloads of different blocks overlap (PCIe is full duplex).  Banded matrices only need a
// look-ahead of the bandwidth; for others need[b] is the last chunk and only the downloads overlap.
static int spmv_host_pipelined(b200_csr_s *A, const double *h_x, const double *h_yin, double *h_y, int mode)
{
  const int nblk = (int)A->pb_tile.size() - 1;
  if ((int)A->pb_evx.size() < nblk) {
    const size_t old = A->pb_evx.size();
    A->pb_evx.resize(nblk); A->pb_evk.resize(nblk);
    for (size_t b = old; b < (size_t)nblk; ++b) {
      B200_CUDA_TRY(cudaEventCreateWithFlags(&A->pb_evx[b], cudaEventDisableTiming));
      B200_CUDA_TRY(cudaEventCreateWithFlags(&A->pb_evk[b], cudaEventDisableTiming));
    }
  }
  auto r0 = [&](int b) { return A->h_tiles[A->pb_tile[b]].x; };
  auto r1 = [&](int b) { return A->h_tiles[A->pb_tile[b + 1] - 1].y; };
  for (int b = 0; b < nblk; ++b) {
    const size_t off = r0(b), len = (size_t)r1(b) - r0(b);
    B200_CUDA_TRY(cudaMemcpyAsync(A->d_hx + off, h_x + off, len * sizeof(double), cudaMemcpyHostToDevice, A->hs[0]));
    if (h_yin) B200_CUDA_TRY(cudaMemcpyAsync(A->d_hy + off, h_yin + off, len * sizeof(double), cudaMemcpyHostToDevice, A->hs[0]));
    B200_CUDA_TRY(cudaEventRecord(A->pb_evx[b], A->hs[0]));
  }
  for (int b = 0; b < nblk; ++b) {
    B200_CUDA_TRY(cudaStreamWaitEvent(A->hs[1], A->pb_evx[std::max(A->pb_need[b], b)], 0));
    const int t0 = A->pb_tile[b], nt = A->pb_tile[b + 1] - t0;
    if (h_yin) {
      if (mode == B200_MODE_EXACT) B200_TRY((launch_stream_range<B200_MODE_EXACT, true>(A, t0, nt, A->d_hx, A->d_hy, A->d_hy, A->hs[1])));
      else B200_TRY((launch_stream_range<B200_MODE_EXACT_FMA, true>(A, t0, nt, A->d_hx, A->d_hy, A->d_hy, A->hs[1])));
    } else {
      if (mode == B200_MODE_EXACT) B200_TRY((launch_stream_range<B200_MODE_EXACT, false>(A, t0, nt, A->d_hx, nullptr, A->d_hy, A->hs[1])));
      else B200_TRY((launch_stream_range<B200_MODE_EXACT_FMA, false>(A, t0, nt, A->d_hx, nullptr, A->d_hy, A->hs[1])));
    }
    B200_CUDA_TRY(cudaEventRecord(A->pb_evk[b], A->hs[1]));
    B200_CUDA_TRY(cudaStreamWaitEvent(A->hs[2], A->pb_evk[b], 0));
    const size_t off = r0(b), len = (size_t)r1(b) - r0(b);
    B200_CUDA_TRY(cudaMemcpyAsync(h_y + off, A->d_hy + off, len * sizeof(double), cudaMemcpyDeviceToHost, A->hs[2]));
  }
  B200_CUDA_TRY(cudaStreamSynchronize(A->hs[2]));
  B200_CUDA_TRY(cudaStreamSynchronize(A->hs[0]));
  return B200_OK;
}

static bool host_pipeline_applies(b200_csr_s *A, int mode)
{
  if (A->pb_tile.size() < 3 || !env_int("B200_HOST_PIPELINE", 1)) return false;
  const int kernel = A->kernel_override ? A->kernel_override : (mode == B200_MODE_FAST ? A->kernel_fast : A->kernel_exact);
  return kernel == B200_KERNEL_STREAM;
}

namespace b200 {
int host_block_count(b200_csr_t A, int mode)
{
  return host_pipeline_applies(A, mode) ? (int)A->pb_tile.size() - 1 : 0;
}
int host_block_info(b200_csr_t A, int b, int *row0, int *row1, int *need)
{
  *row0 = A->h_tiles[A->pb_tile[b]].x;
  *row1 = A->h_tiles[A->pb_tile[b + 1] - 1].y;
  *need = std::max(A->pb_need[b], b);
  return B200_OK;
}
int launch_host_block(b200_csr_t A, int b, const double *x, double *y, int mode, cudaStream_t st)
{
  const int t0 = A->pb_tile[b], nt = A->pb_tile[b + 1] - t0;
  if (mode == B200_MODE_EXACT) return launch_stream_range<B200_MODE_EXACT, false>(A, t0, nt, x, nullptr, y, st);
  return launch_stream_range<B200_MODE_EXACT_FMA, false>(A, t0, nt, x, nullptr, y, st);
}
}  // namespace b200

extern "C" int b200_spmv_host(b200_csr_t A, const double *h_x, double *h_y, int mode)
{
  NvtxRange nvtx_("b200_spmv_host");
  if (!A || (!h_x && A->n) || (!h_y && A->m)) return set_error(B200_ERR_ARG, "b200_spmv_host: null argument");
  B200_TRY(check_mode(mode));
  B200_TRY(host_scratch(A, A->n, A->m));
  if (host_pipeline_applies(A, mode)) return spmv_host_pipelined(A, h_x, nullptr, h_y, mode);
  cudaStream_t s = A->hs[0];
  B200_CUDA_TRY(cudaMemcpyAsync(A->d_hx, h_x, (size_t)A->n * sizeof(double), cudaMemcpyHostToDevice, s));
  B200_TRY(spmv_dispatch<false>(A, A->d_hx, nullptr, A->d_hy, mode, s));
  B200_CUDA_TRY(cudaMemcpyAsync(h_y, A->d_hy, (size_t)A->m * sizeof(double), cudaMemcpyDeviceToHost, s));
  B200_CUDA_TRY(cudaStreamSynchronize(s));
  return B200_OK;
}

extern "C" int b200_spmv_add_host(b200_csr_t A, const double *h_x, const double *h_y, double *h_z, int mode)
{
  if (!A || (!h_x && A->n) || ((!h_y || !h_z) && A->m)) return set_error(B200_ERR_ARG, "b200_spmv_add_host: null argument");
  B200_TRY(check_mode(mode));
  B200_TRY(host_scratch(A, A->n, A->m));
  if (host_pipeline_applies(A, mode)) return spmv_host_pipelined(A, h_x, h_y, h_z, mode);
  cudaStream_t s = A->hs[0];
  B200_CUDA_TRY(cudaMemcpyAsync(A->d_hx, h_x, (size_t)A->n * sizeof(double), cudaMemcpyHostToDevice, s));
  B200_CUDA_TRY(cudaMemcpyAsync(A->d_hy, h_y, (size_t)A->m * sizeof(double), cudaMemcpyHostToDevice, s));
  B200_TRY(spmv_dispatch<true>(A, A->d_hx, A->d_hy, A->d_hy, mode, s));
  B200_CUDA_TRY(cudaMemcpyAsync(h_z, A->d_hy, (size_t)A->m * sizeof(double), cudaMemcpyDeviceToHost, s));
  B200_CUDA_TRY(cudaStreamSynchronize(s));
  return B200_OK;
}

extern "C" int b200_spmv_transpose_host(b200_csr_t A, const double *h_x, double *h_y, int mode)
{
  if (!A || (!h_x && A->m) || (!h_y && A->n)) return set_error(B200_ERR_ARG, "b200_spmv_transpose_host: null argument");
  B200_TRY(check_mode(mode));
  B200_TRY(host_scratch(A, A->m, A->n));
  cudaStream_t s = A->hs[0];
  B200_CUDA_TRY(cudaMemcpyAsync(A->d_hx, h_x, (size_t)A->m * sizeof(double), cudaMemcpyHostToDevice, s));
  B200_TRY(transpose_common(A, A->d_hx, nullptr, A->d_hy, mode, s));
  B200_CUDA_TRY(cudaMemcpyAsync(h_y, A->d_hy, (size_t)A->n * sizeof(double), cudaMemcpyDeviceToHost, s));
  B200_CUDA_TRY(cudaStreamSynchronize(s));
  return B200_OK;
}

extern "C" int b200_spmv_transpose_add_host(b200_csr_t A, const double *h_x, const double *h_z, double *h_y, int mode)
{
  if (!A || (!h_x && A->m) || ((!h_y || !h_z) && A->n)) return set_error(B200_ERR_ARG, "b200_spmv_transpose_add_host: null argument");
  B200_TRY(check_mode(mode));
  B200_TRY(host_scratch(A, A->m, A->n));
  cudaStream_t s = A->hs[0];
  B200_CUDA_TRY(cudaMemcpyAsync(A->d_hx, h_x, (size_t)A->m * sizeof(double), cudaMemcpyHostToDevice, s));
  B200_CUDA_TRY(cudaMemcpyAsync(A->d_hy, h_z, (size_t)A->n * sizeof(double), cudaMemcpyHostToDevice, s));
  B200_TRY(transpose_common(A, A->d_hx, A->d_hy, A->d_hy, mode, s));
  B200_CUDA_TRY(cudaMemcpyAsync(h_y, A->d_hy, (size_t)A->n * sizeof(double), cudaMemcpyDeviceToHost, s));
  B200_CUDA_TRY(cudaStreamSynchronize(s));
  return B200_OK;
}

extern "C" int b200_host_alloc(void **p, size_t bytes)
{
  if (!p) return set_error(B200_ERR_ARG, "null pointer");
  B200_TRY(ensure_device());
  B200_CUDA_TRY(cudaHostAlloc(p, std::max<size_t>(bytes, 1), cudaHostAllocDefault));
  return B200_OK;
}
extern "C" int b200_host_free(void *p)
{
  if (p) B200_CUDA_TRY(cudaFreeHost(p));
  return B200_OK;
}
extern "C" int b200_host_register(void *p, size_t bytes)
{
  B200_TRY(ensure_device());
  B200_CUDA_TRY(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
  return B200_OK;
}
extern "C" int b200_host_unregister(void *p)
{
  B200_CUDA_TRY(cudaHostUnregister(p));
  return B200_OK;
}

// splitmix64 -> uniform [-1,1): the synthetic x of SURVEY 8(d)
extern "C" int b200_gen_vector(double *h_x, int64_t n, uint64_t seed)
{
  if (!h_x && n) return set_error(B200_ERR_ARG, "null pointer");
  for (int64_t i = 0; i < n; ++i) {
    uint64_t z = (seed ^ (uint64_t)i) + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    h_x[i] = (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
  }
  return B200_OK;
}
