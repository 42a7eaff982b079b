// b200_halo.cu -- MatMult_MPIAIJ for one box of B200s: A/B split, garray, scatter lists (host,
// integer work bit-exact with MatSetUpMultiply_MPIAIJ's result) and a push-based NVLink halo.
//
// What it replaces [P376, un-vendored]: mpiaij.c MatMult_MPIAIJ, mmaij.c
// MatSetUpMultiply_MPIAIJ, vscat.c VecScatterBegin/End (MPI persistent send/recv of HOST buffers;
// SURVEY 3.4, 5.9).  Here rank r's pack kernel stores boundary x values directly into the peer's
// lvec through a CUDA-IPC mapping (NVLink 5 / NVSwitch), fences, and releases a per-source flag
// in the peer's window; the peer's off-diagonal kernel acquires the flags and adds B*lvec.
//
// Window layout (one cudaMalloc per rank, exported with cudaIpcGetMemHandle):
//   [0, 1024)                      uint64 flags[size<=120] ; word 126 = error
//   [1024, 8704)                   all-reduce slots red[2][120] = {v0, v1, v2, seq}
//   [16384, 16384 + 8*ngpad)       lvec buffer 0
//   [.. + 8*ngpad, .. + 16*ngpad)  lvec buffer 1          (ngpad = nghost rounded up to 16)
// Buffers alternate with the MatMult sequence number, which makes the exchange safe without a
// reverse "buffer free" signal: a peer can only be one MatMult ahead of this rank.
#include <algorithm>
#include <cstring>
#include <new>
#include <queue>
#include <vector>

#include "../../include/b200_mpiaij.h"
#include "b200_common.h"

using namespace b200;

namespace {

constexpr int    WINDOW_HDR_BYTES = 16384;  // 1 KB of flags + the all-reduce slots
constexpr int    RED_SLOT_OFFSET  = 1024;   // red[parity][source rank] = {v0, v1, v2, seq}: 2 x 120 x 32 B
constexpr int    RED_MAX_VALS     = 3;
constexpr int    MAX_RANKS        = 120;
constexpr int    ERR_WORD         = 126;
constexpr int    PUSH_CHUNK       = 2304;  // elements per CTA of the push role (8 per thread of a 288-thread CTA)

// VecScatterBegin: gather x[send_idx] and store it into each peer's lvec (peer memory), then the
// last CTA of each peer releases flag = seq.
__global__ void __launch_bounds__(256)
    k_halo_push(const PushBlock *__restrict__ blocks, const PushPeer *__restrict__ peers,
                const int *__restrict__ send_idx, const double *__restrict__ x, unsigned *done,
                unsigned long long seq)
{
  HaloArgs h{};
  h.blocks = blocks; h.peers = peers; h.send_idx = send_idx; h.done = done; h.seq = seq;
  halo_push_stores<0>(h, x, blockIdx.x);
  halo_push_signal<0>(h, blockIdx.x);
}

// pack only (transport owned by the caller)
__global__ void k_halo_pack(int count, const int *__restrict__ send_idx, const double *__restrict__ x, double *buf)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count) buf[t] = __ldg(x + send_idx[t]);
}

// VecScatterEnd + MatMultAdd_SeqAIJ(B, lvec, y, y), compressed-row: thread per non-empty row of B.
// WAIT: thread 0 of every CTA acquires the source flags (>= seq) before the CTA reads lvec.
template <int MODE, bool WAIT>
__global__ void __launch_bounds__(128)
    k_offdiag(int nrows, const int *__restrict__ cpi, const int *__restrict__ ridx,
              const int *__restrict__ bj, const double *__restrict__ ba, const double *lvec,
              double *y, const unsigned long long *flags, const int *__restrict__ srcs, int nsrc,
              unsigned long long seq, unsigned long long *err, unsigned long long timeout_ns)
{
  if (WAIT) {
    if (threadIdx.x == 0) {
      HaloArgs h{};
      h.flags = flags; h.srcs = srcs; h.nsrc = nsrc; h.seq = seq; h.err = err; h.timeout_ns = timeout_ns;
      halo_wait_flags(h);
    }
    __syncthreads();
  }
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nrows) return;
  const int lo = cpi[t], hi = cpi[t + 1], i = ridx[t];
  double    sum = y[i];
  for (int k = lo; k < hi; ++k) {
    const double xv = __ldcg(lvec + bj[k]);  // written by a peer GPU: read through L2, not L1
    sum = (MODE == B200_MODE_EXACT) ? __dadd_rn(sum, __dmul_rn(ba[k], xv)) : __fma_rn(ba[k], xv, sum);
  }
  y[i] = sum;
}

// All-reduce (sum) of up to 3 device scalars over the ranks, through the same peer windows:
// rank r stores {values, seq} into slot [seq & 1][r] of EVERY rank's window (its own included),
// then reads all slots of its own window and adds them in rank order -- every rank gets the same
// bits.  Slots alternate with the sequence number; a rank can be at most one all-reduce ahead of
// a peer, so two slots per source are enough.  One CTA, one lane per rank.
// CG rides on it: `skip` (the solve's DONE word, identical on every rank because it is computed from
// all-reduced values) turns the launch into a no-op, and `post` runs the scalar step of the
// iteration (b200_common.h) on the reduced values -- one launch instead of all-reduce + scalar kernel.
struct WinSlot { double v[3]; unsigned long long seq; };
__global__ void __launch_bounds__(128)
    k_allreduce(double *vals, int nvals, int size, int rank, unsigned char *const *windows,
                unsigned long long seq, unsigned long long *err, unsigned long long timeout_ns,
                const int *skip, int post, double *sc, int *st)
{
  pdl_wait();
  pdl_launch_dependents();
  if (skip && *skip) return;
  __shared__ double sv[MAX_RANKS][RED_MAX_VALS];
  const int t = threadIdx.x;
  if (t < size) {
    WinSlot *dst = reinterpret_cast<WinSlot *>(windows[t] + RED_SLOT_OFFSET) + (seq & 1) * MAX_RANKS + rank;
    for (int k = 0; k < nvals; ++k) dst->v[k] = vals[k];
    __threadfence_system();
    st_release_sys(&dst->seq, seq);
    const WinSlot *src = reinterpret_cast<const WinSlot *>(windows[rank] + RED_SLOT_OFFSET) + (seq & 1) * MAX_RANKS + t;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_relaxed_sys(&src->seq) < seq) {
      if (globaltimer_ns() - t0 > timeout_ns) { atomicExch(err, 1ull); break; }
      __nanosleep(32);
    }
    __threadfence_system();
    for (int k = 0; k < nvals; ++k) sv[t][k] = __ldcg(&src->v[k]);
  }
  __syncthreads();
  if (t < nvals) {
    double s = 0.0;
    for (int q = 0; q < size; ++q) s += sv[q][t];
    vals[t] = s;
  }
  if (post) {
    __syncthreads();
    if (t == 0) cg_scalar_post(post, sc, st);
  }
}

}  // namespace

struct b200_mpiaij_s {
  int32_t size = 1, rank = 0, nloc = 0, rstart = 0, rend = 0;
  std::vector<int32_t> base;
  // host blocks
  std::vector<int32_t> Ai, Aj, Bi, Bj, garray, recv_off, cpi, ridx, srcs;
  std::vector<double>  Aa, Ba;
  // send side
  struct Send { int32_t peer, count, peer_offset, peer_ngpad; std::vector<int32_t> idx; void *peer_window = nullptr; bool opened = false; };
  std::vector<Send> sends;  // one per peer that needs something from this rank
  // device
  bool      uploaded = false, push_ready = false;
  b200_csr_t A = nullptr, B = nullptr;
  int       *d_cpi = nullptr, *d_ridx = nullptr, *d_bj = nullptr, *d_srcs = nullptr, *d_send_idx = nullptr;
  double    *d_ba = nullptr;
  unsigned char *d_window = nullptr;
  size_t     window_bytes = 0;
  int32_t    ngpad = 0;
  PushBlock *d_blocks = nullptr;
  PushPeer  *d_peers = nullptr;
  unsigned  *d_done = nullptr;
  int        npush_blocks = 0;
  std::vector<int32_t> send_start;  // start of each peer's run in d_send_idx
  unsigned long long seq = 0;
  cudaStream_t side = nullptr, hstream = nullptr;
  double *d_hx = nullptr, *d_hy = nullptr;
  unsigned long long *h_err = nullptr;   // pinned copy of the time-out word (host-vector entry)
  // row-blocked pipeline of the host-vector entry
  cudaStream_t hs_up = nullptr, hs_dn = nullptr, hs_push = nullptr;
  cudaEvent_t  ev_push = nullptr;
  std::vector<cudaEvent_t> evx, evk;
  double *d_gbuf = nullptr, *h_gbuf = nullptr;
  cudaEvent_t  ev_fork = nullptr, ev_join = nullptr;
  unsigned long long timeout_ns = 10000ull * 1000000ull;
  int *d_cta_ptr = nullptr, *d_cta_rows = nullptr;  // fused launch: B rows grouped by owning CTA
  int4 *d_sched_tiles = nullptr;                    // fused launch: the tile table in CTA-major order
  int  *d_sched_first = nullptr;                    //               and each CTA's range in it
  // all-reduce: every rank's window (own included), in rank order
  std::vector<unsigned char *> all_windows;
  std::vector<char> all_windows_opened;
  unsigned char **d_all_windows = nullptr;
  unsigned long long rseq = 0;
  int   fused_grid = 0;
  bool  fused_ok = false;
};

static int32_t pad16(int32_t n) { return (n + 15) & ~15; }

extern "C" int b200_mpiaij_create(b200_mpiaij_t *out, int32_t size, int32_t rank,
                                  const int32_t *base, const int32_t *h_ai,
                                  const int32_t *h_aj, const double *h_aa)
{
  if (!out || size < 1 || size > MAX_RANKS || rank < 0 || rank >= size || !base || !h_ai)
    return set_error(B200_ERR_ARG, "b200_mpiaij_create: bad argument (size must be 1..%d)", MAX_RANKS);
  b200_mpiaij_s *M = new (std::nothrow) b200_mpiaij_s;
  if (!M) return set_error(B200_ERR_MEM, "out of host memory");
  M->size = size; M->rank = rank;
  M->base.assign(base, base + size + 1);
  M->rstart = base[rank]; M->rend = base[rank + 1];
  M->nloc = M->rend - M->rstart;
  const int nloc = M->nloc, nz = h_ai[nloc];
  if (nz && (!h_aj || !h_aa)) { delete M; return set_error(B200_ERR_ARG, "null aj/aa"); }
  // --- MatSetValues_MPIAIJ split [P376]: owned columns -> A (local ids), others -> B ----------
  M->Ai.assign(nloc + 1, 0); M->Bi.assign(nloc + 1, 0);
  M->Aj.reserve(nz); M->Aa.reserve(nz);
  for (int i = 0; i < nloc; ++i) {
    for (int k = h_ai[i]; k < h_ai[i + 1]; ++k) {
      const int c = h_aj[k];
      if (c < 0 || c >= base[size]) { delete M; return set_error(B200_ERR_ARG, "column %d out of range", c); }
      if (c >= M->rstart && c < M->rend) { M->Aj.push_back(c - M->rstart); M->Aa.push_back(h_aa[k]); }
      else { M->Bj.push_back(c); M->Ba.push_back(h_aa[k]); }
    }
    M->Ai[i + 1] = (int32_t)M->Aj.size();
    M->Bi[i + 1] = (int32_t)M->Bj.size();
  }
  // --- MatSetUpMultiply_MPIAIJ [P376]: garray = sorted unique ghost ids; compact B's columns ---
  M->garray = M->Bj;
  std::sort(M->garray.begin(), M->garray.end());
  M->garray.erase(std::unique(M->garray.begin(), M->garray.end()), M->garray.end());
  for (auto &c : M->Bj) c = (int32_t)(std::lower_bound(M->garray.begin(), M->garray.end(), c) - M->garray.begin());
  // --- receive side of Mvctx: contiguous run of garray per owner ------------------------------
  M->recv_off.assign(size + 1, 0);
  {
    size_t k = 0;
    for (int q = 0; q < size; ++q) {
      M->recv_off[q] = (int32_t)k;
      while (k < M->garray.size() && M->garray[k] < base[q + 1]) ++k;
    }
    M->recv_off[size] = (int32_t)k;
  }
  for (int q = 0; q < size; ++q)
    if (q != rank && M->recv_off[q + 1] > M->recv_off[q]) M->srcs.push_back(q);
  // --- compressed-row index of B (only the rows that touch a ghost) ---------------------------
  M->cpi.push_back(0);
  for (int i = 0; i < nloc; ++i)
    if (M->Bi[i + 1] > M->Bi[i]) { M->cpi.push_back(M->Bi[i + 1]); M->ridx.push_back(i); }
  M->ngpad = pad16((int32_t)M->garray.size());
  *out = M;
  return B200_OK;
}

extern "C" int b200_mpiaij_destroy(b200_mpiaij_t M)
{
  if (!M) return B200_OK;
  for (auto &s : M->sends) if (s.opened && s.peer_window) cudaIpcCloseMemHandle(s.peer_window);
  if (M->A) b200_csr_destroy(M->A);
  if (M->B) b200_csr_destroy(M->B);
  cudaFree(M->d_cpi); cudaFree(M->d_ridx); cudaFree(M->d_bj); cudaFree(M->d_ba); cudaFree(M->d_srcs);
  cudaFree(M->d_cta_ptr); cudaFree(M->d_cta_rows); cudaFree(M->d_sched_tiles); cudaFree(M->d_sched_first);
  cudaFree(M->d_all_windows);
  for (int q = 0; q < (int)M->all_windows.size(); ++q)
    if (q != M->rank && M->all_windows[q] && M->all_windows_opened[q]) cudaIpcCloseMemHandle(M->all_windows[q]);
  cudaFree(M->d_send_idx); cudaFree(M->d_window); cudaFree(M->d_blocks); cudaFree(M->d_peers); cudaFree(M->d_done);
  if (M->side) cudaStreamDestroy(M->side);
  if (M->hstream) cudaStreamDestroy(M->hstream);
  cudaFree(M->d_hx); cudaFree(M->d_hy);
  if (M->h_err) cudaFreeHost(M->h_err);
  if (M->hs_up) { cudaStreamDestroy(M->hs_up); cudaStreamDestroy(M->hs_dn); cudaStreamDestroy(M->hs_push); cudaEventDestroy(M->ev_push); }
  for (auto &e : M->evx) cudaEventDestroy(e);
  for (auto &e : M->evk) cudaEventDestroy(e);
  cudaFree(M->d_gbuf);
  if (M->h_gbuf) cudaFreeHost(M->h_gbuf);
  if (M->ev_fork) cudaEventDestroy(M->ev_fork);
  if (M->ev_join) cudaEventDestroy(M->ev_join);
  delete M;
  return B200_OK;
}

extern "C" int b200_mpiaij_get_sizes(b200_mpiaij_t M, int32_t *s)
{
  if (!M || !s) return set_error(B200_ERR_ARG, "null argument");
  s[0] = M->nloc; s[1] = (int32_t)M->Aj.size(); s[2] = (int32_t)M->Bj.size();
  s[3] = (int32_t)M->garray.size(); s[4] = (int32_t)M->ridx.size(); s[5] = (int32_t)M->srcs.size();
  return B200_OK;
}
extern "C" int b200_mpiaij_get_garray(b200_mpiaij_t M, int32_t *g)
{
  if (!M || (!g && !M->garray.empty())) return set_error(B200_ERR_ARG, "null argument");
  if (!M->garray.empty()) memcpy(g, M->garray.data(), M->garray.size() * sizeof(int32_t));
  return B200_OK;
}
extern "C" int b200_mpiaij_get_recv_offsets(b200_mpiaij_t M, int32_t *off)
{
  if (!M || !off) return set_error(B200_ERR_ARG, "null argument");
  memcpy(off, M->recv_off.data(), M->recv_off.size() * sizeof(int32_t));
  return B200_OK;
}
extern "C" int b200_mpiaij_copy_block(b200_mpiaij_t M, int which, int32_t *ai, int32_t *aj, double *aa)
{
  if (!M || which < 0 || which > 1) return set_error(B200_ERR_ARG, "bad argument");
  const auto &I = which ? M->Bi : M->Ai;
  const auto &J = which ? M->Bj : M->Aj;
  const auto &V = which ? M->Ba : M->Aa;
  if (ai) memcpy(ai, I.data(), I.size() * sizeof(int32_t));
  if (aj && !J.empty()) memcpy(aj, J.data(), J.size() * sizeof(int32_t));
  if (aa && !V.empty()) memcpy(aa, V.data(), V.size() * sizeof(double));
  return B200_OK;
}

extern "C" int b200_mpiaij_set_peer_garray(b200_mpiaij_t M, int32_t peer, const int32_t *pg, int32_t png)
{
  if (!M || peer < 0 || peer >= M->size || png < 0 || (png && !pg)) return set_error(B200_ERR_ARG, "bad argument");
  if (peer == M->rank) return B200_OK;
  if (M->push_ready) return set_error(B200_ERR_STATE, "send lists are frozen after the first MatMult");
  const int32_t *lo = std::lower_bound(pg, pg + png, M->rstart);
  const int32_t *hi = std::lower_bound(pg, pg + png, M->rend);
  for (auto it = M->sends.begin(); it != M->sends.end(); ++it)
    if (it->peer == peer) { M->sends.erase(it); break; }
  if (hi == lo) return B200_OK;
  b200_mpiaij_s::Send s;
  s.peer = peer; s.count = (int32_t)(hi - lo); s.peer_offset = (int32_t)(lo - pg); s.peer_ngpad = pad16(png);
  s.idx.resize(s.count);
  for (int32_t t = 0; t < s.count; ++t) s.idx[t] = lo[t] - M->rstart;
  M->sends.push_back(std::move(s));
  std::sort(M->sends.begin(), M->sends.end(), [](const b200_mpiaij_s::Send &a, const b200_mpiaij_s::Send &b) { return a.peer < b.peer; });
  return B200_OK;
}

extern "C" int b200_mpiaij_get_send_list(b200_mpiaij_t M, int32_t peer, int32_t *count, int32_t *idx, int32_t *peer_offset)
{
  if (!M || !count) return set_error(B200_ERR_ARG, "null argument");
  *count = 0;
  for (auto &s : M->sends)
    if (s.peer == peer) {
      *count = s.count;
      if (peer_offset) *peer_offset = s.peer_offset;
      if (idx) memcpy(idx, s.idx.data(), s.idx.size() * sizeof(int32_t));
    }
  return B200_OK;
}

template <typename T>
static int up(T **d, const std::vector<T> &h)
{
  B200_CUDA_TRY(cudaMalloc((void **)d, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) B200_CUDA_TRY(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return B200_OK;
}

// Tile -> CTA of the fused launch.  Round-robin keeps all CTAs sweeping the matrix in lockstep (the x
// window they gather from stays in L2), but a face of the sub-box that is contiguous in memory puts
// 256 ghost rows into every one of ~90 consecutive tiles, and a CTA that owns one of them would close
// those rows on top of a full share of tiles and finish last; the first `npush` CTAs also carry a
// push block (VecScatterBegin's stores and the fence behind them).  So: the push CTAs start with
// `push_charge` tile times on their account (default 6.0: 8 GPUs, 32 push CTAs: 46.5 us per MatMult at
// 4.0, 45.2 at 6.0; 16 CTAs / 8.0: 48.8; one block per SM / 1.0: 48.1 -- profiles/r02_halo_attribution.md;
// never more than would leave them without a tile); the ghost-heavy tiles (>= 64 ghost rows) are dealt
// out first, each charged 1 + ghost rows / 128 tile times (256 ghost rows ~ two tiles, measured with
// scripts/probe_fused2.py); then the other tiles go, IN INDEX ORDER, to the CTA with the least load so
// far -- round-robin again, except that a CTA holding a heavy tile or a push block sits out as many
// rounds as it was charged.  Tiny matrices, where the charges would leave a CTA without any tile
// (every CTA of the launch must own one), and lpt = false fall back to tile t on CTA t % grid.
static void halo_tile_schedule(int ntiles, const int *gcount, int grid, int npush, double push_charge, bool lpt, int *cta_of)
{
  if (lpt && grid > 0) {
    using Load = std::pair<double, int>;
    std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
    const double charge = std::min(push_charge, std::max(0.0, (double)(ntiles / grid) - 1.0));
    for (int b = 0; b < grid; ++b) heap.push({b < npush ? charge : 0.0, b});
    std::vector<int> count((size_t)grid, 0);
    auto deal = [&](int t) {
      Load top = heap.top();
      heap.pop();
      cta_of[t] = top.second;
      count[top.second]++;
      heap.push({top.first + 1.0 + gcount[t] / 128.0, top.second});
    };
    for (int t = 0; t < ntiles; ++t) if (gcount[t] >= 64) deal(t);
    for (int t = 0; t < ntiles; ++t) if (gcount[t] < 64) deal(t);
    bool starved = false;
    for (int b = 0; b < grid && b < ntiles; ++b) starved = starved || count[b] == 0;
    if (!starved) return;
  }
  for (int t = 0; t < ntiles; ++t) cta_of[t] = grid > 0 ? t % grid : 0;
}
// host only: the schedule for the CPU tests
extern "C" int b200_mpiaij_tile_schedule(int32_t ntiles, const int32_t *ghost_rows_per_tile, int32_t grid, int32_t npush,
                                         double push_charge, int32_t *cta_of)
{
  if (ntiles < 0 || grid < 1 || (ntiles && (!ghost_rows_per_tile || !cta_of))) return set_error(B200_ERR_ARG, "b200_mpiaij_tile_schedule: bad argument");
  halo_tile_schedule(ntiles, ghost_rows_per_tile, grid, npush, push_charge, true, cta_of);
  return B200_OK;
}

// Elements per push block of the fused launch.  Measured on 8 B200s (profiles/r02_halo_attribution.md):
// with one push block on every SM the stores to peer memory and the system-scope fences behind them
// cost the launch 7 us of its 48; concentrated on a few CTAs -- which the tile schedule then gives
// fewer tiles -- the other ~700 CTAs never touch NVLink.  B200_MPIAIJ_PUSH_CTAS (default 32) blocks,
// at least one per destination.
static int fused_push_chunk(b200_mpiaij_s *M, int grid)
{
  long long total = 0;
  for (auto &s : M->sends) total += s.count;
  const int want = std::max(1, std::min(env_int("B200_MPIAIJ_PUSH_CTAS", 32), grid / 4));
  const int room = std::max(1, want - (int)M->sends.size());
  return (int)std::max<long long>(256, ((total + room - 1) / room + 255) / 256 * 256);
}
static int fused_push_blocks(b200_mpiaij_s *M, int grid)
{
  const int chunk = fused_push_chunk(M, grid);
  int n = 0;
  for (auto &s : M->sends) n += (s.count + chunk - 1) / chunk;
  return n;
}

extern "C" int b200_mpiaij_upload(b200_mpiaij_t M)
{
  NvtxRange nvtx_("b200_mpiaij_upload");
  if (!M) return set_error(B200_ERR_ARG, "null handle");
  if (M->uploaded) return B200_OK;
  B200_TRY(ensure_device());
  B200_TRY(b200_csr_create(&M->A, M->nloc, M->nloc, M->Ai.data(), M->Aj.data(), M->Aa.data()));
  B200_TRY(b200_csr_create(&M->B, M->nloc, (int32_t)M->garray.size(), M->Bi.data(), M->Bj.data(), M->Ba.data()));
  B200_TRY(up(&M->d_cpi, M->cpi));
  B200_TRY(up(&M->d_ridx, M->ridx));
  B200_TRY(up(&M->d_bj, M->Bj));
  B200_TRY(up(&M->d_ba, M->Ba));
  B200_TRY(up(&M->d_srcs, M->srcs));
  M->window_bytes = WINDOW_HDR_BYTES + (size_t)2 * M->ngpad * sizeof(double) + 256;
  B200_CUDA_TRY(cudaMalloc((void **)&M->d_window, M->window_bytes));
  B200_CUDA_TRY(cudaMemset(M->d_window, 0, M->window_bytes));
  B200_CUDA_TRY(cudaStreamCreateWithFlags(&M->side, cudaStreamNonBlocking));
  B200_CUDA_TRY(cudaEventCreateWithFlags(&M->ev_fork, cudaEventDisableTiming));
  B200_CUDA_TRY(cudaEventCreateWithFlags(&M->ev_join, cudaEventDisableTiming));
  M->timeout_ns = (unsigned long long)env_int("B200_MPIAIJ_TIMEOUT_MS", 10000) * 1000000ull;
  // fused launch: group B's compressed rows by the stream CTA that owns the tile of the row
  // (tile t belongs to CTA t % grid), so that each CTA finishes exactly the rows it wrote
  {
    int4 *d_tiles = nullptr;
    int   ntiles = 0, grid = 0, threads = 0;
    B200_TRY(stream_plan_tiles(M->A, &d_tiles, &ntiles, &grid, &threads));
    if (ntiles && grid > 0 && env_int("B200_MPIAIJ_FUSED", 1)) {
      std::vector<int4> tiles((size_t)ntiles);
      B200_CUDA_TRY(cudaMemcpy(tiles.data(), d_tiles, sizeof(int4) * (size_t)ntiles, cudaMemcpyDeviceToHost));
      // ghost rows per tile
      std::vector<int> tile_of(M->ridx.size()), gcount((size_t)ntiles, 0);
      {
        int t = 0;
        for (size_t c = 0; c < M->ridx.size(); ++c) {
          while (t < ntiles && M->ridx[c] >= tiles[t].y) ++t;
          tile_of[c] = t;
          gcount[t]++;
        }
      }
      // tile -> CTA (halo_tile_schedule), then the CTA-major copy of the tile table
      std::vector<int> cta_of((size_t)ntiles);
      const bool lpt = env_int("B200_MPIAIJ_SCHED", 1) != 0;
      halo_tile_schedule(ntiles, gcount.data(), grid, lpt ? std::min(grid, fused_push_blocks(M, grid)) : 0,
                         env_int("B200_MPIAIJ_PUSH_CHARGE_TENTHS", 60) / 10.0, lpt, cta_of.data());
      std::vector<int>  first((size_t)grid + 1, 0);
      std::vector<int4> sched_tiles((size_t)ntiles);
      for (int t = 0; t < ntiles; ++t) first[cta_of[t] + 1]++;
      for (int b = 0; b < grid; ++b) first[b + 1] += first[b];
      {
        std::vector<int> next(first.begin(), first.end() - 1);
        for (int t = 0; t < ntiles; ++t) sched_tiles[next[cta_of[t]]++] = tiles[t];   // ascending inside a CTA
      }
      std::vector<int> owner(M->ridx.size()), ptr((size_t)grid + 1, 0), rows(M->ridx.size());
      for (size_t c = 0; c < M->ridx.size(); ++c) {
        owner[c] = cta_of[tile_of[c]];
        ptr[owner[c] + 1]++;
      }
      for (int b = 0; b < grid; ++b) ptr[b + 1] += ptr[b];
      std::vector<int> next(ptr.begin(), ptr.end() - 1);
      for (size_t c = 0; c < M->ridx.size(); ++c) rows[next[owner[c]]++] = (int)c;
      B200_TRY(up(&M->d_cta_ptr, ptr));
      B200_TRY(up(&M->d_cta_rows, rows));
      // the fused kernel always reads the CTA-major table (with B200_MPIAIJ_SCHED=0 it holds the
      // round-robin assignment)
      B200_TRY(up(&M->d_sched_tiles, sched_tiles));
      B200_TRY(up(&M->d_sched_first, first));
      M->fused_grid = grid;
      M->fused_ok = true;
    }
  }
  B200_CUDA_TRY(cudaDeviceSynchronize());
  M->uploaded = true;
  return B200_OK;
}

extern "C" int b200_mpiaij_get_blocks(b200_mpiaij_t M, b200_csr_t *A, b200_csr_t *B)
{
  if (!M || !M->uploaded) return set_error(B200_ERR_STATE, "b200_mpiaij_upload first");
  if (A) *A = M->A;
  if (B) *B = M->B;
  return B200_OK;
}
extern "C" int b200_mpiaij_window_ipc_handle(b200_mpiaij_t M, void *handle64)
{
  if (!M || !M->uploaded || !handle64) return set_error(B200_ERR_STATE, "b200_mpiaij_upload first");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  cudaIpcMemHandle_t h;
  B200_CUDA_TRY(cudaIpcGetMemHandle(&h, M->d_window));
  memcpy(handle64, &h, 64);
  return B200_OK;
}
extern "C" int b200_mpiaij_window_ptr(b200_mpiaij_t M, void **d)
{
  if (!M || !M->uploaded || !d) return set_error(B200_ERR_STATE, "b200_mpiaij_upload first");
  *d = M->d_window;
  return B200_OK;
}
extern "C" int b200_mpiaij_set_peer_window(b200_mpiaij_t M, int32_t peer, void *d_window)
{
  if (!M) return set_error(B200_ERR_ARG, "null handle");
  for (auto &s : M->sends) if (s.peer == peer) { s.peer_window = d_window; s.opened = false; }
  M->push_ready = false;
  return B200_OK;
}
extern "C" int b200_mpiaij_open_peer_window(b200_mpiaij_t M, int32_t peer, const void *handle64)
{
  if (!M || !handle64) return set_error(B200_ERR_ARG, "null argument");
  for (auto &s : M->sends)
    if (s.peer == peer) {
      if (s.peer_window) continue;  // already mapped
      // an IPC handle may be opened once per process: reuse a mapping made for the all-reduce
      if (!M->all_windows.empty() && M->all_windows[peer]) { s.peer_window = M->all_windows[peer]; s.opened = false; continue; }
      cudaIpcMemHandle_t h;
      memcpy(&h, handle64, 64);
      void *p = nullptr;
      B200_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      s.peer_window = p;
      s.opened = true;
    }
  M->push_ready = false;
  return B200_OK;
}

// The two lvec buffers alternate with the MatMult number and nothing tells a sender that a buffer has
// been read: what keeps a sender from running two MatMults ahead of a receiver is that it waits for
// that receiver's flag itself, every MatMult.  That holds exactly when every rank this one sends to
// is also one it receives from (always true for a structurally symmetric matrix such as the
// reference's).  Other patterns are refused here; the caller-owned transport (b200_mpiaij_pack +
// b200_mpiaij_mult_add_ghost, e.g. over NCCL) has no such restriction.
extern "C" int b200_mpiaij_pattern_symmetric(b200_mpiaij_t M, int32_t *first_unmatched_peer)
{
  if (!M) return set_error(B200_ERR_ARG, "null handle");
  if (first_unmatched_peer) *first_unmatched_peer = -1;
  for (auto &s : M->sends)
    if (!std::binary_search(M->srcs.begin(), M->srcs.end(), s.peer)) {
      if (first_unmatched_peer) *first_unmatched_peer = s.peer;
      return set_error(B200_ERR_STATE,
                       "rank %d sends ghost values to rank %d but receives none from it: the push exchange needs a "
                       "symmetric communication pattern (use b200_mpiaij_pack / b200_mpiaij_mult_add_ghost with your own transport)",
                       M->rank, s.peer);
    }
  return B200_OK;
}

// freeze the send side: flat index array, CTA table, per-peer destinations
static int prepare_push(b200_mpiaij_s *M)
{
  if (M->push_ready) return B200_OK;
  B200_TRY(b200_mpiaij_pattern_symmetric(M, nullptr));
  cudaFree(M->d_send_idx); cudaFree(M->d_blocks); cudaFree(M->d_peers); cudaFree(M->d_done);
  M->d_send_idx = nullptr; M->d_blocks = nullptr; M->d_peers = nullptr; M->d_done = nullptr;
  std::vector<int32_t>   flat;
  std::vector<PushBlock> blocks;
  std::vector<PushPeer>  peers;
  M->send_start.clear();
  // fused launch: a few dozen push blocks on the first CTAs (fused_push_chunk); stand-alone push
  // kernel: PUSH_CHUNK elements per CTA
  const int chunk = M->fused_ok ? fused_push_chunk(M, M->fused_grid) : PUSH_CHUNK;
  int slot = 0;
  for (auto &s : M->sends) {
    if (!s.peer_window) return set_error(B200_ERR_STATE, "peer %d needs data but its window is not mapped", s.peer);
    unsigned char *w = (unsigned char *)s.peer_window;
    PushPeer p;
    // dst is biased by -start so that the kernel can index with the flat element number
    const int32_t start = (int32_t)flat.size();
    p.dst[0] = (double *)(w + WINDOW_HDR_BYTES) + s.peer_offset - start;
    p.dst[1] = (double *)(w + WINDOW_HDR_BYTES) + s.peer_ngpad + s.peer_offset - start;
    p.flag   = (unsigned long long *)w + M->rank;
    p.nblocks = (s.count + chunk - 1) / chunk;
    p.pad = 0;
    for (int b = 0; b < p.nblocks; ++b)
      blocks.push_back(PushBlock{slot, start + b * chunk, std::min(chunk, s.count - b * chunk), 0});
    M->send_start.push_back(start);
    flat.insert(flat.end(), s.idx.begin(), s.idx.end());
    peers.push_back(p);
    ++slot;
  }
  M->npush_blocks = (int)blocks.size();
  B200_TRY(up(&M->d_send_idx, flat));
  B200_TRY(up(&M->d_blocks, blocks));
  B200_TRY(up(&M->d_peers, peers));
  B200_CUDA_TRY(cudaMalloc((void **)&M->d_done, sizeof(unsigned) * std::max<size_t>(peers.size(), 1)));
  B200_CUDA_TRY(cudaMemset(M->d_done, 0, sizeof(unsigned) * std::max<size_t>(peers.size(), 1)));
  M->push_ready = true;
  return B200_OK;
}

extern "C" int b200_mpiaij_mult_begin(b200_mpiaij_t M, const double *d_x, void *stream)
{
  if (!M || !M->uploaded) return set_error(B200_ERR_STATE, "b200_mpiaij_upload first");
  B200_TRY(prepare_push(M));
  M->seq += 1;
  if (M->npush_blocks)
    B200_LAUNCH(k_halo_push, M->npush_blocks, 256, 0, (cudaStream_t)stream, M->d_blocks, M->d_peers,
                M->d_send_idx, d_x, M->d_done, M->seq);
  return B200_OK;
}

extern "C" int b200_mpiaij_mult_local(b200_mpiaij_t M, const double *d_x, double *d_y, int mode, void *stream)
{
  if (!M || !M->uploaded) return set_error(B200_ERR_STATE, "b200_mpiaij_upload first");
  return b200_spmv(M->A, d_x, d_y, mode, stream);
}

static int offdiag(b200_mpiaij_s *M, const double *lvec, double *d_y, int mode, bool wait, cudaStream_t st)
{
  const int nrows = (int)M->ridx.size();
  if (!nrows) return B200_OK;
  const unsigned long long *flags = (const unsigned long long *)M->d_window;
  unsigned long long       *err   = (unsigned long long *)M->d_window + ERR_WORD;
  const int grid = (nrows + 127) / 128, nsrc = (int)M->srcs.size();
#define OFFDIAG(MODE_, WAIT_)                                                                   \
  B200_LAUNCH((k_offdiag<MODE_, WAIT_>), grid, 128, 0, st, nrows, M->d_cpi, M->d_ridx, M->d_bj,  \
              M->d_ba, lvec, d_y, flags, M->d_srcs, nsrc, M->seq, err, M->timeout_ns)
  if (mode == B200_MODE_EXACT) { if (wait) OFFDIAG(B200_MODE_EXACT, true); else OFFDIAG(B200_MODE_EXACT, false); }
  else { if (wait) OFFDIAG(B200_MODE_EXACT_FMA, true); else OFFDIAG(B200_MODE_EXACT_FMA, false); }
#undef OFFDIAG
  return B200_OK;
}

extern "C" int b200_mpiaij_mult_end(b200_mpiaij_t M, double *d_y, int mode, void *stream)
{
  if (!M || !M->uploaded) return set_error(B200_ERR_STATE, "b200_mpiaij_upload first");
  const double *lvec = (const double *)(M->d_window + WINDOW_HDR_BYTES) + (size_t)(M->seq & 1) * M->ngpad;
  return offdiag(M, lvec, d_y, mode, true, (cudaStream_t)stream);
}

static HaloArgs fused_args(b200_mpiaij_s *M, bool with_push)
{
  HaloArgs h{};
  h.blocks = M->d_blocks; h.peers = M->d_peers; h.send_idx = M->d_send_idx; h.done = M->d_done;
  h.npush = with_push ? M->npush_blocks : 0;
  h.sched_tiles = M->d_sched_tiles; h.sched_first = M->d_sched_first;
  h.cta_ptr = M->d_cta_ptr; h.cta_rows = M->d_cta_rows; h.cpi = M->d_cpi; h.ridx = M->d_ridx; h.bj = M->d_bj; h.ba = M->d_ba;
  h.lvec = (const double *)(M->d_window + WINDOW_HDR_BYTES) + (size_t)(M->seq & 1) * M->ngpad;
  h.flags = (const unsigned long long *)M->d_window; h.srcs = M->d_srcs; h.nsrc = (int)M->srcs.size();
  h.seq = M->seq; h.err = (unsigned long long *)M->d_window + ERR_WORD; h.timeout_ns = M->timeout_ns;
  return h;
}

// A x + B lvec for the MatMult whose push was issued by b200_mpiaij_mult_begin: one fused launch
// when the diagonal block runs the stream kernel, otherwise the two kernels.
extern "C" int b200_mpiaij_mult_finish(b200_mpiaij_t M, const double *d_x, double *d_y, int mode, void *stream)
{
  if (!M || !M->uploaded) return set_error(B200_ERR_STATE, "b200_mpiaij_upload first");
  if (M->fused_ok) return launch_stream_halo(M->A, d_x, d_y, mode, fused_args(M, false), (cudaStream_t)stream);
  B200_TRY(b200_mpiaij_mult_local(M, d_x, d_y, mode, stream));
  return b200_mpiaij_mult_end(M, d_y, mode, stream);
}

static int mpiaij_mult(b200_mpiaij_t M, const double *d_x, double *d_y, int mode, void *stream, const DotArgs *dot)
{
  NvtxRange nvtx_("b200_mpiaij_mult");
  if (!M || !M->uploaded) return set_error(B200_ERR_STATE, "b200_mpiaij_upload first");
  cudaStream_t st = (cudaStream_t)stream;
  B200_TRY(prepare_push(M));
  if (M->fused_ok && M->npush_blocks <= M->fused_grid) {
    // one launch: push prologue + A x + B lvec (k_stream<..., HALO>), optionally (x, y) too
    M->seq += 1;
    return launch_stream_halo(M->A, d_x, d_y, mode, fused_args(M, true), st, dot);
  }
  if (M->npush_blocks) {
    // fork: the push runs beside A x; join so that the caller may overwrite x afterwards
    B200_CUDA_TRY(cudaEventRecord(M->ev_fork, st));
    B200_CUDA_TRY(cudaStreamWaitEvent(M->side, M->ev_fork, 0));
    B200_TRY(b200_mpiaij_mult_begin(M, d_x, M->side));
    B200_CUDA_TRY(cudaEventRecord(M->ev_join, M->side));
  } else {
    M->seq += 1;
  }
  B200_TRY(b200_mpiaij_mult_local(M, d_x, d_y, mode, st));
  B200_TRY(b200_mpiaij_mult_end(M, d_y, mode, st));
  if (M->npush_blocks) B200_CUDA_TRY(cudaStreamWaitEvent(st, M->ev_join, 0));
  if (dot) {
    // the two-kernel path reduces after the fact
    B200_TRY(b200_vec_dot(d_x, d_y, M->nloc, dot->out, stream));
  }
  return B200_OK;
}

extern "C" int b200_mpiaij_mult(b200_mpiaij_t M, const double *d_x, double *d_y, int mode, void *stream)
{
  return mpiaij_mult(M, d_x, d_y, mode, stream, nullptr);
}

// y rows that touch a ghost, packed: the tail of the pipelined host-vector MatMult
__global__ void k_gather_rows(int n, const int *__restrict__ ridx, const double *__restrict__ y, double *__restrict__ out)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = y[ridx[t]];
}

// Row-blocked pipeline of the host-vector MatMult_MPIAIJ (the reference's step 4,
// src/openacc-step4/MatMult_SeqAIJ.patch:51-72, on the row-partitioned matrix): x goes up in blocks;
// block b's A x starts as soon as the last block it reads has landed and its y rows go down while
// later blocks are still on their way up (PCIe is full duplex); the halo push follows the last
// upload; the ghost rows (~2 % of the rows, scattered over every block) are finished last, packed,
// brought down in one small copy and patched into h_y by the host.  Same arithmetic, same order.
static int mult_host_pipelined(b200_mpiaij_s *M, const double *h_x, double *h_y, int mode, int nblk)
{
  if (!M->hs_up) {
    B200_CUDA_TRY(cudaStreamCreateWithFlags(&M->hs_up, cudaStreamNonBlocking));
    B200_CUDA_TRY(cudaStreamCreateWithFlags(&M->hs_dn, cudaStreamNonBlocking));
    B200_CUDA_TRY(cudaStreamCreateWithFlags(&M->hs_push, cudaStreamNonBlocking));
    B200_CUDA_TRY(cudaEventCreateWithFlags(&M->ev_push, cudaEventDisableTiming));
    const size_t ng = std::max<size_t>(M->ridx.size(), 1);
    B200_CUDA_TRY(cudaMalloc((void **)&M->d_gbuf, ng * sizeof(double)));
    B200_CUDA_TRY(cudaHostAlloc((void **)&M->h_gbuf, ng * sizeof(double), cudaHostAllocDefault));
  }
  while ((int)M->evx.size() < nblk) {
    cudaEvent_t a, b;
    B200_CUDA_TRY(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    B200_CUDA_TRY(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    M->evx.push_back(a); M->evk.push_back(b);
  }
  cudaStream_t sk = M->hstream;
  int r0, r1, need;
  for (int b = 0; b < nblk; ++b) {
    host_block_info(M->A, b, &r0, &r1, &need);
    B200_CUDA_TRY(cudaMemcpyAsync(M->d_hx + r0, h_x + r0, (size_t)(r1 - r0) * sizeof(double), cudaMemcpyHostToDevice, M->hs_up));
    B200_CUDA_TRY(cudaEventRecord(M->evx[b], M->hs_up));
  }
  // VecScatterBegin once every row of x is on the device (the boundary values lie all over it)
  B200_CUDA_TRY(cudaStreamWaitEvent(M->hs_push, M->evx[nblk - 1], 0));
  B200_TRY(b200_mpiaij_mult_begin(M, M->d_hx, M->hs_push));
  B200_CUDA_TRY(cudaEventRecord(M->ev_push, M->hs_push));
  for (int b = 0; b < nblk; ++b) {
    host_block_info(M->A, b, &r0, &r1, &need);
    B200_CUDA_TRY(cudaStreamWaitEvent(sk, M->evx[need], 0));
    B200_TRY(launch_host_block(M->A, b, M->d_hx, M->d_hy, mode, sk));
    B200_CUDA_TRY(cudaEventRecord(M->evk[b], sk));
    B200_CUDA_TRY(cudaStreamWaitEvent(M->hs_dn, M->evk[b], 0));
    B200_CUDA_TRY(cudaMemcpyAsync(h_y + r0, M->d_hy + r0, (size_t)(r1 - r0) * sizeof(double), cudaMemcpyDeviceToHost, M->hs_dn));
  }
  // VecScatterEnd + MatMultAdd on the ghost rows, then their final values in one packed copy
  B200_CUDA_TRY(cudaStreamWaitEvent(sk, M->ev_push, 0));
  B200_TRY(b200_mpiaij_mult_end(M, M->d_hy, mode, sk));
  const int ng = (int)M->ridx.size();
  if (ng) {
    B200_LAUNCH(k_gather_rows, (ng + 255) / 256, 256, 0, sk, ng, (const int *)M->d_ridx, (const double *)M->d_hy, M->d_gbuf);
    B200_CUDA_TRY(cudaMemcpyAsync(M->h_gbuf, M->d_gbuf, (size_t)ng * sizeof(double), cudaMemcpyDeviceToHost, sk));
  }
  B200_CUDA_TRY(cudaMemcpyAsync(M->h_err, (unsigned long long *)M->d_window + ERR_WORD, sizeof(unsigned long long), cudaMemcpyDeviceToHost, sk));
  B200_CUDA_TRY(cudaStreamSynchronize(M->hs_dn));
  B200_CUDA_TRY(cudaStreamSynchronize(sk));
  for (int c = 0; c < ng; ++c) h_y[M->ridx[c]] = M->h_gbuf[c];
  if (*M->h_err) return b200_mpiaij_check(M);
  return B200_OK;
}

// MatMult_MPIAIJ(Mat,Vec,Vec) with host Vecs: upload this rank's x rows, multiply, download y.
extern "C" int b200_mpiaij_mult_host(b200_mpiaij_t M, const double *h_x, double *h_y, int mode)
{
  if (!M || !M->uploaded) return set_error(B200_ERR_STATE, "b200_mpiaij_upload first");
  if (!M->d_hx) {
    B200_CUDA_TRY(cudaMalloc((void **)&M->d_hx, std::max<size_t>(M->nloc, 1) * sizeof(double)));
    B200_CUDA_TRY(cudaMalloc((void **)&M->d_hy, std::max<size_t>(M->nloc, 1) * sizeof(double)));
    B200_CUDA_TRY(cudaStreamCreateWithFlags(&M->hstream, cudaStreamNonBlocking));
    B200_CUDA_TRY(cudaHostAlloc((void **)&M->h_err, sizeof(unsigned long long), cudaHostAllocDefault));
    *M->h_err = 0;
  }
  B200_TRY(prepare_push(M));
  const int nblk = env_int("B200_MPIAIJ_HOST_PIPELINE", 1) ? host_block_count(M->A, mode) : 0;
  if (nblk >= 2) return mult_host_pipelined(M, h_x, h_y, mode, nblk);
  B200_CUDA_TRY(cudaMemcpyAsync(M->d_hx, h_x, (size_t)M->nloc * sizeof(double), cudaMemcpyHostToDevice, M->hstream));
  B200_TRY(b200_mpiaij_mult(M, M->d_hx, M->d_hy, mode, M->hstream));
  B200_CUDA_TRY(cudaMemcpyAsync(h_y, M->d_hy, (size_t)M->nloc * sizeof(double), cudaMemcpyDeviceToHost, M->hstream));
  // the time-out word rides down behind y: this entry synchronises anyway, so it can report it
  B200_CUDA_TRY(cudaMemcpyAsync(M->h_err, (unsigned long long *)M->d_window + ERR_WORD, sizeof(unsigned long long), cudaMemcpyDeviceToHost, M->hstream));
  B200_CUDA_TRY(cudaStreamSynchronize(M->hstream));
  if (*M->h_err) return b200_mpiaij_check(M);
  return B200_OK;
}

extern "C" int b200_mpiaij_pack(b200_mpiaij_t M, int32_t peer, const double *d_x, double *d_buf, void *stream)
{
  if (!M || !M->uploaded) return set_error(B200_ERR_STATE, "b200_mpiaij_upload first");
  // a flat index array is needed; build it without requiring peer windows
  if (!M->d_send_idx) {
    std::vector<int32_t> flat;
    M->send_start.clear();
    for (auto &s : M->sends) { M->send_start.push_back((int32_t)flat.size()); flat.insert(flat.end(), s.idx.begin(), s.idx.end()); }
    B200_TRY(up(&M->d_send_idx, flat));
  }
  for (size_t k = 0; k < M->sends.size(); ++k)
    if (M->sends[k].peer == peer) {
      const int cnt = M->sends[k].count;
      B200_LAUNCH(k_halo_pack, (cnt + 255) / 256, 256, 0, (cudaStream_t)stream, cnt, M->d_send_idx + M->send_start[k], d_x, d_buf);
      return B200_OK;
    }
  return B200_OK;
}

extern "C" int b200_mpiaij_mult_add_ghost(b200_mpiaij_t M, const double *d_lvec, double *d_y, int mode, void *stream)
{
  if (!M || !M->uploaded) return set_error(B200_ERR_STATE, "b200_mpiaij_upload first");
  return offdiag(M, d_lvec, d_y, mode, false, (cudaStream_t)stream);
}

extern "C" int b200_mpiaij_check(b200_mpiaij_t M)
{
  if (!M || !M->uploaded) return B200_OK;
  unsigned long long e = 0;
  B200_CUDA_TRY(cudaMemcpy(&e, (unsigned long long *)M->d_window + ERR_WORD, sizeof e, cudaMemcpyDeviceToHost));
  if (e) {
    // report once: the word is cleared so that a later, healthy MatMult is not blamed for this one
    B200_CUDA_TRY(cudaMemset((unsigned long long *)M->d_window + ERR_WORD, 0, sizeof e));
    return set_error(B200_ERR_TIMEOUT, "rank %d: a halo flag was not seen within %llu ms (B200_MPIAIJ_TIMEOUT_MS); the result of that MatMult is not valid",
                     M->rank, M->timeout_ns / 1000000ull);
  }
  return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// all-reduce and the distributed CG
// ---------------------------------------------------------------------------------------------
// Windows of ALL ranks are needed for the all-reduce (the halo only maps the peers it sends to).
extern "C" int b200_mpiaij_set_rank_window(b200_mpiaij_t M, int32_t q, const void *handle64_or_null, void *d_window_or_null)
{
  if (!M || !M->uploaded || q < 0 || q >= M->size) return set_error(B200_ERR_ARG, "b200_mpiaij_set_rank_window: bad argument");
  if (M->all_windows.empty()) { M->all_windows.assign(M->size, nullptr); M->all_windows_opened.assign(M->size, 0); }
  if (q == M->rank) { M->all_windows[q] = M->d_window; return B200_OK; }
  // reuse a mapping the halo already opened
  for (auto &s : M->sends)
    if (s.peer == q && s.peer_window) { M->all_windows[q] = (unsigned char *)s.peer_window; return B200_OK; }
  if (d_window_or_null) { M->all_windows[q] = (unsigned char *)d_window_or_null; return B200_OK; }
  if (!handle64_or_null) return set_error(B200_ERR_ARG, "no window for rank %d", q);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64_or_null, 64);
  void *p = nullptr;
  B200_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  M->all_windows[q] = (unsigned char *)p;
  M->all_windows_opened[q] = 1;
  return B200_OK;
}

static int allreduce_sum(b200_mpiaij_t M, double *d_vals, int32_t nvals, const int *skip, int post, double *sc, int *sti, void *stream)
{
  NvtxRange nvtx_("b200_mpiaij_allreduce_sum");
  if (!M || !M->uploaded || !d_vals || nvals < 1 || nvals > RED_MAX_VALS) return set_error(B200_ERR_ARG, "b200_mpiaij_allreduce_sum: bad argument");
  if (M->size == 1) return B200_OK;
  if (M->all_windows.empty()) return set_error(B200_ERR_STATE, "b200_mpiaij_set_rank_window for every rank first");
  if (!M->d_all_windows) {
    M->all_windows[M->rank] = M->d_window;
    for (int q = 0; q < M->size; ++q)
      if (!M->all_windows[q]) return set_error(B200_ERR_STATE, "window of rank %d is not mapped", q);
    B200_TRY(up(&M->d_all_windows, M->all_windows));
  }
  M->rseq += 1;
  B200_LAUNCH_PDL(k_allreduce, 1, 128, 0, (cudaStream_t)stream, d_vals, (int)nvals, (int)M->size, (int)M->rank,
                  (unsigned char *const *)M->d_all_windows, M->rseq, (unsigned long long *)M->d_window + ERR_WORD, M->timeout_ns,
                  skip, post, sc, sti);
  return B200_OK;
}

extern "C" int b200_mpiaij_allreduce_sum(b200_mpiaij_t M, double *d_vals, int32_t nvals, void *stream)
{
  return allreduce_sum(M, d_vals, nvals, nullptr, CG_POST_NONE, nullptr, nullptr, stream);
}

// KSPCG + PCJACOBI on the row-partitioned matrix: MatMult_MPIAIJ (one fused launch) + the vector
// kernels of b200_vec.cu on the local rows + all-reduced dot products.
extern "C" int b200_mpiaij_cg_jacobi(b200_mpiaij_t M, const double *d_b, double *d_x, double rtol, double atol,
                                     int32_t max_it, int mode, b200_cg_result_t *res, void *stream)
{
  NvtxRange nvtx_("b200_mpiaij_cg_jacobi");
  if (!M || !M->uploaded || !d_b || !d_x || !res) return set_error(B200_ERR_ARG, "b200_mpiaij_cg_jacobi: bad argument");
  CgOps ops;
  ops.m = M->nloc;
  B200_TRY(b200_csr_device_arrays(M->A, &ops.ai, &ops.aj, &ops.aa));
  if (mode == B200_MODE_FAST) mode = B200_MODE_EXACT_FMA;
  B200_TRY(prepare_push(M));
  const bool fused = M->fused_ok && M->npush_blocks <= M->fused_grid;
  ops.dot_partials = fused ? M->fused_grid : 0;
  ops.mult_dot = [M, mode](const double *p, double *w, const DotArgs &dot, cudaStream_t s) { return mpiaij_mult(M, p, w, mode, s, &dot); };
  if (M->size > 1)
    ops.allreduce = [M](double *v, int n, int post, double *sc, int *sti, cudaStream_t s) {
      return allreduce_sum(M, v, n, sti + CGI_DONE, post, sc, sti, s);
    };
  B200_TRY(cg_jacobi_run(ops, d_b, d_x, rtol, atol, max_it, res, (cudaStream_t)stream));
  return b200_mpiaij_check(M);
}
