/*
 * b200_mpiaij.h -- C ABI of the row-partitioned mat-vec (MatMult_MPIAIJ) for one box of B200s.
 *
 * What it replaces (PETSc 3.7.6, un-vendored, reached from the reference through
 * DMSetMatType(da, MATAIJ) on PETSC_COMM_WORLD, src/helper.cpp:31,39, with aprun -n > 1,
 * runs/single-node-scaling.pbs:60):
 *     Mat_MPIAIJ {A, B, garray, lvec, Mvctx}          mpiaij.h
 *     MatSetValues_MPIAIJ's diagonal / off-diagonal split
 *     MatSetUpMultiply_MPIAIJ                          mmaij.c   (garray, B column compaction)
 *     MatMult_MPIAIJ                                   mpiaij.c  (scatter begin, A x, scatter end,
 *                                                                  y += B lvec)
 *     VecScatter (MPI persistent sends/recvs on host)  vscat.c
 * One process per GPU.  The halo exchange is push based: the sending rank's pack kernel stores its
 * boundary x values straight into the receiving rank's lvec over NVLink (CUDA IPC mapping) and
 * then releases a per-source flag; the receiving rank's off-diagonal kernel acquires the flags.
 * No host round trip, no receive-side copy, and the transfer overlaps A x by construction.
 *
 * Process plumbing (who is rank r, exchanging the 64-byte IPC handles and the garray lists) is the
 * caller's: bench.py / tests use torch.distributed for it.  Everything returns 0 or a b200 error.
 */
#ifndef B200_MPIAIJ_H
#define B200_MPIAIJ_H

#include "b200_seqaij.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_mpiaij_s *b200_mpiaij_t;

/* Host-side construction (no GPU needed): rows [base[rank], base[rank+1]) of the global matrix
 * with GLOBAL column ids ascending in each row.  Splits into A (columns owned by this rank,
 * local ids) and B (the rest), builds garray = sorted unique ghost ids, rewrites B's columns to
 * positions in garray, and the receive offsets per owner rank. */
int b200_mpiaij_create(b200_mpiaij_t *out, int32_t size, int32_t rank, const int32_t *base,
                       const int32_t *h_ai, const int32_t *h_aj_global, const double *h_aa);
int b200_mpiaij_destroy(b200_mpiaij_t M);

/* sizes[6] = nloc, nnz(A), nnz(B), nghost, non-empty rows of B, number of source ranks */
int b200_mpiaij_get_sizes(b200_mpiaij_t M, int32_t *sizes);
int b200_mpiaij_get_garray(b200_mpiaij_t M, int32_t *garray);
int b200_mpiaij_get_recv_offsets(b200_mpiaij_t M, int32_t *off /* size+1 */);
/* host copies of a block: which = 0 (A) or 1 (B, compact columns); any pointer may be NULL */
int b200_mpiaij_copy_block(b200_mpiaij_t M, int which, int32_t *ai, int32_t *aj, double *aa);

/* Tell this rank what `peer` needs: peer's garray (ascending global ids).  The send list is the
 * run of ids inside this rank's ownership range, in that order (VecScatter's convention), landing
 * at lvec offset = position of the run in peer's garray. */
int b200_mpiaij_set_peer_garray(b200_mpiaij_t M, int32_t peer, const int32_t *peer_garray,
                                int32_t peer_nghost);
/* the resulting send list (local x indices) for tests; returns count through *count */
int b200_mpiaij_get_send_list(b200_mpiaij_t M, int32_t peer, int32_t *count, int32_t *local_idx,
                              int32_t *peer_offset);

/* ---- device side ------------------------------------------------------------------------- */
/* Upload A and B (kernel plans included) and allocate this rank's window:
 * [lvec buffer 0][lvec buffer 1][flags: one uint64 per source rank][error word].            */
int b200_mpiaij_upload(b200_mpiaij_t M);
int b200_mpiaij_get_blocks(b200_mpiaij_t M, b200_csr_t *A, b200_csr_t *B); /* borrowed */
int b200_mpiaij_window_ipc_handle(b200_mpiaij_t M, void *handle64);  /* cudaIpcMemHandle_t   */
int b200_mpiaij_window_ptr(b200_mpiaij_t M, void **d_window);
int b200_mpiaij_open_peer_window(b200_mpiaij_t M, int32_t peer, const void *handle64);
int b200_mpiaij_set_peer_window(b200_mpiaij_t M, int32_t peer, void *d_window); /* same process */

/* y = A x_local + B x_ghost.  begin = VecScatterBegin (pack + push + release flags),
 * local = A x, end = VecScatterEnd + MatMultAdd (acquire flags, y += B lvec).
 * b200_mpiaij_mult runs the three in ONE kernel launch when the diagonal block uses the stream
 * kernel (push CTAs first in the grid, B rows folded into the tile loop); otherwise the push goes
 * to an internal side stream and A, B run on `stream`.                                          */
int b200_mpiaij_mult_begin(b200_mpiaij_t M, const double *d_x, void *stream);
int b200_mpiaij_mult_local(b200_mpiaij_t M, const double *d_x, double *d_y, int mode, void *stream);
int b200_mpiaij_mult_end(b200_mpiaij_t M, double *d_y, int mode, void *stream);
int b200_mpiaij_mult(b200_mpiaij_t M, const double *d_x, double *d_y, int mode, void *stream);
/* A x + B lvec after a separate b200_mpiaij_mult_begin (one fused launch when possible).        */
int b200_mpiaij_mult_finish(b200_mpiaij_t M, const double *d_x, double *d_y, int mode, void *stream);
/* The same with HOST vectors (PETSc 3.7.6 Vecs): this rank's x rows up, y rows down, synchronous. */
int b200_mpiaij_mult_host(b200_mpiaij_t M, const double *h_x, double *h_y, int mode);
/* Variant for a transport owned by the caller (NCCL send/recv through torch.distributed):
 * pack boundary values for `peer` into d_buf, and y += B lvec from a caller-filled lvec.       */
int b200_mpiaij_pack(b200_mpiaij_t M, int32_t peer, const double *d_x, double *d_buf, void *stream);
int b200_mpiaij_mult_add_ghost(b200_mpiaij_t M, const double *d_lvec, double *d_y, int mode,
                               void *stream);
/* ---- CG on the partitioned matrix (KSPSolve_CG's VecDot/VecNorm need a global sum) ------------ */
/* The all-reduce needs every rank's window: pass the rank's IPC handle (other process), or its
 * device pointer (same process), or neither when the halo already mapped it.                   */
int b200_mpiaij_set_rank_window(b200_mpiaij_t M, int32_t rank, const void *handle64_or_null,
                                void *d_window_or_null);
/* in-place sum of 1..3 device scalars over all ranks, added in rank order (same bits everywhere) */
int b200_mpiaij_allreduce_sum(b200_mpiaij_t M, double *d_vals, int32_t nvals, void *stream);
/* KSPCG + PCJACOBI; d_b, d_x are this rank's rows; collective over the ranks                   */
int b200_mpiaij_cg_jacobi(b200_mpiaij_t M, const double *d_b, double *d_x, double rtol, double atol,
                          int32_t max_it, int mode, b200_cg_result_t *res, void *stream);
/* B200_ERR_TIMEOUT after a flag wait ran out of its spin budget (B200_MPIAIJ_TIMEOUT_MS, default
 * 10000): the kernel gave up waiting for a peer and the y of that MatMult is not valid.  The
 * asynchronous entries (b200_mpiaij_mult, _mult_finish, _mult_end) cannot return it themselves: call
 * this after synchronising, before trusting y.  b200_mpiaij_mult_host and b200_mpiaij_cg_jacobi,
 * which synchronise anyway, call it for you.  Reporting clears the condition.                     */
int b200_mpiaij_check(b200_mpiaij_t M);
/* The push exchange alternates two receive buffers and relies on every rank this one sends to also
 * being one it receives from (true for structurally symmetric matrices).  B200_OK, or
 * B200_ERR_STATE with the first unmatched peer; b200_mpiaij_mult* refuse such a pattern with the
 * same error.  Host only, valid after every b200_mpiaij_set_peer_garray call.                     */
int b200_mpiaij_pattern_symmetric(b200_mpiaij_t M, int32_t *first_unmatched_peer);

/* The tile -> CTA schedule of the fused launch, host only (for tests): tiles that leave many ghost
 * rows for the closing phase and the first `npush` CTAs (which also carry a push block) are charged
 * extra tile times; everything else is dealt in index order to the least-loaded CTA.  cta_of[ntiles]. */
int b200_mpiaij_tile_schedule(int32_t ntiles, const int32_t *ghost_rows_per_tile, int32_t grid,
                              int32_t npush, double push_charge, int32_t *cta_of);

#ifdef __cplusplus
}
#endif
#endif /* B200_MPIAIJ_H */
