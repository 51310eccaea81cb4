/*
 * tgcn_b200.h — C ABI of libtgcn_b200.so: the sm_100a kernels beneath TextGCN's LightGCN hot path.
 *
 * The reference (sergey-volokhin/TextGCN) has no FFI: its plugin surface is Python method
 * override on BaseModel.  Each entry point below replaces the ATen call chain of
 * one of those methods; the cited file:line is the reference code it stands in for (paths under
 * the reference's TextGCN/ package).  INTEGRATION.md shows the ctypes binding and the subclass a
 * reference maintainer would add.
 *
 * Conventions
 *   - Every pointer named d_* is a DEVICE pointer, h_* a HOST pointer.  No torch types appear here.
 *   - All matrices are row-major fp32; index arrays are int32; every COMPUTE call is asynchronous on
 *     `stream` (a cudaStream_t passed as void*), never synchronises, and allocates nothing:
 *     the caller owns inputs, outputs and workspaces.  Only graph handles own device memory, and only
 *     handle creation (tgcn_graph_create*, tgcn_graph_build_transpose_perm: init-time, like the
 *     reference's dataset construction) synchronises `stream` and walks the row pointer on the host.
 *   - A graph handle serialises its own launches: calls that share a handle must be issued on ONE stream
 *     at a time (the long-row partial sums and arrival counters live in the handle / its workspace);
 *     a launch on another stream while the previous one is still pending is refused with an error.
 *   - Node numbering follows the reference: rows [0, n_users) are users, rows [n_users, n_users +
 *     n_items) are items (dataset.py:130-131).  N = n_users + n_items.
 *   - Return value 0 = ok; otherwise tgcn_last_error() (thread-local) describes the failure.
 *     There is no CPU fallback: without a CUDA device every compute call fails.
 *   - Embedding width d must be a multiple of 4 and rows 16-byte aligned (128-bit loads).
 */
#ifndef TGCN_B200_H
#define TGCN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGCN_ABI_VERSION 3
#define TGCN_MAX_LAYERS 15
#define TGCN_MAX_PEERS 8
#define TGCN_MAX_TOPK 128
#define TGCN_ADV_MAX_CANDIDATES 2048

typedef struct tgcn_graph tgcn_graph_t;
typedef void* tgcn_stream_t; /* cudaStream_t */

int tgcn_abi_version(void);
const char* tgcn_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Graph handle.  Consumes Â exactly as dataset.norm_matrix holds it (dataset.py:122-157): the
 * coalesced COO sorted by (row, col) IS CSR order, so the caller passes rowptr (N+1), col (nnz)
 * and val (nnz) on the device.  The arrays are borrowed and must outlive the handle.  The handle
 * owns: the long-row segment list used for load balancing and (built on first need) the
 * transpose permutation tperm[p] = position of entry (c, r) for entry p = (r, c), which lets the
 * backward pass apply Â_dropᵀ without rebuilding a matrix (base_model.py:77-86).
 * --------------------------------------------------------------------------------------------- */
int tgcn_graph_create(tgcn_graph_t** out, int64_t n_users, int64_t n_items, int64_t nnz,
                      const int32_t* d_rowptr, const int32_t* d_col, const float* d_val,
                      tgcn_stream_t stream);
/* Row-block variant for multi-GPU: the handle covers rows [row_begin, row_begin + n_local_rows) of Â;
 * d_rowptr has n_local_rows + 1 entries starting at 0; d_col holds indices into the gathered table. */
int tgcn_graph_create_block(tgcn_graph_t** out, int64_t n_users, int64_t n_items, int64_t row_begin,
                            int64_t n_local_rows, int64_t nnz_local, const int32_t* d_rowptr,
                            const int32_t* d_col, const float* d_val, tgcn_stream_t stream);
int tgcn_graph_build_transpose_perm(tgcn_graph_t* g, tgcn_stream_t stream);
void tgcn_graph_destroy(tgcn_graph_t* g);
int64_t tgcn_graph_num_segments(const tgcn_graph_t* g);
/* Eval masks read the user rows of a handle: local row = user id - row_begin, an entry equals col_offset + item id.
 * col_offset defaults to n_users (global column numbering); a row block whose columns are item ids sets 0. */
int tgcn_graph_set_mask_col_offset(tgcn_graph_t* g, int64_t col_offset);
/* bytes of caller-provided workspace for propagate_fwd / propagate_bwd */
int64_t tgcn_propagate_workspace_bytes(const tgcn_graph_t* g, int64_t d, int32_t n_layers);

/* a4  BaseModel.layer_aggregation (base_model.py:141-148): Y = Â · X, X and Y are (N, d). */
int tgcn_spmm_fwd(const tgcn_graph_t* g, int64_t d, const float* d_x, float* d_y,
                  void* d_workspace, int64_t workspace_bytes, tgcn_stream_t stream);

/* a4 with a3/a5 fused: one SpMM pass  Y = (add_0 + ... + add_{n_add-1} + Â·X) / divisor  [+ Y if accumulate].
 * X is read through two base pointers (d_x_item NULL = contiguous after the user rows); d_keep/dropout apply the
 * edge-dropout mask (transposed != 0 applies Â_dropᵀ through the transpose permutation).  h_add_user / h_add_item
 * are HOST arrays of n_add device pointers (item part NULL = contiguous).  This is the per-hop building block the
 * multi-GPU host code calls between all-gathers; for a row-block handle X is the gathered table indexed by d_col
 * and Y / addends are the local rows. */
int tgcn_spmm_ex(const tgcn_graph_t* g, int64_t d, const float* d_x_user, const float* d_x_item,
                 const uint8_t* d_keep, float dropout, int32_t transposed, int32_t n_add,
                 const float* const* h_add_user, const float* const* h_add_item, float divisor,
                 int32_t accumulate, float* d_y, void* d_workspace, int64_t workspace_bytes, tgcn_stream_t stream);

/* a5 alone: out = (add_0 + ... + add_{n_add-1}) / divisor over n floats (n % 4 == 0) — layer_combination
 * (base_model.py:150-157) for a table that no local SpMM produces (the all-reduced item table in multi-GPU runs). */
int tgcn_layer_mean(int64_t n, int32_t n_add, const float* const* h_add, float divisor, float* d_out,
                    tgcn_stream_t stream);

/* a2+a3+a4×L+a5  BaseModel.representation (base_model.py:88-106, :150-164).
 * E0 is read through two base pointers (no torch.cat); d_keep is the Bernoulli keep-mask over the
 * nnz of Â drawn by the caller (base_model.py:82) or NULL in eval mode; survivors are scaled by
 * 1/(1-dropout) (:84).  single != 0 returns the last layer (:159-164) instead of the mean (:157).
 * d_out is (N, d): rows [0,n_users) users_emb, the rest items_emb (:106). */
int tgcn_propagate_fwd(const tgcn_graph_t* g, int64_t d, int32_t n_layers, int32_t single,
                       const float* d_user_w, const float* d_item_w, const uint8_t* d_keep, float dropout,
                       float* d_out, void* d_workspace, int64_t workspace_bytes, tgcn_stream_t stream);

/* e (multi-GPU)  representation with the FEATURE dimension sliced across GPUs.  E' = Â·E acts on every column
 * independently, so GPU p can run all L hops on columns [col_off, col_off + d_slice) of the d_full-wide tables with no
 * exchange between hops: d_user_slice (n_users, d_slice) and d_item_slice (n_items, d_slice) are its contiguous column
 * slices of the layer-0 tables, the graph handle is the whole Â (replicated).  The exchange is fused into the LAST pass:
 * its epilogue stores the layer mean (or the last layer if single) directly into the row-sharded full-width result
 * tables of the peers through peer-mapped pointers (tgcn_peer_open): user row r goes to
 * h_peer_user_out[r / users_per_rank] at local row r % users_per_rank, item rows go to every h_peer_item_out[q]
 * (n_items, d_full), both at column col_off.  After a barrier every GPU holds users_emb for its user range and the whole
 * items_emb, as base_model.py:106 returns them.  n_peers = 1 with local pointers assembles slices on one GPU. */
int tgcn_propagate_sliced(const tgcn_graph_t* g, int64_t d_slice, int32_t n_layers, int32_t single,
                          const float* d_user_slice, const float* d_item_slice, const uint8_t* d_keep, float dropout,
                          int64_t d_full, int64_t col_off, int32_t n_peers, int64_t users_per_rank,
                          float* const* h_peer_user_out, float* const* h_peer_item_out, void* d_workspace,
                          int64_t workspace_bytes, tgcn_stream_t stream);

/* The two building blocks of the GRID scheme (feature slices x user partitions; textgcn_b200.dist.GridPropagator), where
 * the hops run through tgcn_spmm_ex on row-block handles and only the last pass changes layout:
 * spmm_scatter: one pass  (add_0 + ... + Â_block·X) / divisor  whose rows go to the peers' full-width tables: user row u
 *   (global id = the block's row_begin + local row) is stored at local row (u - user_row0) % users_per_rank of
 *   h_peer_user_out[(u - user_row0) / users_per_rank] — with user_row0 = the first user of the rank's row partition and the
 *   n_peers = G feature-slice partners of that partition, the rows never leave their row group (tgcn_propagate_sliced is the
 *   user_row0 = 0, n_peers = P case); item rows go to every h_peer_item_out[q];
 * layer_mean_scatter: (add_0 + ... + add_{n_add-1}) / divisor over (n_rows, d_slice) tables, stored at column col_off of
 *   rows [row0, row0 + n_rows) of every h_dst[q] (d_full-wide) — the item table's layer mean and its all-gather in one. */
int tgcn_spmm_scatter(const tgcn_graph_t* g, int64_t d_slice, const float* d_x_user, const float* d_x_item, int32_t n_add,
                      const float* const* h_add_user, const float* const* h_add_item, float divisor, int64_t d_full,
                      int64_t col_off, int32_t n_peers, int64_t users_per_rank, int64_t user_row0,
                      float* const* h_peer_user_out, float* const* h_peer_item_out, void* d_workspace,
                      int64_t workspace_bytes, tgcn_stream_t stream);
int tgcn_layer_mean_scatter(int64_t n_rows, int64_t d_slice, int32_t n_add, const float* const* h_add, float divisor,
                            int64_t d_full, int64_t col_off, int64_t row0, int32_t n_dst, float* const* h_dst,
                            tgcn_stream_t stream);

/* Peer memory for the call above (one process per GPU): tgcn_peer_alloc cudaMalloc's `bytes` on the current device and
 * fills a 64-byte CUDA IPC handle the caller ships to the other ranks (any transport; torch.distributed here);
 * tgcn_peer_open maps a peer's handle into this process (peer access enabled lazily); close / free undo them. */
#define TGCN_PEER_HANDLE_BYTES 64
int tgcn_peer_alloc(int64_t bytes, void** d_ptr, uint8_t* h_handle);
int tgcn_peer_open(const uint8_t* h_handle, void** d_ptr);
int tgcn_peer_close(void* d_ptr);
int tgcn_peer_free(void* d_ptr);
/* Stream-ordered barrier between the n_peers GPUs that share peer-mapped flag arrays (h_peer_flags[q] = rank q's array of
 * n_peers int64 slots, zero-initialised, opened with tgcn_peer_open): returns on `stream` once every peer has reached the
 * barrier with the same `epoch` (1, 2, 3, ... — strictly increasing per call).  One tiny kernel; replaces the 1-float NCCL
 * all-reduce that bracketed the peer stores of the grid / sliced schemes.  Everything this GPU stored to peer memory on
 * `stream` before the call is visible to the peers after their barrier returns. */
int tgcn_peer_barrier(int32_t n_peers, int32_t rank, int64_t* const* h_peer_flags, int64_t epoch, tgcn_stream_t stream);

/* e (multi-GPU)  The path's collectives as helpers that take an ncclComm_t (passed as void*; any communicator of the
 * process, e.g. one created below).  unique_id / init_rank / destroy: communicator plumbing for consumers that do not run
 * torch.distributed — rank 0 of the group fills the 128-byte id, ships it over any transport, every rank calls init_rank.
 * allreduce_sum_f32: the per-hop exchange of the grid / bipartite schemes — in-place sum of the (n_items, d/G) item-table
 *   slice over the row group.  allgather_f32: the north_star's row-block scheme — every rank contributes n_per_rank floats.
 * topk_exchange: item-sharded eval — block q (rows_per_rank x k) of the partial tables goes to rank q; afterwards
 *   d_recv_* holds, block by block, every shard's candidates for THIS rank's user slice, ready for tgcn_topk_merge. */
#define TGCN_COMM_ID_BYTES 128
int tgcn_comm_unique_id(uint8_t* h_id);
int tgcn_comm_init_rank(void** comm, int32_t n_ranks, int32_t rank, const uint8_t* h_id);
int tgcn_comm_destroy(void* comm);
int tgcn_allreduce_sum_f32(void* comm, float* d_buf, int64_t n, tgcn_stream_t stream);
int tgcn_allgather_f32(void* comm, const float* d_send, float* d_recv, int64_t n_per_rank, tgcn_stream_t stream);
int tgcn_topk_exchange(void* comm, int32_t n_ranks, int64_t rows_per_rank, int32_t k, const int32_t* d_part_ids,
                       const float* d_part_scores, int32_t* d_recv_ids, float* d_recv_scores, tgcn_stream_t stream);

/* Backward of the above (what autograd does at base_model.py:125 through :148/:157): given
 * d_grad_out = dL/d(out) (N, d) computes dL/dE0 into d_grad_in (N, d) with L transposed SpMMs in
 * Horner form and no saved activations.  accumulate != 0 adds into d_grad_in instead of storing. */
int tgcn_propagate_bwd(tgcn_graph_t* g, int64_t d, int32_t n_layers, int32_t single,
                       const float* d_grad_out, const uint8_t* d_keep, float dropout, int32_t accumulate,
                       float* d_grad_in, void* d_workspace, int64_t workspace_bytes, tgcn_stream_t stream);

/* Host-buffer form of representation for callers that keep the tables in host memory: copies
 * h_user_w / h_item_w to the device staging buffers the caller provides, runs propagate_fwd and
 * copies the (N, d) result back to h_out, all on `stream`.  d_stage must hold 2·N·d floats. */
int tgcn_propagate_host(const tgcn_graph_t* g, int64_t d, int32_t n_layers, int32_t single,
                        const float* h_user_w, const float* h_item_w, float* h_out,
                        float* d_stage, void* d_workspace, int64_t workspace_bytes, tgcn_stream_t stream);

/* a7-a10  bpr_loss + reg_loss (base_model.py:166-171, :186-210) and their gradients, one kernel.
 * d_emb is the (N, d) result of propagate_fwd; d_user_w/d_item_w the layer-0 tables.  d_negs is
 * (n_neg, batch).  Loss is mean over negatives of mean(SELU(neg - pos)) (:194); the regulariser is
 * reg_lambda/(2·batch)·(‖U0[users]‖² + ‖I0[pos]‖² + ‖I0[negs]‖²_F) (:200-210).
 * Outputs: d_losses[0] = bpr, d_losses[1] = reg; dL/d(emb) is atomically ADDED into d_grad_emb (N, d)
 * and the regulariser's gradient into d_grad_w0 (N, d) (either may be NULL to skip; caller zeroes).
 * A row whose user / positive id lies outside [0, n_users) / [0, n_items) — the -1 sentinels tgcn_sample_bpr_batch writes
 * for rows it could not complete — is skipped (no loss term, no gradient, no memory access); so is a single negative with
 * such an id.  The divisors stay batch and batch·n_neg.
 * d_workspace holds per-warp partial sums: tgcn_bpr_workspace_bytes(batch). */
int64_t tgcn_bpr_workspace_bytes(int64_t batch);
int tgcn_bpr_fwd_bwd(int64_t n_users, int64_t n_items, int64_t d, int64_t batch, int32_t n_neg,
                     const int32_t* d_users, const int32_t* d_pos, const int32_t* d_negs,
                     const float* d_emb, const float* d_user_w, const float* d_item_w, float reg_lambda,
                     float* d_losses, float* d_grad_emb, float* d_grad_w0,
                     void* d_workspace, int64_t workspace_bytes, tgcn_stream_t stream);

/* a11+a12  score_batchwise + train-item masking + topk (base_model.py:173-179, :255-261) fused:
 * the (n_rank, n_items) score matrix never reaches HBM.  Ranks users d_users[0..n_rank) (NULL = ids
 * 0..n_rank-1) against items [item_begin, item_end).  Vectors: d_user_vecs row u at u·ldu, item row i
 * at i·ldi, K floats each (K % 4 == 0).  score = <user, item> (+ d_user_bias[u]) (+ d_item_bias[i]).
 * vecs_by_position != 0: user vectors and d_user_bias are indexed by list position m instead of user id
 * d_users[m] (packed LTR operands); d_users is then used for the mask only.
 * Items the user interacted with in `mask_graph` (user rows of Â = train_user_dict) are excluded;
 * when fewer than k unmasked items exist the list is completed with masked items, lowest id first,
 * score -inf.  Order is the canonical strict order (score desc, item id asc).
 * precision: 1 = exact fp32 FMA (SIMT kernel); 2 = 3xTF32 on the tcgen05 tensor cores (k <= 64; K is zero-padded to whole
 * 32-wide chunks, bias terms ride in one extra chunk; the user tile stays resident in shared memory for K <= 128 and is
 * streamed with the item tile above that; scores within ~1e-6 norm-wise of fp32, bit-exact for TF32-representable
 * inputs); 3 = screened (k <= 24): ONE TF32 product per score (raw user rows x a rounded copy of the item rows; bias terms in an extra
 * chunk) keeps the best 40 items per user by approximate score; with eps = 1.6e-3·|u|·max|i| (+ the bias terms' share) bounding the
 * TF32 error, the k best of them are re-scored in exact fp32 FMA, further entries as long as approximate score + eps reaches the
 * smallest of those exact scores, and the re-scored entries are sorted on the exact scores: provably the exact top-k unless no entry of
 * a full list can be ruled out — such rows are queued on the device and ranked again by the 3xTF32 variant in the same call (no host
 * synchronisation: the second pass reads the queue length on the device) and then given the same exact scores; 0 = screened when
 * eligible without bias terms at K <= 128 and the item range is long enough for it to pay (65 536 rows at K = 128, 98 304 at K = 96,
 * 131 072 at K <= 64: shorter sweeps are dominated by the list updates of their opening; wider contractions and bias terms have a
 * screened form too, reachable with precision 3, which does not beat 3xTF32 yet), else 3xTF32 when eligible, else fp32.  tgcn_eval_resolve_precision says what 0 resolves to for a shape.
 * Outputs (n_rank, k) int32 ids and fp32 scores.  k <= TGCN_MAX_TOPK. */
int64_t tgcn_eval_workspace_bytes(int64_t n_rank, int64_t n_items_range, int64_t K, int32_t k);
/* Byte offset, inside the eval workspace, of the int32 in which a screened call (precision 0 / 3) leaves the number of rows it
 * sent to its second pass; -1 when the screened path does not apply to this shape.  Diagnostics only (bench, tests). */
int64_t tgcn_eval_screen_queue_offset(int64_t n_rank, int64_t n_items_range, int64_t K, int32_t k, int32_t has_bias);
/* The precision (1, 2 or 3) a tgcn_eval_topk call with this shape runs at: `precision` itself unless it is 0 (auto). */
int32_t tgcn_eval_resolve_precision(int64_t n_items_range, int64_t K, int32_t k, int32_t has_bias, int32_t precision);
int tgcn_eval_topk(const tgcn_graph_t* mask_graph, int64_t n_rank, const int32_t* d_users,
                   const float* d_user_vecs, int64_t ldu, const float* d_item_vecs, int64_t ldi, int64_t K,
                   int64_t item_begin, int64_t item_end, const float* d_user_bias, const float* d_item_bias,
                   int32_t vecs_by_position, int32_t precision, int32_t k, int32_t finalize, int32_t* d_out_ids,
                   float* d_out_scores,
                   void* d_workspace, int64_t workspace_bytes, tgcn_stream_t stream);

/* Cross-shard / cross-GPU merge of n_parts partial top-k tables (n_parts, n_rows, k) under the same
 * order; finalize != 0 applies the G9 completion using mask_graph and d_users. */
int tgcn_topk_merge(const tgcn_graph_t* mask_graph, int64_t n_rows, const int32_t* d_users, int32_t n_parts,
                    int32_t k, const int32_t* d_part_ids, const float* d_part_scores, int32_t finalize,
                    int32_t* d_out_ids, float* d_out_scores, tgcn_stream_t stream);

/* a15+a16  AdvSamplModel.score_pairwise_adv + per-user sort / positive removal / top max(k)
 * (advanced_sampling.py:37-44, :61-65; utils.py:121-128) as one kernel: for each batch row scores its
 * n_cand candidates, orders them (score desc, candidate position asc), drops the user's train items and
 * writes the first kmax to d_out_negs (batch, kmax), -1 padded, with counts in d_out_counts. */
int tgcn_adv_select(const tgcn_graph_t* mask_graph, int64_t d, int64_t batch, int32_t n_cand,
                    const int32_t* d_users, const int32_t* d_cands, const float* d_emb, int32_t kmax,
                    int32_t* d_out_negs, int32_t* d_out_counts, float* d_out_scores, tgcn_stream_t stream);

/* a11 / a15 / a18 / a19 / a20 as CALLABLES with the reference's signatures, for callers that want the dense result
 * rather than the fused top-k (predict / get_loss never come through here: tgcn_eval_topk / tgcn_adv_select fuse these
 * products so that the intermediates never reach HBM).  Exact fp32 FMA arithmetic.
 * score_batchwise:  out[m·ldo_row + n·ldo_col] = <A[m, :K], B[n, :K]> (+ d_row_bias[m]) (+ d_col_bias[n]) — BaseModel.
 *   score_batchwise (base_model.py:173-179: A = users_emb batch, B = items_emb, ldo = (n_items, 1)); one plane of
 *   LTRBase.get_features_batchwise (ltr_models.py:131-146: ldo = (5·n_items, 5)); LTRLinear(WPop).score_batchwise_ltr
 *   (ltr_models.py:200-204, :227-232) on the packed operands of tgcn_ltr_pack_* with the head's bias terms.
 * score_pairwise_adv: out[b, c] = <users_emb[b], items_emb[b, c]> for (batch, d) x (batch, n_cand, d)
 *   (AdvSamplModel.score_pairwise_adv, advanced_sampling.py:37-44; shape stays (batch, n_cand), no squeeze: G14).
 * ltr_features_rows: the five dot products of ROW-ALIGNED vectors, reference order [emb·emb, rev·rev, desc·desc, rev·desc,
 *   desc·rev] (LTRBase.get_features_pairwise(u_vecs, i_vecs), ltr_models.py:148-166) into d_out[b·ldo + 0..4]. */
int tgcn_score_batchwise(int64_t n_rows, int64_t n_cols, int64_t K, const float* d_a, int64_t lda, const float* d_b,
                         int64_t ldb, const float* d_row_bias, const float* d_col_bias, float* d_out, int64_t ldo_row,
                         int64_t ldo_col, tgcn_stream_t stream);
int tgcn_score_pairwise_adv(int64_t batch, int32_t n_cand, int64_t d, const float* d_users_emb, int64_t ldu,
                            const float* d_items_emb, float* d_out, tgcn_stream_t stream);
int tgcn_ltr_features_rows(int64_t batch, int64_t d, int64_t D, const float* d_ue, int64_t ld_ue, const float* d_ie,
                           int64_t ld_ie, const float* d_ur, int64_t ld_ur, const float* d_ud, int64_t ld_ud,
                           const float* d_ir, int64_t ld_ir, const float* d_id, int64_t ld_id, float* d_out, int64_t ldo,
                           tgcn_stream_t stream);

/* a13 / n4  utils.calculate_metrics (utils.py:11-63) on the device: d_pred_ids is the (n_rows, kmax) id table of
 * tgcn_eval_topk, the true test items of row r are d_true_ids[d_true_ptr[r] .. d_true_ptr[r+1]) (true_test_lil as a CSR).
 * Writes d_out[ki·5 + m], m = {recall, precision, hit, ndcg, f1} at k = h_ks[ki], each the float64 mean over rows.
 * Deterministic (fixed-order reduction).  d_workspace: tgcn_topk_metrics_workspace_bytes(). */
int64_t tgcn_topk_metrics_workspace_bytes(void);
int tgcn_topk_metrics(int64_t n_rows, int32_t kmax, const int32_t* d_pred_ids, const int64_t* d_true_ptr,
                      const int32_t* d_true_ids, int32_t n_ks, const int32_t* h_ks, double* d_out, void* d_workspace,
                      int64_t workspace_bytes, tgcn_stream_t stream);

/* a17-a21  LTR feature assembly (ltr_models.py:116-166, :200-241).
 * pairwise_features: for pair b = (users[b], items[b]) writes the 5 raw dot products
 *   [emb·emb, rev·rev, desc·desc, rev·desc, desc·rev] (:154-163) to d_out (batch, n_feat) columns 0..4,
 *   and, if n_feat == 7, popularity_users[u], popularity_items[i] to columns 5, 6 (:234-241).
 * pairwise_emb_bwd: scatter-adds d(feature 0)/d(emb) given d_gf0 (batch) into d_grad_emb (N, d).
 * pack_items / pack_users: operands of the collapsed batchwise score (the head has no activation,
 *   ltr_models.py:186-190): item row = [w0·Ie | w1·Ir + w3·Id | w2·Id + w4·Ir], user row = [Ue | Ur | Ud],
 *   so score_batchwise_ltr (:200-204) is one contraction of width d + 2·D fed to tgcn_eval_topk. */
int tgcn_ltr_pairwise_features(int64_t n_users, int64_t d, int64_t D, int64_t batch, int32_t n_feat,
                               const int32_t* d_users, const int32_t* d_items, const float* d_emb,
                               const float* d_users_rev, const float* d_users_desc,
                               const float* d_items_rev, const float* d_items_desc,
                               const float* d_pop_users, const float* d_pop_items,
                               float* d_out, tgcn_stream_t stream);
int tgcn_ltr_pairwise_emb_bwd(int64_t n_users, int64_t d, int64_t batch, const int32_t* d_users,
                              const int32_t* d_items, const float* d_emb, const float* d_gf0,
                              float* d_grad_emb, tgcn_stream_t stream);
int tgcn_ltr_pack_items(int64_t n_items, int64_t d, int64_t D, const float* d_items_emb,
                        const float* d_items_rev, const float* d_items_desc, const float* h_w5,
                        float* d_out, tgcn_stream_t stream);
int tgcn_ltr_pack_users(int64_t n_rank, const int32_t* d_users, int64_t d, int64_t D,
                        const float* d_users_emb, const float* d_users_rev, const float* d_users_desc,
                        float* d_out, tgcn_stream_t stream);

/* n1 (next row)  GPU negative sampler replacing BaseDataset._cache_samples / __getitem__ (dataset.py:167-193):
 * for each of `batch` user ids writes the int64 row [user, uniform positive, n_neg uniform non-positive items
 * (distinct within the row)] into d_out (batch, 2 + n_neg).  Counter-based RNG keyed by (seed, row).  Rows for which
 * no negative exists get -1 and bump *d_fail_count (the reference loops forever there). */
int tgcn_sample_bpr_batch(const tgcn_graph_t* g, int64_t batch, int32_t n_neg, const int32_t* d_users, uint64_t seed,
                          int32_t max_tries, int64_t* d_out, int32_t* d_fail_count, tgcn_stream_t stream);

/* n2 (next row)  the edge-dropout keep mask of base_model.py:82 drawn on the device: d_keep[p] = 1 with probability
 * 1 - dropout (counter-based hash of (seed, p)), one byte per nnz, consumed by tgcn_propagate_fwd/_bwd/_spmm_ex. */
int tgcn_dropout_mask(int64_t nnz, float dropout, uint64_t seed, uint8_t* d_keep, tgcn_stream_t stream);

/* a14 / n1  AdvSamplDataset.__getitem__ (advanced_sampling.py:21-22): row b of d_out (batch, 1 + n_cand) int64 is
 * [d_users[b], n_cand distinct uniform item ids] — the head of a keyed random permutation of the items. */
int tgcn_sample_candidates(int64_t n_items, int64_t batch, int32_t n_cand, const int32_t* d_users, uint64_t seed,
                           int64_t* d_out, tgcn_stream_t stream);

/* advanced_sampling.py:63-64: min(n_pos, deg(u)) distinct random positives of every batch user into d_out
 * (batch, n_pos) int64, -1 padded — the device counterpart of random.sample(positives, 5). */
int tgcn_sample_positives(const tgcn_graph_t* g, int64_t batch, int32_t n_pos, const int32_t* d_users, uint64_t seed,
                          int64_t* d_out, tgcn_stream_t stream);

/* n3 (next row)  dense Adam step over one table, torch.optim.Adam defaults (base_model.py:111, :126):
 * p, m, v updated in place from g; step is the 1-based step count. */
int tgcn_adam_step(int64_t n, float* d_p, const float* d_g, float* d_m, float* d_v, float lr, float beta1,
                   float beta2, float eps, int64_t step, tgcn_stream_t stream);

/* The same two with their per-step scalars in DEVICE memory, so that a whole training step (mask draw, propagate, fused
 * BPR, backward, Adam) can be captured once in a CUDA graph and replayed (textgcn_b200.train_graph.GraphedTrainStep):
 * adam_prepare does ++*d_step and writes d_bc = {1 - beta1^step, sqrt(1 - beta2^step)}; adam_step_dev reads d_bc;
 * counter_inc does ++*d_counter; dropout_mask_dev draws with seed = base_seed + *d_draws · 0xD6E8FEB86659FD93. */
int tgcn_adam_prepare(int64_t* d_step, float* d_bc, float beta1, float beta2, tgcn_stream_t stream);
int tgcn_adam_step_dev(int64_t n, float* d_p, const float* d_g, float* d_m, float* d_v, float lr, float beta1, float beta2,
                       float eps, const float* d_bc, tgcn_stream_t stream);
int tgcn_counter_inc(uint64_t* d_counter, tgcn_stream_t stream);
int tgcn_dropout_mask_dev(int64_t nnz, float dropout, uint64_t base_seed, const uint64_t* d_draws, uint8_t* d_keep,
                          tgcn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TGCN_B200_H */
