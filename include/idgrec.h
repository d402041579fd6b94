/* idgrec.h -- C ABI of libidgrec_sm100.so, the B200 (sm_100a) implementation of
 * the ID-GRec hot path: normalised-adjacency CSR build -> K-layer propagation
 * (SpMM) and its backward -> fused BPR / InfoNCE losses -> full-ranking top-K.
 *
 * The reference (BlueGhostYi/ID-GRec) is pure Python and has no FFI: every entry
 * point below replaces a *library call made from* the cited reference lines
 * (torch.sparse.mm, torch.matmul, torch.topk, scipy dok/lil algebra, the numpy
 * sampling loop).  The Python host mirror (id-grec_b200/{models,utility}) binds
 * these symbols with ctypes (id-grec_b200/idgrec/_lib.py); INTEGRATION.md shows
 * the same binding applied directly inside the reference's files.
 *
 * Conventions
 *   - every function returns 0 on success, <0 for an argument error, >0 for a
 *     cudaError_t passed through; idg_last_error() returns a thread-local text.
 *   - all `d_*` pointers are DEVICE pointers on the current CUDA device, row-major,
 *     contiguous, caller-allocated and caller-owned; `h_*` are HOST pointers.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and
 *     never synchronise unless the comment says so.
 *   - embeddings are fp32; ids are int64 at the API (the reference's LongTensors)
 *     and int32 inside CSR structures (scipy's index dtype, data_graph.py:33-55).
 */
#ifndef IDGREC_H_
#define IDGREC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct idg_graph idg_graph; /* opaque device CSR + schedule */
typedef struct idg_peers idg_peers; /* opaque table of IPC-mapped peer slabs (multi-GPU) */
/* Adam fused into the epilogue of the last backward propagation layer (trainer.py:54-56 without a gradient pass):
 * p/m/v [N,d] tables, regc [N] per-row L2-reg coefficient (written by idg_bpr_backward, may be NULL),
 * d_scalars = {lr/(1-beta1^t), sqrt(1-beta2^t)} produced by idg_adam_prepare for the current step. */
typedef struct idg_adam_args {
    float* p; float* m; float* v;
    const float* regc;
    const float* d_scalars;
    float beta1, beta2, eps;
} idg_adam_args;

/* Scalar work of one captured train step folded into the single-thread epilogue of the BPR loss reduction
 * (idg_bpr_forward_tail): d_loss_acc[0..1] += {bpr, reg} (epoch sums, read once per epoch instead of trainer.py:52's
 * .item() per batch) and, when d_step is given, the job of idg_adam_prepare for this step.  Any pointer may be NULL. */
typedef struct idg_step_tail {
    double* d_loss_acc; /* float64, like the reference's Python-float sums of .item() values (trainer.py:52-53) */
    int32_t* d_step;
    float* d_scalars;
    float lr, beta1, beta2;
} idg_step_tail;

int idg_version(void);
/* d_acc[i] += (double)d_x[i], i < n: per-loss epoch sums kept on the device in float64 (the reference adds
 * loss.item() values into Python floats, trainer.py:52-53), read once per epoch. */
int idg_accumulate_f64(double* d_acc, const float* d_x, int32_t n, void* stream);
const char* idg_last_error(void);
/* number of kernels this library has launched so far in this process (bench.py "gpu_launches") */
int64_t idg_launch_count(void);

/* ---- a2: utility/utility_data/data_graph.py:7-55 -------------------------
 * D^-1/2 [[0,R],[R^T,0]] D^-1/2 (add_self: +I before normalising) as canonical
 * CSR: rows ascending, columns ascending inside a row, duplicate (u,i) pairs
 * merged with their multiplicity as weight (data_loader.py:42-43).  Two phases,
 * because the reference's d = np.power(deg, -0.5) is numpy's own powf/pow and is
 * not correctly rounded: the host computes d from d_deg with
 * numpy, exactly like data_graph.py:46, and hands it back.
 *   structure: d_user/d_item are the E train pairs (int64, device).  Outputs hold
 *     up to 2E(+N) entries; d_mult = multiplicity a (1.0, 2.0, ...), d_deg[N] = row
 *     sums (incl. the self loop when add_self).  *h_nnz = merged entry count.
 *     SYNCHRONISES the stream once (to return nnz).
 *   normalise: data[k] = (d[row]*a[k])*d[col] evaluated in fp32 from d_dinv32
 *     (no-self variant, data_graph.py:48-51) or in fp64 from d_dinv64 and rounded
 *     to fp32 once (with-self variant + tools.py:101).  Exactly one of the two
 *     dinv pointers is non-NULL. */
int idg_csr_structure(const int64_t* d_user, const int64_t* d_item, int64_t E, int32_t U, int32_t I,
                      int add_self, int32_t* d_indptr, int32_t* d_indices, float* d_mult, double* d_deg,
                      int64_t* h_nnz, void* stream);
int idg_csr_normalise(const int32_t* d_indptr, const int32_t* d_indices, const float* d_mult, int32_t n_rows,
                      int64_t nnz, const float* d_dinv32, const double* d_dinv64, float* d_data, void* stream);

/* ---- graph handle: replaces `self.Graph` (models/LightGCN.py:30-32) -------
 * Takes rows [0,n_rows) of a CSR whose column space is [0,n_cols).  For a
 * row-partitioned rank pass its slice (indptr rebased to 0) and set row_offset to
 * the first global row: outputs are then written at Y[row_offset + r].  Builds the
 * degree-sorted, chunked work schedule.  SYNCHRONISES the stream (one-off). */
int idg_graph_create(const int32_t* d_indptr, const int32_t* d_indices, const float* d_data,
                     int32_t n_rows, int32_t n_cols, int64_t nnz, int32_t row_offset,
                     idg_graph** out, void* stream);
void idg_graph_destroy(idg_graph* g);
int64_t idg_graph_nnz(const idg_graph* g);
int32_t idg_graph_rows(const idg_graph* g);
/* schedule classes of the handle: 2 when the work items run "all rows that gather from above their own index, then all rows that
 * gather from below" (bipartite adjacency whose gather table exceeds the L2: user rows, then item rows), else 1.  Scheduling
 * only -- results are bit-identical (IDG_SPMM_CLASS_SPLIT = 0 | 1 overrides the size rule). */
int32_t idg_graph_classes(const idg_graph* g);

/* ---- a6/a7/a11: one propagation layer = torch.sparse.mm(self.Graph, X) -----
 * (models/LightGCN.py:44, SimGCL.py:48, XSimGCL.py:51, NGCF.py:85) with the
 * element-wise work that follows it in the reference fused into the epilogue:
 *   y   = sum_k data[k] * X[col[k]]              (fixed order, no float atomics)
 *   y  += addend[row]                             if d_addend  (Horner backward h <- G + A h)
 *   y  += sign(y) * noise[row]/max(|noise[row]|_2,1e-12) * eps   if d_noise (SimGCL.py:49-51)
 *   Y[row]       = y                              if d_Y
 *   acc_out[row] = ((d_acc_in ? acc_in[row] : 0) + y) / acc_div   if d_acc_out (layer mean, LightGCN.py:47-48)
 * d in {32, 64, 128}.  X must not alias Y / acc_out. */
int idg_spmm_layer(const idg_graph* g, const float* d_X, float* d_Y, const float* d_addend,
                   const float* d_noise, float eps, const float* d_acc_in, float* d_acc_out,
                   float acc_div, int32_t d, void* stream);

/* K-layer forward in one call (single-GPU convenience; same kernels).
 *   include_layer0 = 1: mean over X0..XK (LightGCN.py:41); 0: X1..XK (SimGCL.py:45).
 *   d_noise: K stacked [N,d] U[0,1) tensors or NULL.  cl_layer > 0 also copies the
 *   post-noise output of layer cl_layer into d_out_cl (XSimGCL.py:57-58).
 *   d_work: scratch of 2*N*d floats. */
int idg_propagate_fwd(const idg_graph* g, const float* d_X0, int32_t d, int32_t K, int include_layer0,
                      const float* d_noise, float eps, int32_t cl_layer, float* d_out_mean,
                      float* d_out_cl, float* d_work, void* stream);
/* Backward of the above w.r.t. X0 (A symmetric => same CSR; trainer.py:55 autograd):
 *   gX0 = (inc0*G + A(G + A(G + ... A G)))/cnt  [+ A^cl_layer Gcl]   cnt = K + inc0. */
int idg_propagate_bwd(const idg_graph* g, const float* d_G, const float* d_Gcl, int32_t d, int32_t K,
                      int include_layer0, int32_t cl_layer, float* d_gX0, float* d_work, void* stream);

/* ---- work the loss never looks at (identical results, SURVEY.md 8 d "row-restricted last layer") ----
 * idg_batch_rows: unique rows {user, U+pos, U+neg} of a mini-batch as a list (order unspecified) with
 * its length in *d_count, and as bits set in d_bitmap (ceil(N/32) words, all zero on entry).
 * idg_batch_rows_clear zeroes those bits again.  d_rowlist holds up to 3B ints. */
int idg_batch_rows(const int64_t* d_user, const int64_t* d_pos, const int64_t* d_neg, int32_t B, int32_t U,
                   int32_t* d_rowlist, int32_t* d_count, uint32_t* d_bitmap, void* stream);
/* same + the distinct users and (U + positive item) rows of the batch, in order of first appearance, counts on
 * the device: replaces torch.unique(user) / torch.unique(positive) (SimGCL.py:80-81) without a host sync */
int idg_batch_rows_unique(const int64_t* d_user, const int64_t* d_pos, const int64_t* d_neg, int32_t B, int32_t U,
                          int32_t* d_rowlist, int32_t* d_count, uint32_t* d_bitmap, unsigned char* d_lead,
                          int64_t* d_uidx, int32_t* d_ucnt, int64_t* d_iidx, int32_t* d_icnt, void* stream);
int idg_batch_rows_clear(const int32_t* d_rowlist, const int32_t* d_count, int32_t max_rows, uint32_t* d_bitmap,
                         void* stream);
/* Neighbourhood of the batch ("closure" = batch rows + rows with a batch neighbour) as a bitmap, all-zero on entry;
 * computed for the rows of the handle (a row-partitioned rank computes its slice).  Once registered with
 * idg_graph_set_closure (NULL to clear), idg_propagate_fwd_ex evaluates layer K-1 only on closure rows (the batch-
 * restricted last layer reads nothing else) and idg_propagate_bwd_ex/_adam gather only closure columns in the
 * second backward product (H_{K-1} = G + A.G is zero elsewhere).  Pays off when the closure is a small part of the
 * graph (1M x 1M scale-up: ~20 %); identical results.  idg_spmm_layer_masked is the masked layer alone. */
int idg_closure_bitmap(const idg_graph* g, const uint32_t* d_batch_bitmap, uint32_t* d_closure, void* stream);
int idg_graph_set_closure(idg_graph* g, const uint32_t* d_closure);
/* The same bitmap from the batch side (the structure of A_hat is symmetric: rows with a neighbour in the batch = union of the
 * column lists of the batch rows): cost ~ batch rows x degree instead of one pass over all nonzeros.  Whole-graph handle only. */
int idg_closure_from_rows(const idg_graph* g, const int32_t* d_rowlist, const int32_t* d_count, int32_t max_rows,
                          const uint32_t* d_batch_bitmap, uint32_t* d_closure, void* stream);
int idg_spmm_layer_masked(const idg_graph* g, const float* d_X, float* d_Y, const float* d_noise, float eps,
                          const float* d_acc_in, float* d_acc_out, float acc_div, int32_t d,
                          const uint32_t* d_rowmask, void* stream);
/* scratch ints needed by the row-restricted entry points for up to max_rows listed rows */
int64_t idg_graph_worklist_ints(const idg_graph* g, int32_t max_rows);
/* idg_spmm_layer evaluated only on the listed rows (other rows of the outputs are left untouched);
 * the layer sum may take up to three inputs: acc_out = (((acc_in + acc_in2) + acc_in3) + y) / acc_div */
int idg_spmm_layer_rows(const idg_graph* g, const float* d_X, float* d_Y, const float* d_noise, float eps,
                        const float* d_acc_in, const float* d_acc_in2, const float* d_acc_in3, float* d_acc_out,
                        float acc_div, int32_t d, const int32_t* d_rowlist, const int32_t* d_count,
                        int32_t max_rows, int32_t* d_worklist, void* stream);
/* idg_spmm_layer for an X that is zero outside the rows flagged in d_bitmap: streams the CSR structure
 * but gathers only flagged columns (first backward layer: dL/dF is non-zero on the batch rows only).
 * skip_zero_rows: rows whose result is exactly zero are not stored (d_Y must be zero-filled by the caller);
 * on the multi-GPU path this keeps them off NVLink. */
int idg_spmm_layer_sparse_in(const idg_graph* g, const float* d_X, float* d_Y, const float* d_addend,
                             const float* d_acc_in, float* d_acc_out, float acc_div, int32_t d,
                             const uint32_t* d_bitmap, int skip_zero_rows, void* stream);
/* idg_propagate_fwd whose LAST layer and layer mean are evaluated only on the listed rows
 * (d_rowlist NULL => identical to idg_propagate_fwd); idg_propagate_bwd whose FIRST product uses the
 * sparse-input kernel (d_bitmap NULL => identical to idg_propagate_bwd). */
int idg_propagate_fwd_ex(const idg_graph* g, const float* d_X0, int32_t d, int32_t K, int include_layer0,
                         const float* d_noise, float eps, int32_t cl_layer, float* d_out_mean, float* d_out_cl,
                         float* d_work, const int32_t* d_rowlist, const int32_t* d_count, int32_t max_rows,
                         int32_t* d_worklist, void* stream);
int idg_propagate_bwd_ex(const idg_graph* g, const float* d_G, const float* d_Gcl, int32_t d, int32_t K,
                         int include_layer0, int32_t cl_layer, float* d_gX0, float* d_work,
                         const uint32_t* d_bitmap, void* stream);

/* ---- a9: models/LightGCN.py:54-72 + utility_function/losses.py:4-21 ---------
 * Fused gather + dot + -log(sigmoid+1e-7) + L2-reg, forward and backward.
 *   F  [N,d]  final embeddings (users rows 0..U-1, items U..N-1)
 *   E0 [N,d]  ego embeddings (cat(user_w,item_w))
 *   d_loss[0] = bpr, d_loss[1] = reg_lambda * reg     (losses.py:13,19-21)
 *   d_G  [N,d]: dL/dF.  Only the rows touched by the batch are written (each row is
 *               the ordered sum of its samples' contributions: deterministic, no
 *               float atomics); every other row must already be ZERO.  The caller
 *               zero-fills G once; idg_bpr_finish re-zeroes the touched rows.
 * reg_mask bit0/1/2 = include user/pos/neg ego rows (NGCF.py:120-125 uses 0b110).
 * d_ws: workspace of idg_bpr_workspace_bytes(B) bytes, private to one forward/backward/finish triple. */
int64_t idg_bpr_workspace_bytes(int32_t B);
/* forward: d_loss[0] = bpr, d_loss[1] = reg_lambda * reg; keeps c_b and the row keys in d_ws */
int idg_bpr_forward(const float* d_F, const float* d_E0, const int64_t* d_user, const int64_t* d_pos,
                    const int64_t* d_neg, int32_t B, int32_t U, int32_t N, int32_t d, float reg_lambda,
                    int reg_mask, float* d_loss, void* d_ws, void* stream);
/* backward: writes the touched rows of d_G = upstream[0] * d bpr / dF.  d_upstream: device float[2]
 * = (dL/d bpr, dL/d reg) handed down by autograd, or NULL for (1, 1).  d_regc (may be NULL): [N] floats, all
 * zero on entry; receives upstream[1]*reg_lambda/B*multiplicity for the touched rows (Adam-fused path). */
int idg_bpr_forward_tail(const float* d_F, const float* d_E0, const int64_t* d_user, const int64_t* d_pos,
                         const int64_t* d_neg, int32_t B, int32_t U, int32_t N, int32_t d, float reg_lambda, int reg_mask,
                         float* d_loss, const idg_step_tail* tail, void* d_ws, void* stream);
int idg_bpr_backward(const float* d_F, int32_t B, int32_t d, int reg_mask, const float* d_upstream,
                     float* d_G, float reg_lambda, float* d_regc, void* d_ws, void* stream);
/* After the propagation backward has consumed G: adds the L2-reg gradient
 * (upstream[1] * reg_lambda/B * multiplicity * E0[row]) into d_gE0 (may be NULL) and zeroes the
 * touched rows of d_G and of d_regc (each may be NULL).  Same B / d_ws as the matching forward/backward calls. */
int idg_bpr_finish(const float* d_E0, float* d_gE0, float* d_G, int32_t B, int32_t d, float reg_lambda,
                   const float* d_upstream, float* d_regc, void* d_ws, void* stream);

/* small utilities used between the fused kernels */
/* idg_bpr_finish that also clears the step's batch-row bitmap (the rows it visits are exactly the rows
 * idg_batch_rows listed), replacing a separate idg_batch_rows_clear launch. */
int idg_bpr_finish_clear(const float* d_E0, float* d_gE0, float* d_G, int32_t B, int32_t d, float reg_lambda,
                         const float* d_upstream, float* d_regc, uint32_t* d_bitmap, void* d_ws, void* stream);
int idg_axpby(float* d_out, float a, const float* d_x, float b, const float* d_y, int64_t n, void* stream);
int idg_zero_rows(float* d_buf, const int64_t* d_idx, int32_t n, int32_t d, void* stream);
/* rows of d_buf [n_rows, d] whose bit is set in d_bitmap are zeroed (the row-partitioned step re-zeroes only the rows its
 * sparse first backward product can have written: the batch neighbourhood, not the whole [N,d] buffer) */
int idg_zero_rows_bitmap(float* d_buf, const uint32_t* d_bitmap, int32_t n_rows, int32_t d, void* stream);

/* ---- a10: utility_function/losses.py:24-35 (in-batch InfoNCE) --------------
 * V1,V2 [N,d] views; d_idx n sorted-unique row ids (torch.unique, SimGCL.py:80-81)
 * already offset into [0,N).  loss_scale multiplies the loss and the gradients
 * (ssl_lambda).  d_loss[0] += loss_scale * mean(-log(pos/ttl + 1e-5)).
 * Gradients are ACCUMULATED into rows d_idx of d_gV1 / d_gV2 (both [N,d]).
 * The n x n matrix is never materialised (tile-wise recompute in shared memory). */
int64_t idg_infonce_workspace_bytes(int32_t n, int32_t d);
int idg_infonce_fwd_bwd(const float* d_V1, const float* d_V2, const int64_t* d_idx, int32_t n, int32_t d,
                        float temperature, float loss_scale, float* d_loss, float* d_gV1, float* d_gV2,
                        void* d_ws, void* stream);

/* Device-side variants for CUDA-graph capture of the contrastive steps: idg_unique_rows = torch.unique(ids)
 * + offset (sorted distinct values, count in *d_out_count, n <= 4096, no host sync); idg_infonce_fwd_bwd_dev takes
 * the row count from the device (*d_n <= n_max; workspace sized with idg_infonce_workspace_bytes(n_max, d)). */
int idg_unique_rows(const int64_t* d_ids, int32_t n, int64_t offset, int64_t* d_out, int32_t* d_out_count, void* stream);
int idg_infonce_fwd_bwd_dev(const float* d_V1, const float* d_V2, const int64_t* d_idx, const int32_t* d_n, int32_t n_max,
                            int32_t d, float temperature, float loss_scale, float* d_loss, float* d_gV1, float* d_gV2,
                            void* d_ws, void* stream);

/* ---- a8: models/NGCF.py:87-106, dense part of one layer (side = A_hat.E comes from idg_spmm_layer) ----------
 * forward:  S = side W_gcn + b_gcn + (E*side) W_bi + b_bi;  D = LeakyReLU_0.2(S) * keep/(1-p);  out = D/max(|D|,1e-12)
 *           d_keep: [N,64] 0/1 dropout draws or NULL (no dropout); out rows have stride out_stride floats (a 64-column
 *           block of the [N,256] concat or a plain [N,64] tensor).  D is kept for the backward; d_S (the pre-activation)
 *           may be NULL in both calls -- the tensor-core backward reads the sign of S off D (same sign wherever keep = 1,
 *           zero gradient wherever keep = 0); only the CUDA-core cross-check (IDG_NGCF_BWD=fma) needs it.
 * backward: given dO (stride dO_stride) and dD_ext (gradient reaching D from the next layer, may be NULL):
 *           dside, dE_direct [N,64] (the caller finishes dE = dE_direct + A_hat.dside with idg_spmm_layer),
 *           dW_gcn, dW_bi [64,64], db [64] (same for both biases).  d_ws: idg_ngcf_workspace_bytes() bytes. */
int idg_ngcf_dense_fwd(const float* d_E, const float* d_side, const float* d_Wg, const float* d_bg, const float* d_Wb,
                       const float* d_bb, const float* d_keep, float drop_p, int32_t N, float* d_S, float* d_D,
                       float* d_out, int32_t out_stride, void* stream);
int64_t idg_ngcf_workspace_bytes(void);
/* The dropout draws of one NGCF step (NGCF.py:99-100, nn.Dropout constructed inline = always active) for all layers in one
 * launch: d_keep [n_layers, per_layer] floats, keep = 1 with probability h_keep_prob[l] = 1 - p_l.  Philox stream
 * (seed, subsequence = element quad, offset = *d_step): a captured step replays with fresh, launch-geometry independent draws. */
int idg_ngcf_keep_masks(float* d_keep, int64_t per_layer, int32_t n_layers, const float* h_keep_prob, uint64_t seed,
                        const int32_t* d_step, void* stream);
/* dst[idx[i] + row_offset, 0:d] = src[idx[i] + row_offset, 0:d] with row strides in floats (the ego block of NGCF's [N,256]
 * concat is read at the batch rows only). */
int idg_copy_rows_strided(const float* d_src, int32_t src_stride, const int64_t* d_idx, int32_t n, int32_t row_offset, int32_t d,
                          float* d_dst, int32_t dst_stride, void* stream);
int idg_ngcf_dense_bwd(const float* d_E, const float* d_side, const float* d_Wg, const float* d_Wb, const float* d_keep,
                       float drop_p, const float* d_S, const float* d_D, const float* d_dO, int32_t dO_stride,
                       const float* d_dD_ext, int32_t N, float* d_dside, float* d_dE_direct, float* d_dWg, float* d_dWb,
                       float* d_db, void* d_ws, void* stream);
/* The same three calls with the dropout draws packed 64 bits per row (d_bits [n_layers][N][2] words: word w, bit b = column
 * 32 w + b; same Philox stream as idg_ngcf_keep_masks, draw for draw): the dense kernels read 8 bytes per row instead of a
 * 256-byte float mask row.  Tensor-core kernels only; no pre-activation argument. */
int idg_ngcf_keep_bits(uint32_t* d_bits, int32_t N, int32_t n_layers, const float* h_keep_prob, uint64_t seed, const int32_t* d_step,
                       void* stream);
int idg_ngcf_dense_fwd_bits(const float* d_E, const float* d_side, const float* d_Wg, const float* d_bg, const float* d_Wb,
                            const float* d_bb, const uint32_t* d_keep_bits, float drop_p, int32_t N, float* d_S, float* d_D,
                            float* d_out, int32_t out_stride, void* stream);
int idg_ngcf_dense_bwd_bits(const float* d_E, const float* d_side, const float* d_Wg, const float* d_Wb, const uint32_t* d_keep_bits,
                            float drop_p, const float* d_D, const float* d_dO, int32_t dO_stride, const float* d_dD_ext, int32_t N,
                            float* d_dside, float* d_dE_direct, float* d_dWg, float* d_dWb, float* d_db, void* d_ws, void* stream);

/* ---- a13/a14: get_rating_for_test + Test (models/LightGCN.py:74-80,
 * utility_train/batch_test.py:52-68) fused: score = <Fu[user], Fi[item]>, train
 * positives removed, top-K by (score desc, item id asc).  No [b, I] matrix is
 * materialised.  Exactness: candidates come from an fp32 (or tensor-core) pass,
 * the survivors are rescored with the fp64 sequential dot product of the fp32
 * embeddings and every user whose candidate margin cannot be proven is recomputed
 * exhaustively, so ids equal the exact-rank oracle bit for bit.
 *   d_users [nu] int64 user ids; mask CSR = user_item_net (int32, U rows, sorted)
 *   d_out_ids [nu,K] int64, d_out_scores [nu,K] fp32 (may be NULL).  Any d >= 1 and any K <= I are accepted (the
 *   reference takes any embedding_size / top_K): d in {32, 64, 128, 256} with K <= 48 run the tiled candidate
 *   kernels (d = 64 on tcgen05); every other shape ranks each user exhaustively in fp64 -- same ids, slower. */
int64_t idg_eval_workspace_bytes(int32_t nu, int32_t I, int32_t d, int32_t K);
int idg_eval_topk(const float* d_Fu, const float* d_Fi, int32_t U, int32_t I, int32_t d,
                  const int32_t* d_mask_indptr, const int32_t* d_mask_indices, const int64_t* d_users,
                  int32_t nu, int32_t K, int64_t* d_out_ids, float* d_out_scores, void* d_ws, void* stream);

/* Self-check of the tensor-core candidate pass (d = 64, nu <= 128): writes the raw accumulators of the first item tile,
 * d_out_tile[128,128] (row = position in d_users, column = item id 0..127), i.e. the per-item UPPER bounds
 * s~ + c |u||i| the epilogue filters on.  Test infrastructure for the exactness proof; Test() never calls it. */
int idg_eval_tc_bounds(const float* d_Fu, const float* d_Fi, int32_t U, int32_t I, int32_t d, const int32_t* d_mask_indptr,
                       const int32_t* d_mask_indices, const int64_t* d_users, int32_t nu, float* d_out_tile, void* d_ws,
                       void* stream);

/* get_rating_for_test as the reference writes it (models/LightGCN.py:74-80): d_out [nu, I] = sigmoid(Fu[users] . Fi^T).
 * API completeness only: the evaluator ranks through idg_eval_topk and never materialises this matrix. */
int idg_rating_matrix(const float* d_Fu, const float* d_Fi, const int64_t* d_users, int32_t nu, int32_t I, int32_t d,
                      float* d_out, void* stream);

/* metrics.py:4-58 + batch_test.py:80-91 on device: sums over users of
 * recall/precision/ndcg at each k in h_ks (nk <= 8) -> d_sums[3*nk] (float64).
 * test CSR = test_dict rows (int32, sorted inside a row, duplicates kept: metrics.py:27 uses
 * the list length).  d_ws: idg_eval_workspace_bytes(nu, ...) bytes (may be the top-k workspace). */
int idg_eval_metrics(const int64_t* d_topk_ids, const int64_t* d_users, int32_t nu, int32_t K,
                     const int32_t* d_test_indptr, const int32_t* d_test_indices, const int32_t* h_ks,
                     int32_t nk, double* d_sums, void* d_ws, void* stream);

/* ---- a12: torch.optim.Adam step (trainer.py:11,54-56) fused, fp32 ----------
 * p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps), torch's non-capturable formula order. */
int idg_adam_step(float* d_p, const float* d_g, float* d_m, float* d_v, int64_t n, float lr, float beta1,
                  float beta2, float eps, int32_t step, void* stream);

/* Adam fused into the last backward layer: idg_adam_prepare computes the step scalars on the device (and
 * advances *d_step); idg_propagate_bwd_adam is idg_propagate_bwd_ex whose last product updates p/m/v in its
 * epilogue instead of writing the gradient; idg_spmm_layer_adam is that last product alone (multi-GPU path:
 * updated parameter rows inside the peer slab are stored to every peer as well). */
int idg_adam_prepare(int32_t* d_step, float* d_scalars, float lr, float beta1, float beta2, void* stream);
int idg_propagate_bwd_adam(const idg_graph* g, const float* d_G, const float* d_Gcl, int32_t d, int32_t K,
                           int include_layer0, int32_t cl_layer, float* d_work, const uint32_t* d_bitmap,
                           const idg_adam_args* adam, void* stream);
int idg_spmm_layer_adam(const idg_graph* g, const float* d_X, const float* d_addend, float acc_div, int32_t d,
                        const idg_adam_args* adam, void* stream);
/* SimGCL (models/SimGCL.py:62-66): the clean propagation and the two perturbed views of one step start from the same ego
 * table, so their first product A_hat . E0 is the same gather.  One launch: Y = A X (clean) and, per view v with noise_v != NULL,
 * Y_v = Y + sign(Y) * normalize(noise_v, dim=-1) * eps -- bit-identical to three separate idg_spmm_layer calls. */
int idg_spmm_layer_views(const idg_graph* g, const float* d_X, float* d_Y, const float* d_noise_a, float* d_Y_a,
                         const float* d_noise_b, float* d_Y_b, float eps, int32_t d, void* stream);
/* idg_spmm_layer_sparse_in on the rows flagged in d_rowmask only (the others are left untouched): for an output that is
 * pre-zeroed and can only be non-zero on the batch rows and their neighbours (the closure bitmap). */
int idg_spmm_layer_sparse_in_masked(const idg_graph* g, const float* d_X, float* d_Y, const float* d_addend, int32_t d,
                                    const uint32_t* d_bitmap, const uint32_t* d_rowmask, int skip_zero_rows, void* stream);
/* One step of the backward Horner chain on the handle's rows with a second addend: Y = A X + addend + scale2 * addend2
 * (XSimGCL.py:57-58,64-66: the gradient of the captured contrast layer joins the chain at that layer); d_bitmap != NULL
 * makes it the sparse-input product (X zero outside the flagged rows).  Used by the row-partitioned contrastive steps,
 * where the whole-graph idg_propagate_bwd_ex cannot be. */
int idg_spmm_layer_add2(const idg_graph* g, const float* d_X, float* d_Y, const float* d_addend, const float* d_addend2,
                        float scale2, int32_t d, const uint32_t* d_bitmap, int skip_zero_rows, void* stream);

/* Same update with the step counter on the device (*d_step = steps already taken; incremented by
 * the call): lets a captured CUDA graph of the whole train step be replayed unchanged. */
int idg_adam_step_dev(float* d_p, const float* d_g, float* d_m, float* d_v, int64_t n, float lr, float beta1,
                      float beta2, float eps, int32_t* d_step, void* stream);

/* ---- multi-GPU (SURVEY.md 8 e): row-partitioned nodes, one process per GPU --------------------
 * Each rank allocates one slab (idg_device_alloc), exports it (idg_ipc_get_handle, 64 bytes), opens the
 * peers' (idg_ipc_open) and registers the table (idg_peers_create; bases[rank] is ignored).  After
 * idg_graph_set_peers, every idg_spmm_layer* call whose d_Y lies inside the slab ALSO stores each finished
 * row at the same offset of every peer's slab over NVLink: the per-layer all-gather is fused into the SpMM
 * epilogue.  idg_peers_barrier is a device-side flag barrier over the slabs (d_state: 64 ints inside the
 * slab, zero-initialised, same offset on every rank); idg_peers_push copies a byte range of the local slab
 * to the same offset of all peers (parameter rows after the Adam step). */
int idg_device_alloc(int64_t bytes, void** out);
int idg_device_free(void* p);
int idg_ipc_get_handle(const void* d_ptr, void* handle64);
int idg_ipc_open(const void* handle64, void** out);
int idg_ipc_close(void* p);
int idg_peers_create(void* local_base, int64_t bytes, int32_t rank, int32_t world, void* const* bases, idg_peers** out);
void idg_peers_destroy(idg_peers* p);
/* optional: the NVSwitch multicast (NVLS) mapping of the same slab; row stores and pushes then go out once
 * as multimem.st and the switch replicates them into every GPU's copy (1/(G-1) of the NVLink egress) */
int idg_peers_set_multicast(idg_peers* p, void* mc_base);
int idg_graph_set_peers(idg_graph* g, const idg_peers* p);
int idg_peers_push(const idg_peers* p, const void* d_src, int64_t bytes, void* stream);
/* Same copy by a persistent grid of n_ctas CTAs: a few SMs stream a finished block of rows to the peers while the rest of the
 * GPU computes the next block (chunked exchange of idgrec/dist.py; replaces the in-epilogue peer stores where NVLink
 * back-pressure stalled the propagation kernel's gathers). */
int idg_peers_push_ctas(const idg_peers* p, const void* d_src, int64_t bytes, int32_t n_ctas, void* stream);
int idg_peers_barrier(const idg_peers* p, int32_t* d_state, void* stream);
/* The barrier's wait is bounded (default 20 s per barrier, idg_peers_set_timeout_ms): a peer that never arrives is
 * recorded in d_state[1] and the kernel returns instead of spinning; every later barrier on that slab returns at
 * once.  idg_peers_status SYNCHRONISES the stream and returns 0, or IDG_ERR_PEER_TIMEOUT + r when rank r was missing
 * (the results of the steps since the last healthy status are then invalid). */
#define IDG_ERR_PEER_TIMEOUT 100000
int idg_peers_set_timeout_ms(idg_peers* p, int64_t ms);
int idg_peers_status(const idg_peers* p, const int32_t* d_state, void* stream);

/* ---- a4: data_loader.py:108-127, exact replay on the HOST -------------------
 * h_cand: candidate stream = np.random.randint(0, I, size=n_cand) drawn from the
 * reference's RNG state.  Walks the train edges in file order, skipping candidates
 * that are positives of the current user (CSR of user_item_net, sorted).  Writes
 * h_neg[E]; *h_consumed = number of candidates used (caller re-winds the numpy
 * stream to exactly that many draws).  Returns -2 if the stream ran out. */
int idg_neg_sample_replay(const int64_t* h_train_user, int64_t E, const int32_t* h_pos_indptr,
                          const int32_t* h_pos_indices, const int64_t* h_cand, int64_t n_cand,
                          int64_t* h_neg, int64_t* h_consumed);

/* Resumable form: edges [e_begin, E) against ONE CHUNK of the candidate stream; *h_edges_done = first edge not yet
 * placed (== E when finished), *h_consumed = candidates of this chunk that were used.  A candidate rejected for the
 * edge at which the chunk ran out stays consumed, exactly as in the reference's while-loop. */
int idg_neg_sample_walk(const int64_t* h_train_user, int64_t e_begin, int64_t E, const int32_t* h_pos_indptr,
                        const int32_t* h_pos_indices, const int64_t* h_cand, int64_t n_cand, int64_t* h_neg,
                        int64_t* h_edges_done, int64_t* h_consumed);

/* a5: tools.shuffle (tools.py:35-52) applied on the device: d_out[3,n] = (a, b, c)[perm] (users, positives, negatives
 * of one epoch; only the permutation and the negatives cross PCIe, the train edges are resident). */
int idg_permute3(const int64_t* d_a, const int64_t* d_b, const int64_t* d_c, const int64_t* d_perm, int64_t n,
                 int64_t* d_out, void* stream);
/* The step's mini-batch fetched on the device from the epoch's (shuffled) sample arrays -- the first node of a captured train step,
 * so that an epoch is nothing but graph replays on the host (trainer.py:40-47 slices the three arrays per batch).  d_ptrs: device
 * int64[4] = base addresses of the user / positive / negative arrays and the step-counter value of the epoch's first batch; the
 * batch read is [(*d_step - start) * stride, + B); d_slab: [3, slab_stride] int64. */
int idg_batch_fetch(const void* d_ptrs, const int32_t* d_step, int32_t stride, int32_t B, int64_t* d_slab, int32_t slab_stride, void* stream);

/* nn.Tanh between the propagation layers of EGCF (models/EGCF.py:42,52-53,71): y = tanh(x); gx = gy (1 - y^2). */
int idg_tanh_fwd(const float* d_x, float* d_y, int64_t n, void* stream);
int idg_tanh_bwd(const float* d_y, const float* d_gy, float* d_gx, int64_t n, void* stream);

/* ---- a1: data_loader.py:48-70, the dataset text format ("user item item ..." per line) parsed on the HOST in one
 * pass.  Two-call protocol: pair_cap = line_cap = 0 counts (*n_pairs, *n_lines); the second call fills
 * h_user/h_item [n_pairs] (file order = inter_users/inter_items), h_line_user/h_line_len [n_lines] (unique_users and
 * the per-line item counts, 0 for a user with an empty line).  *max_user and *max_item follow data_loader.py:62-63
 * (lines with at least one item only; -1 when there is none).  -3: cannot open, -4: not an integer token. */
int idg_parse_ratings(const char* path, int64_t* h_user, int64_t* h_item, int64_t pair_cap, int64_t* n_pairs,
                      int64_t* h_line_user, int64_t* h_line_len, int64_t line_cap, int64_t* n_lines,
                      int64_t* max_user, int64_t* max_item);

/* ---- section 8 f (rank 4): batch x batch losses of the LightGCN-backbone models, forward + backward ----------
 * X, Y: dense [n,d] blocks (batch rows of the propagated tables, see idg_gather_rows); both are L2-normalised inside
 * (F.normalize, eps 1e-12) and the gradients are returned w.r.t. the UN-normalised X / Y for an upstream gradient of 1.
 *   kind 0  LightCCF neighbourhood-aggregation loss (models/LightCCF.py:81-94)            p0 = temperature
 *   kind 1  LightCSCF margin loss (models/LightCSCF.py:93-104)                            p0 = temperature, p1 = margin
 *   kind 2  SCCF "down" = log(sum_ij psi(S_ij) / p1) over ALL batch pairs (models/SCCF.py:72-79; summing over batch
 *           pairs equals the reference's count-weighted sum over unique users x unique items)  p0 = temperature,
 *           p1 = n_unique_users * n_unique_items
 *   kind 3  SCCF "-up" = -mean_i log psi(<a_i,b_i>) (models/SCCF.py:64-70)                 p0 = temperature
 *   kind 4  DirectAU alignment  mean_i |a_i - b_i|^2 (utility_function/losses.py:61-64)
 *   kind 5  DirectAU uniformity log mean_{i<j} exp(-2 |a_i - a_j|^2) (losses.py:67-69); Y, gY unused (may be NULL)
 * d_loss: one float.  d_gX/d_gY [n,d] may both be NULL (forward only).  n <= 8192.  Deterministic (fixed-order
 * reductions, no float atomics).  Workspace: idg_pair_loss_workspace_bytes(n, d) (two n x n fp32 matrices). */
int64_t idg_pair_loss_workspace_bytes(int32_t n, int32_t d);
int idg_pair_loss(int32_t kind, const float* d_X, const float* d_Y, int32_t n, int32_t d, float p0, float p1,
                  float* d_loss, float* d_gX, float* d_gY, void* d_ws, void* stream);

/* Same, for the fused train step: *d_loss += scale * loss and the gradients are multiplied by scale (ssl_lambda, gamma / 2,
 * ...); kind 2 may take its denominator from two device counts (n_unique_users, n_unique_items written by
 * idg_batch_rows_unique) instead of p1, so the whole step stays free of host syncs. */
int idg_pair_loss_ex(int32_t kind, const float* d_X, const float* d_Y, int32_t n, int32_t d, float p0, float p1, float scale,
                     const int32_t* d_cnt_a, const int32_t* d_cnt_b, float* d_loss, float* d_gX, float* d_gY, void* d_ws,
                     void* stream);

/* out[i,:] = T[idx[i],:]  (all_user_embeddings[user.long()], models/LightCCF.py:66) and its backward:
 * T[idx[i],:] += G[i,:] with duplicates summed in entry order by the first occurrence (index_put(accumulate) without
 * float atomics; bit-reproducible). */
int idg_gather_rows(const float* d_T, const int64_t* d_idx, int32_t n, int32_t d, float* d_out, void* stream);
int idg_scatter_add_rows(const float* d_G, const int64_t* d_idx, int32_t n, int32_t d, float* d_T, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IDGREC_H_ */
