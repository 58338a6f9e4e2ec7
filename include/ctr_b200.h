/*
 * ctr_b200.h — C ABI of libctr_b200.so: the B200 (sm_100a) CTR embedding + interaction hot path.
 *
 * The reference (shenweichen/DeepCTR-Torch) has NO native/FFI layer: its hot path is Python over
 * stock torch ops.  Each entry point below therefore cites the reference *Python* code whose
 * arithmetic it replaces (paths relative to the reference root, v0.2.9).  INTEGRATION.md shows the
 * ctypes stub a reference maintainer would add to bind them.
 *
 * Conventions
 *  - every tensor is caller-owned, device-resident, contiguous along its last axis, fp32 unless
 *    stated; `ld*` = leading dimension in ELEMENTS; `stream` is a cudaStream_t passed as void*.
 *  - the library never allocates, frees or synchronises; launches are asynchronous on `stream`.
 *  - return value: 0 = OK, <0 = argument error, >0 = cudaError_t.  `ctr_last_error()` returns a
 *    thread-local message for the last non-zero return on this thread.
 *  - ids travel inside X, one 4-byte cell per id.  `id_mode` says how a cell is read:
 *      CTR_IDS_F32     (0) fp32-encoded integers (reference models/basemodel.py:242,369), decoded by
 *                          truncation exactly like `.long()` — exact only below 2^24;
 *      CTR_IDS_I32BITS (1) the cell holds the int32 id itself (bit pattern of an int32 stored in the
 *                          fp32 matrix) — any id below 2^31, no float round trip.
 *    Out-of-range ids set bit 0 of `*err_flag` (the reference raises IndexError on CPU); the row is
 *    then read from id 0 so no memory is touched out of bounds.
 *  - "pointer arrays" (`const float* const*`) are DEVICE arrays of device pointers.
 */
#ifndef CTR_B200_H_
#define CTR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ---------------------------------------------------------------------------- */
int         ctr_version(void);            /* ABI version, currently 2 */
const char* ctr_last_error(void);
int         ctr_debug_set_buffer(void* dev_u64_buffer);   /* optional: per-stage clock64 timeline of the tensor-core GEMM (>= 256 u64), NULL = off */
int64_t     ctr_launch_count(void);       /* kernels launched by this library so far (process-wide) */
/* Tensor-core passes of the fp32 GEMMs: 3 (default, PARITY mode: 3xTF32, fp32-grade accuracy) or 1 (FAST mode,
 * NOT parity: single-pass TF32, ~1e-3 relative error like torch.backends.cuda.matmul.allow_tf32).  Process-wide,
 * read at launch time (a CUDA graph keeps the mode it was captured in).  Returns the previous value. */
int         ctr_set_gemm_passes(int passes);

/* activation codes shared by the dense ops (reference layers/activation.py:57-84) */
enum { CTR_ACT_LINEAR = 0, CTR_ACT_RELU = 1, CTR_ACT_SIGMOID = 2, CTR_ACT_TANH = 3 };
enum { CTR_IDS_F32 = 0, CTR_IDS_I32BITS = 1 };

/* ---- a4+a5+a6+a7: fused multi-slot gather + linear term + FM + dnn_input assembly --------
 * replaces: BaseModel.input_from_feature_columns  models/basemodel.py:354-380
 *           Linear.forward                        models/basemodel.py:63-92
 *           combined_dnn_input                    inputs.py:126-138
 *           FM.forward                            layers/interaction.py:26-34
 * blk[b, f*D + d] = emb_tables[f][id(b,f), d]        f < n_emb  (the [B,F,D] embedding block)
 * blk[b, n_emb*D + k] = X[b, dense_cols[k]]          k < n_dense
 * lin[b] = sum_f lin_tables[f][id(b,f)] + sum_k X[b, lin_dense_cols[k]] * lin_dense_w[k]
 * fm[b]  = 0.5 * sum_d ((sum_f E)^2 - sum_f E^2)     (only if fm != NULL)
 * Any of {blk, lin, fm} may be NULL (branch switched off).  D is the (uniform) embedding dim.
 * n_shards > 1: every table is row-sharded over the GPUs of the node (p2p.cu): the pointer arrays
 * hold n_slots * n_shards device pointers, entry [f*n_shards + s] = the rows {id : id % n_shards
 * == s} of field f stored at local index id / n_shards, on this GPU or a peer (NVLink P2P loads).
 */
int ctr_gather_fwd(const float* X, int64_t ldx, int64_t B,
                   int n_emb, int D, const float* const* emb_tables, const int32_t* emb_cols,
                   const int32_t* emb_vocab,
                   int n_lin, const float* const* lin_tables, const int32_t* lin_cols,
                   const int32_t* lin_vocab,
                   int n_dense, const int32_t* dense_cols,
                   int n_lin_dense, const int32_t* lin_dense_cols, const float* lin_dense_w,
                   float* blk, int64_t ld_blk, float* lin, float* fm,
                   int32_t* err_flag, int n_shards, int id_mode, void* stream);

/* FM on an already assembled block (used when pooled VarLen fields were added to it).
 * replaces FM.forward layers/interaction.py:26-34.  E = blk viewed as [B, F, D], row stride ld. */
int ctr_fm_fwd(const float* blk, int64_t ld, int64_t B, int F, int D, float* fm, void* stream);
/* d_blk[b,f,:] += g[b] * (S[b,:] - E[b,f,:])  (accumulates into d_blk) */
int ctr_fm_bwd(const float* blk, int64_t ld, int64_t B, int F, int D, const float* g,
               float* d_blk, int64_t ld_d, void* stream);

/* ---- backward of the gather ("embedding_dense_backward" x 52 in the reference) -----------
 * replaces: autograd of nn.Embedding (aten::embedding_dense_backward) reached from
 *           models/basemodel.py:261, plus FM backward and the dense part of Linear.
 * Row gradient of field f for sample b:
 *     r[b,f,:] = d_blk[b, f*D:(f+1)*D]  (+ g_fm[b] * (S[b,:] - E[b,f,:]) if g_fm != NULL)
 *     linear:    rl[b,f] = g_lin[b]
 *
 * (1) dense-compat mode: atomically accumulate rows into caller-zeroed dense [V,D] / [V,1]
 *     gradient tables (what the reference's sparse=False nn.Embedding produces).
 */
int ctr_scatter_bwd_dense(const float* X, int64_t ldx, int64_t B,
                          int n_emb, int D, float* const* emb_grads, const int32_t* emb_cols,
                          const int32_t* emb_vocab,
                          int n_lin, float* const* lin_grads, const int32_t* lin_cols,
                          const int32_t* lin_vocab,
                          const float* blk, int64_t ld_blk,
                          const float* d_blk, int64_t ld_dblk,
                          const float* g_fm, const float* g_lin, int id_mode, void* stream);

/* device pointer table for the calls above, written by a kernel from by-value arguments (so it can be
 * recorded into a CUDA graph; a host->device copy of a Python list cannot): dev_out[i] = ptrs[i], n <= 4096 */
int ctr_write_ptrs(const void* const* host_ptrs, int n, void** dev_out, void* stream);

/* (2) row-wise mode (B200-native): per id column a duplicate-free list of touched rows.
 *   ctr_unique_plan  : for each of n_cols id columns builds, with an open-addressing hash table
 *                      of H (power of two >= 2B) slots per column,
 *                        uniq[c*B + u]   = u-th distinct id of column c      (u < n_uniq[c])
 *                        inv [b*n_cols+c] = u such that uniq[c*B+u] == id(b,c)
 *                        cnt [c*B + u]   = multiplicity of that id in the batch
 *                      hash_keys/hash_vals: int32 [n_cols*H] scratch, n_uniq: int32 [n_cols];
 *                      all three are (re)initialised inside.  Entries u >= n_uniq[c] get
 *                      uniq = 0, cnt = 0.  col_count (may be NULL): int32 [n_cols], only the first
 *                      col_count[c] entries of column c exist (ragged receive lists, see
 *                      ctr_rowgrad_combine).
 *   ctr_scatter_bwd_rowwise: field f writes the [B, D] buffer at emb_rowgrad + f*emb_rowgrad_stride
 *                      (linear field f the [B] buffer at lin_rowgrad + f*lin_rowgrad_stride);
 *                      row u of field f receives the summed gradient of uniq id u of the plan
 *                      column emb_plan_col[f]; rows u >= n_uniq are zero-filled so that
 *                      (uniq, rowgrad) is always a valid padded sparse-COO pair of length B.
 */
int64_t ctr_unique_plan_hash_slots(int64_t B);
int ctr_unique_plan(const float* X, int64_t ldx, int64_t B, int n_cols, const int32_t* cols,
                    const int32_t* vocab, int32_t* hash_keys, int32_t* hash_vals, int64_t H,
                    int32_t* n_uniq, int32_t* uniq, int32_t* inv, int32_t* cnt,
                    int32_t* err_flag, int id_mode, const int32_t* col_count, void* stream);
int ctr_scatter_bwd_rowwise(int64_t B, int n_plan_cols, const int32_t* inv, const int32_t* cnt,
                            const int32_t* n_uniq,
                            int n_emb, int D, float* emb_rowgrad, int64_t emb_rowgrad_stride,
                            const int32_t* emb_plan_col,
                            int n_lin, float* lin_rowgrad, int64_t lin_rowgrad_stride,
                            const int32_t* lin_plan_col,
                            const float* blk, int64_t ld_blk,
                            const float* d_blk, int64_t ld_dblk,
                            const float* g_fm, const float* g_lin, void* stream);

/* ---- multi-GPU: row-sharded tables over NVLink peer memory (replaces torch.nn.DataParallel,
 * reference models/basemodel.py:206-209) ---------------------------------------------------
 * ctr_p2p_alloc/free : cudaMalloc'ed (zero-filled) buffer that can be exported to peers — the
 *                      only entry points that allocate; ctr_p2p_export writes the 64-byte CUDA
 *                      IPC handle, ctr_p2p_open maps a peer's buffer (peer access enabled lazily).
 * ctr_rowgrad_push   : after ctr_scatter_bwd_rowwise, appends every unique row id of plan column pc
 *                      (local row = id / G) to the receive list pc of its owner GPU id % G:
 *                      slot = remote atomicAdd on recv_count[owner][pc] (one claim per block and owner),
 *                      recv_ids[owner][pc*cap + slot] = local row, and EVERY field reading column pc
 *                      delivers its gradient into that slot: recv_emb_rows[owner][(f*cap+slot)*D ..]
 *                      for embedding field f (emb_plan_col[f] == pc), recv_lin_rows[owner][fl*cap+slot]
 *                      for linear field fl.  One list per id column means one owner-side plan per
 *                      column (ctr_unique_plan with col_count = recv_count) whatever the number of
 *                      tables fed by it.  recv_* are DEVICE arrays of n_shards peer pointers.  A full
 *                      list sets bit 1 of *err_flag.
 */
int ctr_p2p_alloc(int64_t bytes, void** ptr);
int ctr_p2p_free(void* ptr);
int ctr_p2p_export(void* ptr, unsigned char* handle64);
int ctr_p2p_open(const unsigned char* handle64, void** peer_ptr);
int ctr_p2p_close(void* peer_ptr);
/* device-side barrier of the n_shards ranks over peer memory (no NCCL): peer_flags[r] = rank r's int32
 * [n_shards] flag array (peer pointers, zero-initialised), epoch_ctr = this rank's private int32 counter.
 * Orders everything this stream wrote to peer memory before the call ahead of every peer's work after it.
 * A peer that never arrives sets bit 2 of *err_flag after a bounded spin. */
int ctr_p2p_barrier(int32_t* const* peer_flags, int32_t* epoch_ctr, int n_shards, int rank, int32_t* err_flag,
                    void* stream);
int ctr_rowgrad_push(int64_t B, int n_shards, int n_plan_cols, const int32_t* n_uniq, const int32_t* uniq,
                     int n_emb, int D, const float* emb_rowgrad, int64_t emb_rowgrad_stride,
                     const int32_t* emb_plan_col,
                     int n_lin, const float* lin_rowgrad, int64_t lin_rowgrad_stride,
                     const int32_t* lin_plan_col,
                     int32_t* const* recv_count, int32_t* const* recv_ids,
                     float* const* recv_emb_rows, float* const* recv_lin_rows,
                     int64_t cap, int32_t* err_flag, void* stream);

/* Forward exchange of cross-shard rows (csrc/p2p.cu; "owner computes" all-to-all over peer memory):
 *   ctr_shard_request : per (sample, distinct id column): ids owned by this rank get where = -1, the others
 *                       are appended as (column, local row) to the owner's inbox (peer pointers
 *                       inbox_req[o] = owner o's request list for THIS rank, capacity cap entries of 8
 *                       bytes) and where[b, c] = (owner << 26) | slot; finally the per-owner counts are
 *                       published to inbox_cnt[o][rank].  cnt_to: [n_shards] local scratch counters.
 *   -- all ranks synchronise (e.g. a tiny NCCL all-reduce) --
 *   ctr_shard_serve   : for every requester q, gathers the req_cnt[q] requested rows from this rank's
 *                       LOCAL tables (emb_of_col / lin_of_col: table of each id column or NULL) and streams
 *                       them into resp_emb[q] / resp_lin[q] (peer pointers: q's response buffers for THIS
 *                       owner, rows in request order).
 *   -- all ranks synchronise --
 *   ctr_gather_fwd_exchanged : ctr_gather_fwd whose rows of other shards come from the local response
 *                       buffers (resp_emb[o], resp_lin[o]: this rank's buffers for owner o) via where;
 *                       emb_plan_col / lin_plan_col: id-column index of each slot inside where. */
int ctr_shard_request(const float* X, int64_t ldx, int64_t B, int n_cols, const int32_t* cols,
                      const int32_t* vocab, int n_shards, int rank, int32_t* cnt_to,
                      int32_t* const* inbox_req, int32_t* const* inbox_cnt, int32_t* where,
                      int64_t cap, int32_t* err_flag, int id_mode, void* stream);
int ctr_shard_serve(int n_shards, int rank, int D, const int32_t* req_cnt, const int32_t* req, int64_t cap,
                    const float* const* emb_of_col, const float* const* lin_of_col,
                    float* const* resp_emb, float* const* resp_lin, void* stream);
int ctr_gather_fwd_exchanged(const float* X, int64_t ldx, int64_t B, int n_emb, int D,
                             const float* const* emb_tables, const int32_t* emb_cols,
                             const int32_t* emb_vocab, int n_lin, const float* const* lin_tables,
                             const int32_t* lin_cols, const int32_t* lin_vocab, int n_dense,
                             const int32_t* dense_cols, int n_lin_dense,
                             const int32_t* lin_dense_cols, const float* lin_dense_w, float* blk,
                             int64_t ld_blk, float* lin, float* fm, int32_t* err_flag,
                             int n_shards, int rank, const int32_t* where, int n_plan,
                             const int32_t* emb_plan_col, const int32_t* lin_plan_col,
                             const float* const* resp_emb, const float* const* resp_lin, int id_mode,
                             void* stream);

/* gradient of Linear's dense weight: dw[k] = sum_b g[b] * X[b, cols[k]]   (basemodel.py:88-90) */
int ctr_lin_dense_wgrad(const float* X, int64_t ldx, int64_t B, int n, const int32_t* cols,
                        const float* g, float* dw, void* stream);

/* ---- a8: DNN tower (reference layers/core.py:120-134: Linear(+bias) -> activation) -------
 * The weight is addressed as W(n,k) = W[n*swn + k*swk] so that both nn.Linear weights
 * ([N,K] row-major: swn = K, swk = 1) and the [K,N] factors of CrossNetMix (swn = 1, swk = N)
 * are consumed in place.  bias may be NULL.
 * fwd : Y[B,N] = act(X[B,K] @ W^T + bias[N])
 * bwd : dZ = dY (.) act'(Y)  applied on the fly;
 *       dX[B,K] (+)= dZ @ W           (dX may be NULL; accumulate_dx != 0 adds into dX)
 *       dW(n,k)  = sum_b dZ[b,n] X[b,k]  stored at dW[n*sdwn + k*sdwk] (one of them == 1),
 *       db[N]    = colsum(dZ)          (dW / db may be NULL)
 */
int ctr_dnn_layer_fwd(const float* X, int64_t ldx, const float* W, int64_t swn, int64_t swk,
                      const float* bias, float* Y, int64_t ldy, int64_t B, int K, int N, int act,
                      void* stream);
int ctr_dnn_layer_bwd(const float* X, int64_t ldx, const float* W, int64_t swn, int64_t swk,
                      const float* Y, int64_t ldy, const float* dY, int64_t lddy,
                      float* dX, int64_t lddx, int accumulate_dx,
                      float* dW, int64_t sdwn, int64_t sdwk, float* db,
                      int64_t B, int K, int N, int act, void* stream);

/* Backward of one layer INSIDE a tower (reference layers/core.py:120-134 stacks Linear -> act):
 * like ctr_dnn_layer_bwd, plus the two fusions that remove the activation-mask traffic between
 * consecutive layers:
 *   dy_is_dz != 0 : dY already IS dZ of this layer (the layer above produced it), Y is not read;
 *   dx_act        : activation of the layer BELOW (whose output is this layer's input X): the
 *                   input gradient is written as dX (.) dx_act'(X), i.e. as the dZ of that layer
 *                   (CTR_ACT_LINEAR = plain dX). */
int ctr_dnn_layer_bwd_chain(const float* X, int64_t ldx, const float* W, int64_t swn, int64_t swk,
                            const float* Y, int64_t ldy, const float* dY, int64_t lddy,
                            float* dX, int64_t lddx, float* dW, int64_t sdwn, int64_t sdwk, float* db,
                            int64_t B, int K, int N, int act, int dy_is_dz, int dx_act, void* stream);
/* 1 when ctr_dnn_layer_bwd_chain(..., dX = NULL, dW, db) with these operands runs on the engine that streams both
 * operands from the batch-major activations (no caller scratch from ctr_set_scratch is touched): such a call may be
 * issued on a second stream while the first one continues with the input gradient's consumers.  0 otherwise. */
int ctr_dnn_wgrad_is_scratch_free(const float* X, int64_t ldx, const float* Y, int64_t ldy, const float* dY,
                                  int64_t lddy, const float* dW, int64_t sdwn, int64_t sdwk, int64_t B, int K,
                                  int N, int act, int dy_is_dz);

/* Scratch for the tensor-core GEMM engine (csrc/gemm_pk.cu): every GEMM-shaped entry point
 * (ctr_dnn_layer_*, ctr_sgemm, ctr_cross_matrix_*, CrossNetMix projections) first re-tiles its two
 * operands into pre-split (hi, lo) TF32 tiles inside a CALLER-OWNED device buffer.  The library
 * never allocates: register one buffer per device with ctr_set_scratch (128-byte aligned; it must
 * stay alive and must not be shared by launches that may run concurrently on different streams);
 * ctr_gemm_scratch_bytes gives the requirement of one C[M,N] = A[M,K] B[N,K]^T.  Without enough
 * scratch the GEMMs run on the slower in-kernel-split engine (csrc/gemm_tc.cu). */
int64_t ctr_gemm_scratch_bytes(int64_t M, int64_t N, int64_t K);
int ctr_set_scratch(void* ptr, int64_t bytes);

/* generic fp32 GEMM used by the layer kernels and exposed for composition:
 *   C[m,n] (+)= sum_k A[m*sam + k*sak] * Bm[n*sbn + k*sbk]      (any strides, in elements) */
int ctr_sgemm(int64_t M, int64_t N, int64_t K,
              const float* A, int64_t sam, int64_t sak,
              const float* Bm, int64_t sbn, int64_t sbk,
              float* C, int64_t ldc, int accumulate, void* stream);

/* row-dot head: out[b] (+)= sum_n H[b,n] * w[n]   (the bias-free dnn_linear / cin_linear,
 * reference models/deepfm.py:59-60,81; xdeepfm.py:73,88; dcn.py:64-65,86)
 * bwd: dH[b,n] (+)= g[b]*w[n];  dw[n] = sum_b g[b]*H[b,n] */
int ctr_rowdot_fwd(const float* H, int64_t ldh, const float* w, int64_t B, int N,
                   float* out, int accumulate, void* stream);
int ctr_rowdot_bwd(const float* H, int64_t ldh, const float* w, const float* g, int64_t B, int N,
                   float* dH, int64_t lddh, int accumulate_dh, float* dw, void* stream);

/* ---- a9: PredictionLayer (reference layers/core.py:154-160) ------------------------------
 * y[b] = sigmoid(sum_t terms[t][b] + bias)  (task_binary) or the plain sum (regression);
 * logit[b] (optional) receives the pre-sigmoid value.  terms: HOST array of device pointers.
 * bwd: dlogit[b] = dy[b] * y(1-y) (binary) or dy[b];  dbias[0] = sum_b dlogit[b] */
int ctr_predict_fwd(const float* const* terms, int n_terms, const float* bias, int64_t B,
                    int task_binary, float* logit, float* y, void* stream);
int ctr_predict_bwd(const float* y, const float* dy, int64_t B, int task_binary,
                    float* dlogit, float* dbias, void* stream);

/* ---- a10: CIN (reference layers/interaction.py:207-248) ----------------------------------
 * one layer:  Z[b,n,d] = sum_{h,m} W[n, h*M+m] * Xp[b,h,d] * X0[b,m,d] + bias[n];  Y = act(Z)
 *   Xp: [B,H,D] with batch stride sxp (elements), X0: [B,M,D] with batch stride sx0;
 *   Y : [B,N,D] contiguous (kept for the backward).  Channels [0, n_hidden) are the next
 *       layer's Xp (n_hidden is only needed by the backward); channels [direct_start, N) are
 *       "direct connect":  out[b, n - direct_start] = sum_d Y[b,n,d], written at out + b*ld_out.
 *       split_half, not last layer: n_hidden = direct_start = N/2; last layer: n_hidden = 0,
 *       direct_start = 0; split_half=False: n_hidden = N, direct_start = 0.
 * bwd: dZ[b,n,d] = ([n < n_hidden] dYh[b,n,d] + [n >= direct_start] dout[b, n-direct_start])
 *                  * act'(Y[b,n,d])            (materialised in the caller's dZ buffer [B,N,D]);
 *      dW[n,hm] = sum_{b,d} dZ * P ; dbias[n] = sum_{b,d} dZ ;
 *      dXp[b,h,d] = sum_{n,m} W dZ X0   (written, batch stride sdxp; NULL when Xp == X0: the
 *                                        first layer, whose Xp gradient is folded into dX0)
 *      dX0[b,m,d] += sum_{n,h} W dZ Xp  (accumulated, batch stride sdx0)
 */
int ctr_cin_layer_fwd(const float* Xp, int64_t sxp, int H, const float* X0, int64_t sx0, int M,
                      int D, const float* W, const float* bias, int N, int direct_start, int act,
                      float* Y, float* out, int64_t ld_out, int64_t B, void* stream);
int ctr_cin_layer_bwd(const float* Xp, int64_t sxp, int H, const float* X0, int64_t sx0, int M,
                      int D, const float* W, int N, int n_hidden, int direct_start, int act,
                      const float* Y, const float* dYh, int64_t sdyh,
                      const float* dout, int64_t ld_dout,
                      float* dZ, float* dW, float* dbias,
                      float* dXp, int64_t sdxp, float* dX0, int64_t sdx0,
                      int64_t B, void* stream);

/* ---- a11: CrossNet (reference layers/interaction.py:438-453) ------------------------------
 * vector: x_{l+1} = x0 * (x_l . w_l) + b_l + x_l, all L layers in one kernel.
 *   kernels/bias: [L, n] (the reference's [L,n,1] squeezed).  s[b,l] = x_l . w_l is saved.
 * matrix: x_{l+1} = x0 (.) (x_l W_l^T + b_l) + x_l ; one call per layer; U = x_l W_l^T + b_l saved.
 */
int ctr_cross_vector_fwd(const float* x0, int64_t ldx, const float* kernels, const float* bias,
                         int L, int n, float* out, int64_t ldo, float* s, int64_t B, void* stream);
int ctr_cross_vector_bwd(const float* x0, int64_t ldx, const float* kernels, const float* bias,
                         int L, int n, const float* s, const float* dout, int64_t lddo,
                         float* dx0, int64_t lddx, int accumulate_dx,
                         float* dkernels, float* dbias, int64_t B, void* stream);
int ctr_cross_matrix_layer_fwd(const float* x0, int64_t ldx0, const float* xl, int64_t ldxl,
                               const float* W, const float* bias, int n,
                               float* U, int64_t ldu, float* xnext, int64_t ldn,
                               int64_t B, void* stream);
/* g = d x_{l+1} [B,n] (read), U saved.  Produces dU = g (.) x0 (in dU buffer), dx0 += g (.) U,
 * dW = dU^T x_l, db = colsum(dU), gprev = g + dU W  (gprev may alias nothing; written) */
int ctr_cross_matrix_layer_bwd(const float* x0, int64_t ldx0, const float* xl, int64_t ldxl,
                               const float* W, const float* U, int64_t ldu,
                               const float* g, int64_t ldg, int n,
                               float* dU, int64_t lddu, float* dx0, int64_t lddx0,
                               float* dW, float* db, float* gprev, int64_t ldgp,
                               int64_t B, void* stream);

/* ---- a12: CrossNetMix helpers (reference layers/interaction.py:499-534) ------------------
 * The low-rank expert projections are ctr_sgemm / ctr_dnn_layer calls; these two fuse the rest:
 * mix fwd : x_{l+1}[b,:] = sum_e softmax_e(gate[b,:])_e * x0[b,:] (.) (uv_e[b,:] + bias) + x_l[b,:]
 *           uv: [E][B,n] stacked (expert stride = B*ldu), gate: [B,E]
 * mix bwd : given g = d x_{l+1}: d uv_e = p_e * g (.) x0 ; dgate (through softmax) ;
 *           dx0 += sum_e p_e g (.) (uv_e+bias) ; dbias = colsum(sum_e p_e g (.) x0) ; dxl = g
 */
int ctr_cross_mix_fwd(const float* x0, int64_t ldx0, const float* xl, int64_t ldxl,
                      const float* uv, int64_t ldu, const float* gate, const float* bias,
                      int E, int n, float* xnext, int64_t ldn, int64_t B, void* stream);
int ctr_cross_mix_bwd(const float* x0, int64_t ldx0, const float* uv, int64_t ldu,
                      const float* gate, const float* bias, const float* g, int64_t ldg,
                      int E, int n, float* duv, float* dgate, float* dx0, int64_t lddx0,
                      float* dbias, int64_t B, void* stream);

/* ---- a13: SENET (reference layers/interaction.py:93-101) ---------------------------------
 * Z = mean_d E; A = relu(W2 relu(W1 Z)); V = E * A[:,:,None].  E,V: [B,F,D] batch stride se/sv.
 * W1: [R,F], W2: [F,R]. */
int ctr_senet_fwd(const float* E, int64_t se, int F, int D, const float* W1, const float* W2, int R,
                  float* V, int64_t sv, int64_t B, void* stream);
int ctr_senet_bwd(const float* E, int64_t se, int F, int D, const float* W1, const float* W2, int R,
                  const float* dV, int64_t sdv, float* dE, int64_t sde, int accumulate_de,
                  float* dW1, float* dW2, int64_t B, void* stream);

/* ---- a14: BilinearInteraction (reference layers/interaction.py:140-156) ------------------
 * out[b,p,:] = (E[b,i_p,:] @ W_{w(p)}^T) (.) E[b,j_p,:]  for pairs p in combinations order.
 * W: [n_w, D, D]; wsel = 0: all (one W), 1: each (W_i), 2: interaction (W_p).
 * out: [B, P, D] written at out + b*so (batch stride so, lets two calls fill one DNN input). */
int ctr_bilinear_fwd(const float* E, int64_t se, int F, int D, const float* W, int wsel,
                     float* out, int64_t so, int64_t B, void* stream);
int ctr_bilinear_bwd(const float* E, int64_t se, int F, int D, const float* W, int wsel,
                     const float* dout, int64_t sdo, float* dE, int64_t sde, float* dW,
                     int64_t B, void* stream);

/* ---- a16: whole-table L2 term (reference models/basemodel.py:412-428) --------------------
 * out[0] += scale * sum(w^2) over n elements */
int ctr_sumsq_acc(const float* w, int64_t n, float scale, float* out, void* stream);

/* ---- a16: the loss of fit(): F.binary_cross_entropy(y_pred, y, reduction='sum') (reference
 * models/basemodel.py:254) with ATen's clamps (log >= -100; gradient denominator >= 1e-12).
 * fwd: out[0] = sum_b -(y log p + (1-y) log(1-p));  bwd: d_pred[b] = g[0] * (p - y) / max(p (1-p), 1e-12) */
int ctr_bce_sum_fwd(const float* y_pred, const float* y, int64_t B, float* out, void* stream);
int ctr_bce_sum_bwd(const float* y_pred, const float* y, const float* g, int64_t B, float* d_pred, void* stream);

/* PReLU with one learnable slope (reference layers/activation.py:61-62, `dnn_activation='prelu'`): elementwise over n
 * values; bwd zero-fills and accumulates dalpha[0] = sum dy * min(z, 0) */
int ctr_prelu_fwd(const float* z, const float* alpha, int64_t n, float* y, void* stream);
int ctr_prelu_bwd(const float* z, const float* alpha, const float* dy, int64_t n, float* dz, float* dalpha,
                  void* stream);

/* ---- VarLenSparseFeat pooled lookup (reference inputs.py:141-155, layers/sequence.py:49-77)
 * ids X[b, col .. col+T); mask = (id != 0) when len_col < 0, else t < (int)X[b,len_col].
 * mode 0 sum, 1 mean (divide by count + 1e-8), 2 max (masked positions contribute row - 1e9).
 * dst[b*ld + d] = pooled (dst points at the field's slot inside blk).
 * bwd (dense-compat): atomically adds into the caller-zeroed dense gradient table. */
int ctr_varlen_pool_fwd(const float* X, int64_t ldx, int64_t B, int col, int T, int len_col,
                        const float* table, int vocab, int D, int mode,
                        float* dst, int64_t ld, int32_t* err_flag, int id_mode, void* stream);
int ctr_varlen_pool_bwd(const float* X, int64_t ldx, int64_t B, int col, int T, int len_col,
                        const float* table, int vocab, int D, int mode,
                        const float* ddst, int64_t ld, float* dtable, int id_mode, void* stream);

/* ---- f2: fused row-wise optimizer on the row-gradient stream ------------------------------
 * replaces: optim.step() + the embedding part of get_regularization_loss().backward()
 *           (reference models/basemodel.py:262, 412-428, 447-461) for the tables, consuming the
 *           (uniq, rowgrad) pairs of ctr_scatter_bwd_rowwise in place — no dense [V,D] gradient, no
 *           sparse-COO hop.  Only rows that occur in the batch are touched ("lazy" semantics: identical
 *           to the dense torch optimizer for sgd / adagrad at l2 = 0, to torch.optim.SparseAdam-style
 *           row-wise moments for adam / rmsprop; L2 is applied to the touched rows as g += l2x2 * w).
 *   hp (device, 8 floats, written by ctr_rowopt_tick so that a captured CUDA graph sees fresh values):
 *      [0] step  [1] lr  [2] beta1 (rmsprop: alpha)  [3] beta2  [4] eps  [5] 1-beta1^step  [6] 1-beta2^step
 *   ctr_rowopt_tick : step += 1 and refresh the bias corrections (one thread).
 *   ctr_rowopt_step : for field f < n_fields and u < n_uniq[plan_col[f]]:
 *        id = uniq[plan_col[f]*cap + u],  g = rowgrad[f*rowgrad_stride + u*D ..] + l2x2 * w
 *        kind 0 sgd     : w -= lr * g
 *        kind 1 adagrad : s1 += g^2;  w -= lr * g / (sqrt(s1) + eps)
 *        kind 2 adam    : s1 = b1*s1 + (1-b1) g;  s2 = b2*s2 + (1-b2) g^2;
 *                         w -= (lr / bc1) * s1 / (sqrt(s2) / sqrt(bc2) + eps)       (torch.optim.Adam)
 *        kind 3 rmsprop : s1 = a*s1 + (1-a) g^2;  w -= lr * g / (sqrt(s1) + eps)
 *      tables / state1 / state2: device arrays of n_fields pointers to [V_f, D] tensors (state unused by
 *      the kind may be NULL).  row_div > 1: the tables are row shards, row = id (ids are already local).
 *   ctr_rowgrad_combine : owner side of the sharded backward: sums the received
 *        (row, gradient) pairs of field f that share a row, using a unique plan built over the receive
 *        lists (ctr_unique_plan with col_count; one list per id column pc, see ctr_rowgrad_push):
 *        out[f][inv[b, pc]] += recv[f][b] for b < count[pc], pc = plan_col[f]; count: int32 [n_plan_cols].
 *        The rows u < n_uniq[pc] of out are zeroed inside (n_uniq: the plan's distinct-row counts).
 */
enum { CTR_OPT_SGD = 0, CTR_OPT_ADAGRAD = 1, CTR_OPT_ADAM = 2, CTR_OPT_RMSPROP = 3 };
int ctr_rowopt_tick(float* hp, void* stream);
int ctr_rowopt_step(int kind, int64_t cap, const int32_t* n_uniq, const int32_t* uniq,
                    int n_fields, int D, const float* rowgrad, int64_t rowgrad_stride,
                    const int32_t* plan_col, float* const* tables, float* const* state1,
                    float* const* state2, const float* hp, float l2x2, void* stream);
int ctr_rowgrad_combine(int64_t cap, int n_fields, int D, const int32_t* count, const int32_t* n_uniq,
                        const int32_t* inv, int n_plan_cols, const int32_t* plan_col, const float* recv,
                        int64_t recv_stride, float* out, int64_t out_stride, void* stream);

/* ---- f4: interaction ops of the adjacent models (WDL / NFM / AFM / IFM / DIFM reuse everything above)
 * ctr_bipool_*  : BiInteractionPooling, reference layers/interaction.py:54-61:
 *                 out[b,d] = 0.5 * ((sum_f E[b,f,d])^2 - sum_f E[b,f,d]^2);  bwd writes dE[b,f,d] = g[b,d] * (S[b,d] - E[b,f,d])
 * ctr_refine_*  : input-aware re-weighting, reference models/ifm.py:83-88, models/difm.py:104-109 and the
 *                 sparse_feat_refine_weight branch of Linear.forward (models/basemodel.py:82-84):
 *                 m = softmax ? F * softmax_f(P) : P;  Er[b,f,:] = m[b,f] * E[b,f,:];  lin[b] = sum_f m[b,f] * L[b,f]
 *                 (L = per-field linear weights of the sample, may be NULL; lin may be NULL).
 *                 bwd: dP (through the softmax when set), dE = dEr * m, dL = dlin * m.
 * ctr_afm_*     : AFMLayer, reference layers/interaction.py:250-331 (pairs in itertools.combinations order):
 *                 out[b,:] = sum_pairs softmax_pairs(h . relu(W^T (e_i*e_j) + b)) (e_i*e_j)   ([B,D]: the attention
 *                 output; dropout and the projection p are applied by the caller);  W [D,A], b [A], h [A].
 *                 bwd (g = d out [B,D]) zero-fills and accumulates dW/db/dh, writes dE.
 * ctr_fieldattn_*: the attention core of InteractingLayer (reference layers/interaction.py:372-392, used by DIFM):
 *                 Y[b,i,:] = relu(concat_heads(softmax_j(Q_i . K_j * scale) V_j) + R[b,i,:]);  Q, K, V, R, Y: [B, F, D]
 *                 contiguous (the four projections are GEMMs on the [B*F, D] view); bwd writes dQ, dK, dV, dR. */
int ctr_fieldattn_fwd(const float* Q, const float* K, const float* V, const float* R, int F, int D, int H,
                      float scale, float* Y, int64_t B, void* stream);
int ctr_fieldattn_bwd(const float* Q, const float* K, const float* V, const float* Y, const float* dY, int F, int D,
                      int H, float scale, float* dQ, float* dK, float* dV, float* dR, int64_t B, void* stream);
int ctr_bipool_fwd(const float* E, int64_t se, int F, int D, float* out, int64_t so, int64_t B, void* stream);
int ctr_bipool_bwd(const float* E, int64_t se, int F, int D, const float* g, int64_t sg, float* dE, int64_t sde,
                   int64_t B, void* stream);
int ctr_refine_fwd(const float* P, const float* E, int64_t se, const float* L, int F, int D, int softmax, float* m,
                   float* Er, int64_t ser, float* lin, int64_t B, void* stream);
int ctr_refine_bwd(const float* m, const float* E, int64_t se, const float* L, int F, int D, int softmax,
                   const float* dEr, int64_t sder, const float* dlin, float* dP, float* dE, int64_t sde,
                   float* dL, int64_t B, void* stream);
int ctr_afm_fwd(const float* E, int64_t se, int F, int D, int A, const float* W, const float* b, const float* h,
                float* out, int64_t B, void* stream);
int ctr_afm_bwd(const float* E, int64_t se, int F, int D, int A, const float* W, const float* b, const float* h,
                const float* g, float* dE, int64_t sde, float* dW, float* db, float* dh, int64_t B, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTR_B200_H_ */
