/*
 * ncf_b200.h - C ABI of libncf_b200.so: the B200 (sm_100a) implementation of the AdvancedNCF
 * training + scoring hot path of ethanshenley/Neural-Collaborative-Filtering-Demo.
 *
 * The reference has no FFI: the path sits behind a PyTorch nn.Module (reference
 * src/model/architecture.py:121 `AdvancedNCF`).  This header is the boundary a maintainer binds
 * with ctypes from that module; every entry point cites the reference
 * code it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory unless named host_*;
 *   - the library never allocates or frees device memory: scratch is a caller-provided
 *     workspace sized by the *_workspace_bytes queries;
 *   - every launch goes on the `stream` argument (a cudaStream_t passed as void*); no hidden
 *     synchronisation, so calls are CUDA-graph capturable;
 *   - return value: 0 on success, negative ncf_status on error (message: ncf_last_error());
 *     no C++ exception crosses the boundary;
 *   - model geometry is the reference's configuration (config/config.yaml:53-69):
 *     embedding dim 64 (both towers), MLP 96->256->128->64, 4 heads x 16, temporal dim 32.
 *   - ids are int64 (the KeyedJaggedTensor `values` dtype, data_prep.py:286-298); a sample row
 *     n uses user_ids[n] and item_ids[n]; rows b*S .. b*S+S-1 form one interaction group.
 */
#ifndef NCF_B200_H
#define NCF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NCF_ABI_VERSION 2
#if defined(__GNUC__)
#define NCF_API __attribute__((visibility("default")))
#else
#define NCF_API
#endif
#define NCF_D 64          /* embedding dim of all four tables */
#define NCF_H1 256
#define NCF_H2 128
#define NCF_H3 64
#define NCF_HEADS 4
#define NCF_TDIM 32       /* temporal_dim */
#define NCF_MAX_S 8       /* 1 + negative_samples supported by the attention kernels */

typedef enum ncf_status {
  NCF_OK = 0,
  NCF_ERR_ARG = -1,        /* bad argument (null pointer, size, alignment, unsupported S) */
  NCF_ERR_WORKSPACE = -2,  /* workspace too small */
  NCF_ERR_CUDA = -3,       /* a CUDA runtime call or launch failed */
  NCF_ERR_UNSUPPORTED = -4
} ncf_status;

/* Table order used everywhere: the four EmbeddingBagCollection tables of
 * architecture.py:153-190 (SUM-pooled bags of length 1 == row gather). */
enum { NCF_T_USER_MF = 0, NCF_T_ITEM_MF = 1, NCF_T_USER_MLP = 2, NCF_T_ITEM_MLP = 3 };

typedef struct ncf_tables {
  float* w[4];            /* fp32 [rows, 64] row-major, 16-byte aligned */
  float* m[4];            /* Adam first moment  (may be NULL when no table update is requested) */
  float* v[4];            /* Adam second moment */
  float* g[4];            /* dense gradient [rows,64], only for NCF_EMB_MATERIALIZE (caller zeroes) */
  uint8_t* touched[2];    /* [rows_user], [rows_item] scratch flags, zero between steps */
  int64_t rows_user;
  int64_t rows_item;
  int32_t* status;        /* optional [NCF_STATUS_WORDS] sticky error words (see ncf_check_ids); device memory or pinned
                             host memory mapped into the device address space (then the host reads them without a copy) */
} ncf_tables;

/* Id validation.  nn.EmbeddingBag raises on an id outside [0, rows) (reference architecture.py:286-287 through torchrec);
 * a kernel cannot raise, so every kernel of this library that indexes a table with a caller-supplied id CLAMPS it into
 * range (no out-of-bounds access, no cross-table corruption of the sorted-id backward) and the kernels that see the ids
 * first (K1, ncf_gather_ln, the scoring kernels, ncf_check_ids) store 1 into the status word of the offending kind.  The
 * words are never cleared by the library: the host side clears them and raises (IndexError in the Python module) at its
 * next synchronisation point. */
enum { NCF_STATUS_BAD_USER_ID = 0, NCF_STATUS_BAD_ITEM_ID = 1, NCF_STATUS_BAD_HOUR = 2, NCF_STATUS_WORDS = 4 };

/* Offsets (in floats) of the 30 dense tensors `forward` uses inside ONE flat fp32 buffer
 * (83,909 values padded to 16-byte boundaries).  Names are the reference state_dict keys
 * (SURVEY Appendix C).  ncf_dense_offset(id) returns the offset, ncf_dense_numel() the length. */
typedef enum ncf_dense_id {
  NCF_P_MF_NORM_W = 0, NCF_P_MF_NORM_B, NCF_P_MLP_NORM_W, NCF_P_MLP_NORM_B,
  NCF_P_Q_W, NCF_P_K_W, NCF_P_V_W, NCF_P_O_W,   /* k_proj and v_proj adjacent: one [128,64] GEMM */
  NCF_P_Q_B, NCF_P_K_B, NCF_P_V_B, NCF_P_O_B,
  NCF_P_MLP0_W, NCF_P_MLP0_B, NCF_P_LN0_W, NCF_P_LN0_B,
  NCF_P_MLP1_W, NCF_P_MLP1_B, NCF_P_LN1_W, NCF_P_LN1_B,
  NCF_P_MLP2_W, NCF_P_MLP2_B, NCF_P_LN2_W, NCF_P_LN2_B,
  NCF_P_MF_OUT_W, NCF_P_MF_OUT_B, NCF_P_MLP_OUT_W, NCF_P_MLP_OUT_B,
  NCF_P_FINAL_W, NCF_P_FINAL_B,
  NCF_P_COUNT
} ncf_dense_id;

typedef enum ncf_precision {
  NCF_FP32 = 0,      /* CUDA-core fp32 towers: <=1e-5 relative vs the reference (BASELINE config 1) */
  NCF_BF16_TC = 1    /* tcgen05 bf16 towers, fp32 accumulate (BASELINE config 2) */
} ncf_precision;

typedef struct ncf_run_cfg {
  int32_t S;              /* rows per interaction group: 1+negative_samples in train, 1 in eval
                             (architecture.py:275) */
  int32_t training;       /* nn.Module.training */
  int32_t precision;      /* ncf_precision */
  float dropout_p;        /* architecture.py:51, 238 (train only) */
  uint64_t seed;          /* Philox key of the in-kernel dropout */
  uint64_t step;          /* Philox stream offset: one value per forward call */
} ncf_run_cfg;

typedef enum ncf_emb_mode {
  NCF_EMB_NONE = 0,         /* no table gradient (tables frozen) */
  NCF_EMB_MATERIALIZE = 1,  /* write the dense table gradient into tables->g (reference layout:
                               aten::_embedding_bag_dense_backward, trainer.py:276) */
  NCF_EMB_ADAM_SPARSE = 2,  /* fused scatter+Adam on touched rows only */
  NCF_EMB_ADAM_DENSE_EQUIV = 3 /* touched rows as above, then every untouched row gets the
                               reference's g = wd*w Adam step (trainer.py:71-75, 285) */
} ncf_emb_mode;

typedef struct ncf_adam_cfg {
  float lr, beta1, beta2, eps, weight_decay;
  int32_t step;             /* 1-based optimizer step (bias correction) */
  int32_t emb_mode;         /* ncf_emb_mode */
} ncf_adam_cfg;

/* ---- library ------------------------------------------------------------------------- */
NCF_API int ncf_version(void);
NCF_API const char* ncf_last_error(void);
NCF_API int64_t ncf_launch_count(void);   /* kernels this library has launched so far (process-wide) */
/* Every call enqueues on its `stream` argument only and never synchronises.  One opt-in exception: with an
 * auxiliary stream set here, ncf_train_step forks the id sort of the embedding backward and, in
 * NCF_EMB_ADAM_DENSE_EQUIV mode, the sweep of the rows the batch does not name (both depend on the ids alone and
 * touch nothing the rest of the step reads or writes) onto it, ordered with events against the stream argument;
 * the fork is joined again before ncf_train_step's embedding backward, which then runs its item side on the
 * auxiliary stream next to the user side (so does a direct ncf_emb_bwd_adam_both call); everything is joined
 * back into the stream argument before the call's last kernel.  With the auxiliary stream set, ncf_backward /
 * ncf_train_step / ncf_shard_backward* (NCF_BF16_TC, S = 5, >= 128 rows per SM) also run the MLP weight-gradient
 * kernel on a second, library-owned stream (created with the first use, per device) over 28 SMs next to the
 * attention backward, which leaves those SMs free; joined like the rest (NCF_WGRAD_SMS=0 in the environment: off).
 * The caller keeps the stream alive while it is set.  NULL (default) switches all of that off. */
NCF_API int ncf_set_aux_stream(void* stream);
/* SMs the persistent tower kernels (attention / MLP forward and backward: CTAs that own an SM's registers and shared
 * memory) leave free, process-wide, 0 by default.  The sharded step sets 1: its small collectives (a communication
 * kernel on another stream) then start at once instead of waiting for the running tower kernel to end - nothing can
 * co-reside with those CTAs.  At 65,536 interactions per step the tile counts per CTA do not change (2,560 MLP tiles
 * over 147 CTAs = 18 rounds, as over 148), so the towers lose nothing. */
NCF_API int ncf_set_sm_reserve(int32_t sms);
/* Early loss read-back for ncf_train_step (opt-in, per device like the auxiliary stream): the loss of a step is final after
 * the forward, a third of the way into the step.  With a pinned host float and an event set here, ncf_train_step copies
 * the loss to the host right after the loss kernel and records the event behind the copy, so a training loop that reads
 * its loss every step (trainer.py:289 `loss.item()`) waits for THAT event instead of draining the whole step, and
 * enqueues the next step while the backward still runs.  NULL, NULL switches it off. */
NCF_API int ncf_set_loss_readback(float* host_loss_pinned, void* cuda_event);
NCF_API int64_t ncf_dense_numel(void);
NCF_API int64_t ncf_dense_offset(int32_t dense_id);
NCF_API int64_t ncf_dense_size(int32_t dense_id);

/* Stand-alone validator (callers that hand ids to entry points without an ncf_tables argument, e.g. the sharded
 * requester): status[NCF_STATUS_BAD_USER_ID] = 1 if any user id is outside [0, rows_user), likewise item ids / rows_item
 * and hour (optional) outside [0, 24).  item_ids may be NULL. */
NCF_API int ncf_check_ids(const int64_t* user_ids, const int64_t* item_ids, int64_t N, int64_t rows_user, int64_t rows_item,
                  const int64_t* hour, int32_t* status, void* stream);

/* ---- forward ---------------------------------------------------------------------------
 * AdvancedNCF.forward (architecture.py:258-381) and forward_simple (:409-485).
 * hour == NULL: the temporal columns are zeros (:329-340).  hour != NULL: forward_simple's
 * hour path; tmod [24,64] and tail1 [24,256] come from ncf_temporal_tables.
 * out: probabilities [N].  The workspace keeps the activations ncf_backward needs. */
NCF_API int64_t ncf_workspace_bytes(int64_t N, const ncf_run_cfg* cfg);
NCF_API int ncf_forward(const ncf_run_cfg* cfg, const ncf_tables* tables, const float* dense,
                const int64_t* user_ids, const int64_t* item_ids, int64_t N,
                const int64_t* hour, const float* tmod, const float* tail1,
                float* out, void* workspace, int64_t workspace_bytes, void* stream);

/* loss.backward() of trainer.py:276 for `out` produced by ncf_forward with the same workspace:
 * grad_out = dL/d out [N].  Accumulates (+=) into dense_grad (flat layout above) and handles the
 * four tables according to adam->emb_mode (the fused sorted-id scatter + Adam, or the dense
 * gradient the reference materialises). */
NCF_API int ncf_backward(const ncf_run_cfg* cfg, const ncf_adam_cfg* adam, const ncf_tables* tables,
                 const float* dense, float* dense_grad,
                 const int64_t* user_ids, const int64_t* item_ids, int64_t N,
                 const float* grad_out, void* workspace, int64_t workspace_bytes, void* stream);

/* The four stages of the dense towers on their own, operating IN PLACE on the workspace of a
 * training- or eval-mode ncf_forward (same N, cfg, workspace): what ncf_forward / ncf_backward run
 * after K1 / before K6.  They exist for profiling and unit tests (bench.py times each against its
 * roofline); a caller that wants the model output uses ncf_forward.
 *   ncf_attn_fwd  MultiHeadAttention block (architecture.py:18-57, 315-326): xu, xp -> a
 *   ncf_mlp_fwd   self.mlp + mlp_output + final (architecture.py:230-252, 329-354): a, mf_pred -> out
 *   ncf_mlp_bwd   backward of the head and the MLP: grad_out -> d mf_pred, da (+ parameter gradients)
 *   ncf_attn_bwd  backward of the attention block: da -> dxu, dxp (+ parameter gradients) */
NCF_API int ncf_attn_fwd(const ncf_run_cfg* cfg, const float* dense, int64_t N,
                 void* workspace, int64_t workspace_bytes, void* stream);
NCF_API int ncf_mlp_fwd(const ncf_run_cfg* cfg, const float* dense, int64_t N, float* out,
                void* workspace, int64_t workspace_bytes, void* stream);
NCF_API int ncf_mlp_bwd(const ncf_run_cfg* cfg, const float* dense, float* dense_grad, int64_t N,
                const float* grad_out, void* workspace, int64_t workspace_bytes, void* stream);
NCF_API int ncf_attn_bwd(const ncf_run_cfg* cfg, const float* dense, float* dense_grad, int64_t N,
                 void* workspace, int64_t workspace_bytes, void* stream);

/* nn.BCELoss() mean reduction with torch's log clamp at -100 (trainer.py:78, 271) and its
 * gradient: loss_out[0] = mean loss, grad_out[n] = dL/d out[n]. */
NCF_API int ncf_bce_loss(const float* out, const float* targets, int64_t N, float* loss_out,
                 float* grad_out, void* stream);

/* torch.optim.Adam (L2-coupled weight decay) on a flat buffer (trainer.py:71-75, 285). */
NCF_API int ncf_dense_adam(float* w, const float* g, float* m, float* v, int64_t n,
                   const ncf_adam_cfg* adam, void* stream);

/* One iteration of ModelTrainer.train_epoch's loop body (trainer.py:253-289): forward, BCELoss,
 * zero_grad, backward, Adam on the dense flat buffer and on the tables.  loss_out: device float. */
NCF_API int ncf_train_step(const ncf_run_cfg* cfg, const ncf_adam_cfg* adam, const ncf_tables* tables,
                   float* dense, float* dense_grad, float* dense_m, float* dense_v,
                   const int64_t* user_ids, const int64_t* item_ids, const float* targets,
                   int64_t N, float* out, float* loss_out,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* ---- individual kernels (also used by the unit tests) ---------------------------------- */
/* K1: fused dual-tower gather + LayerNorm + GMF product (architecture.py:286-287, 305-312).
 * mf_pred[N] = mf_output(LN(U_mf[u]) * LN(P_mf[p])); xu/xp [N,64] = mlp_norm of the MLP rows.
 * y_item_mf / y_user_mf (optional, [N,64]) receive mf_norm(P_mf[p]) / mf_norm(U_mf[u]) for the backward. */
NCF_API int ncf_gather_ln_gmf_fwd(const ncf_tables* tables, const float* dense,
                          const int64_t* user_ids, const int64_t* item_ids, int64_t N,
                          const int64_t* hour, const float* tmod,
                          float* mf_pred, float* xu, float* xp, float* y_item_mf, float* y_user_mf,
                          void* stream);

/* The same kernel writing xu / xp as bf16 rows [N,64] (128 B per row): the form ncf_forward hands to the fused
 * tcgen05 attention block (precision NCF_BF16_TC, S = 5), which would round them to bf16 anyway.  The saved
 * mf_norm rows (y_item_mf / y_user_mf, optional) are bf16 rows too in this variant: only
 * ncf_emb_bwd_adam_both_bf16 reads them, and it sums per id in fp32. */
NCF_API int ncf_gather_ln_gmf_fwd_bf16(const ncf_tables* tables, const float* dense,
                               const int64_t* user_ids, const int64_t* item_ids, int64_t N,
                               const int64_t* hour, const float* tmod,
                               float* mf_pred, void* xu_bf16, void* xp_bf16, void* y_item_mf_bf16, void* y_user_mf_bf16,
                               void* stream);

/* get_user_embeddings / get_product_embeddings rows (architecture.py:383-407): LN'd rows of one
 * side. side 0 = user, 1 = item. */
NCF_API int ncf_gather_ln(const ncf_tables* tables, const float* dense, int32_t side,
                  const int64_t* ids, int64_t n, float* mf_out, float* mlp_out, void* stream);

/* K6: sorted-id fused embedding backward + Adam for one side (0 user, 1 item).  d_mf_pred [N],
 * d_x [N,64] = gradient wrt the LN'd MLP row of this side.  The GMF product needs the OTHER side's
 * LN'd MF row of every sample: other_y_mf [N,64] if the forward saved it, else NULL = gather it from
 * the other side's table, which must then still hold the forward's values (so with the Adam modes
 * run the item side first with NULL, then the user side with the saved item rows).  Sort workspace
 * from ncf_emb_bwd_workspace_bytes. */
NCF_API int64_t ncf_emb_bwd_workspace_bytes(int64_t N);
NCF_API int ncf_emb_bwd_adam(const ncf_adam_cfg* adam, const ncf_tables* tables, const float* dense,
                     float* dense_grad, int32_t side,
                     const int64_t* user_ids, const int64_t* item_ids, int64_t N,
                     const float* d_mf_pred, const float* d_x, const float* other_y_mf,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* Both sides in one call with a single radix sort (what ncf_backward runs): item side first, then the
 * user side with y_item_mf = the item rows saved by ncf_gather_ln_gmf_fwd.  y_user_mf (optional) = the
 * user rows it saved: with them no table row is gathered or LayerNorm-ed a second time. */
NCF_API int ncf_emb_bwd_adam_both(const ncf_adam_cfg* adam, const ncf_tables* tables, const float* dense,
                          float* dense_grad, const int64_t* user_ids, const int64_t* item_ids, int64_t N,
                          const float* d_mf_pred, const float* d_xu, const float* d_xp, const float* y_item_mf,
                          const float* y_user_mf, void* workspace, int64_t workspace_bytes, void* stream);

/* The same with the four per-sample row arrays as bf16 rows [N,64] (what ncf_backward runs with precision
 * NCF_BF16_TC and S = 5: rows written by ncf_gather_ln_gmf_fwd_bf16 and by the fused attention backward); the
 * per-id sums, the LayerNorm backward and Adam stay fp32.  trainer.py:284-285 (loss.backward + optimizer.step). */
NCF_API int ncf_emb_bwd_adam_both_bf16(const ncf_adam_cfg* adam, const ncf_tables* tables, const float* dense,
                          float* dense_grad, const int64_t* user_ids, const int64_t* item_ids, int64_t N,
                          const float* d_mf_pred, const void* d_xu_bf16, const void* d_xp_bf16,
                          const void* y_item_mf_bf16, const void* y_user_mf_bf16, void* workspace,
                          int64_t workspace_bytes, void* stream);

/* the "every untouched row" half of NCF_EMB_ADAM_DENSE_EQUIV; clears tables->touched. */
NCF_API int ncf_emb_adam_sweep(const ncf_adam_cfg* adam, const ncf_tables* tables, void* stream);

/* TemporalEncoding.forward (architecture.py:86-94): out [n,32]. */
NCF_API int ncf_temporal_fwd(const float* hour_embed, const float* day_embed, const float* month_embed,
                     const float* pe, const int64_t* hour, const int64_t* day, const int64_t* month,
                     const int64_t* days_since, int64_t n, float* out, void* stream);

/* forward_simple hour path tables (architecture.py:433-444, 454-456, 466-468):
 * tmod[h] = 1 + 0.3*(proj_w . hour_embed[h] + proj_b)  [24,64]
 * tail1[h] = mlp.0.weight[:,64:96] . hour_embed[h]      [24,256] */
NCF_API int ncf_temporal_tables(const float* hour_embed, const float* proj_w, const float* proj_b,
                        const float* dense, float* tmod, float* tail1, void* stream);

/* keep-mask of dropout site (0 attention probs [B,H,S,S]; 1,2,3 MLP layers [N,256|128|64]) exactly
 * as the forward kernels draw it; used by the parity tests to feed the oracle. */
NCF_API int ncf_dropout_mask(const ncf_run_cfg* cfg, int32_t site, int64_t numel, uint8_t* keep, void* stream);

/* ---- full-catalogue scoring (app.py:43-77) ---------------------------------------------- */
/* eval-mode factorisation: P_hat [I,64], g [I] with logit(u,i) = LN_mf(U_mf[u]).P_hat[i] + g[i]. */
NCF_API int64_t ncf_item_fold_workspace_bytes(int64_t I);
NCF_API int ncf_item_fold(const ncf_tables* tables, const float* dense, float* p_hat, float* g,
                  void* workspace, int64_t workspace_bytes, void* stream);
/* scores + top-k per user: order = score descending, ties -> lowest item index (nlargest keep='first').
 * topk_idx int64 [n_users,k], topk_score fp32 [n_users,k]. */
NCF_API int64_t ncf_score_topk_workspace_bytes(int64_t n_users, int64_t I, int32_t k);
NCF_API int ncf_score_topk(const ncf_tables* tables, const float* dense, const float* p_hat, const float* g,
                   const int64_t* user_ids, int64_t n_users, int64_t I, int32_t k,
                   int64_t* topk_idx, float* topk_score,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* The same result through a tensor-core pre-filter (large catalogues, many users per call): item tiles as
 * bf16 operand images + per-item error margin (ncf_item_image, built once per fold); a tcgen05 GEMM bounds
 * every logit from above (|error| <= 1.01 * 2^-7 ||u|| ||p_i||) and only pairs that can enter a list are re-scored
 * with the exact fp32 arithmetic of ncf_score_topk: bit-identical indices and scores. */
NCF_API int64_t ncf_item_image_bytes(int64_t I);
NCF_API int ncf_item_image(const float* p_hat, const float* g, int64_t I, void* image, void* stream);
NCF_API int64_t ncf_score_topk_tc_workspace_bytes(int64_t n_users, int64_t I, int32_t k);
NCF_API int ncf_score_topk_tc(const ncf_tables* tables, const float* dense, const float* p_hat, const float* g,
                      const void* image, const int64_t* user_ids, int64_t n_users, int64_t I, int32_t k,
                      int64_t* topk_idx, float* topk_score,
                      void* workspace, int64_t workspace_bytes, void* stream);

/* Exact top-k of sigmoid(q . v_i + bias_i) for n raw 64-d query rows against I vectors, same kernels and the same order
 * (score descending, ties -> lowest index) as ncf_score_topk: the nearest-neighbour step behind the embedding export
 * (src/inference/generate_embeddings.py:184-236 feeds a cosine Tree-AH index, setup_tree_ah_endpoint.py:25-32; with
 * L2-normalised rows the dot product IS the cosine, and the monotone sigmoid keeps the sort keys positive).
 * image: ncf_item_image(vectors, bias) for the tensor-core pre-filter, or NULL. */
NCF_API int64_t ncf_dot_topk_workspace_bytes(int64_t n, int64_t I, int32_t k, int32_t with_image);
NCF_API int ncf_dot_topk(const float* queries, int64_t n, const float* vectors, const float* bias, const void* image, int64_t I,
                 int32_t k, int64_t* topk_idx, float* topk_score, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- row-wise sharding (torchrec ROW_WISE convention) ------------------------- */
/* block = ceil(rows/world); owner = id / block; local = id % block.  Buckets ids by owner:
 * counts[world], perm[n] (stable: position of each id in owner-major order), local_ids[n]
 * written in owner-major order. */
NCF_API int64_t ncf_shard_bucketize_workspace_bytes(int64_t n, int32_t world);
NCF_API int ncf_shard_bucketize(const int64_t* ids, int64_t n, int64_t rows, int32_t world,
                        int64_t* counts, int64_t* perm, int64_t* local_ids,
                        void* workspace, int64_t workspace_bytes, void* stream);

/* Same with adjacent-run compression: equal ADJACENT ids (the user id repeated over the S rows of an
 * interaction, data_prep.py:286-303) are exchanged once.  counts[world] = run heads per owner;
 * local_ids: the run heads' local ids in owner-major order (first sum(counts) entries valid);
 * pos[n] = index into that order (= into the exchanged row buffers) for EVERY sample. */
NCF_API int64_t ncf_shard_bucketize_runs_workspace_bytes(int64_t n, int32_t world);
NCF_API int ncf_shard_bucketize_runs(const int64_t* ids, int64_t n, int64_t rows, int32_t world,
                             int64_t* counts, int64_t* local_ids, int64_t* pos,
                             void* workspace, int64_t workspace_bytes, void* stream);

/* Full de-duplication of BOTH sides with one radix sort: every distinct id of the batch is exchanged once.
 * counts [2][world] (side-major) = distinct ids per owner; local_ids [2][N]: the distinct ids' local ids in
 * owner-major (= ascending id) order, first sum(counts[side]) entries of each half valid; pos [2][N] =
 * index of every sample's id in that order.  route_ws keeps the sorted ids for ncf_shard_backward. */
NCF_API int64_t ncf_shard_route_workspace_bytes(int64_t N);
NCF_API int ncf_shard_route(const int64_t* user_ids, const int64_t* item_ids, int64_t N,
                    int64_t rows_user, int64_t rows_item, int32_t world,
                    int64_t* counts, int64_t* local_ids, int64_t* pos,
                    void* route_ws, int64_t route_ws_bytes, void* stream);

/* ---- row-sharded step: owner-side and requester-side halves ------------------- */
/* owner: rows [n,128] = [mf_norm(T_mf[id]) | mlp_norm(T_mlp[id])] of its LOCAL ids for one side. */
NCF_API int ncf_shard_owner_rows(const ncf_tables* local_tables, const float* dense, int32_t side,
                         const int64_t* local_ids, int64_t n, float* rows, void* stream);
/* requester: forward on its own N samples from the received rows; sample n uses rows_u[pos_u[n]]
 * and rows_i[pos_i[n]].  Same workspace contract as ncf_forward. */
NCF_API int ncf_shard_forward(const ncf_run_cfg* cfg, const float* dense, const float* rows_u, const float* rows_i,
                      const int64_t* pos_u, const int64_t* pos_i, int64_t N, float* out,
                      void* workspace, int64_t workspace_bytes, void* stream);
/* requester: backward; writes ONE upstream gradient row [128] = [d/d mf_norm row | d/d mlp_norm row] per
 * exchanged row at its owner-order position (samples that share a position must be adjacent, as
 * ncf_shard_bucketize_runs arranges: their gradients are summed here; or anywhere in the batch when
 * route_ws = the buffer ncf_shard_route filled for this batch, NULL otherwise) and accumulates the dense
 * gradients (except the LayerNorm affine gradients of mf_norm / mlp_norm, which the owners add). */
NCF_API int ncf_shard_backward(const ncf_run_cfg* cfg, const float* dense, float* dense_grad,
                       const float* rows_u, const float* rows_i, const int64_t* pos_u, const int64_t* pos_i,
                       int64_t N, const float* grad_out, float* grad_rows_u, float* grad_rows_i,
                       const void* route_ws, void* workspace, int64_t workspace_bytes, void* stream);
/* owner: sorted-id segment sum of the received gradient rows, LayerNorm backward once per unique
 * local id, fused Adam (adam->emb_mode as in ncf_backward); workspace: ncf_emb_bwd_workspace_bytes(n). */
NCF_API int ncf_shard_owner_update(const ncf_adam_cfg* adam, const ncf_tables* local_tables, const float* dense,
                           float* dense_grad, int32_t side, const int64_t* local_ids, int64_t n,
                           const float* grad_rows, void* workspace, int64_t workspace_bytes, void* stream);

/* The same in two halves, for the one-sided step below: ncf_shard_pull_rows has already written the ids a requester will
 * push rows for, so the owner sorts them early (on another stream, next to the towers) and only the segment sum + update
 * wait for the rows.  Same workspace in both calls; local_ids / n as in ncf_shard_owner_update.  (trainer.py:84-88, 284-285) */
NCF_API int ncf_shard_owner_sort(const ncf_tables* local_tables, int32_t side, const int64_t* local_ids, int64_t n,
                         void* workspace, int64_t workspace_bytes, void* stream);
NCF_API int ncf_shard_owner_update_sorted(const ncf_adam_cfg* adam, const ncf_tables* local_tables, const float* dense,
                                  float* dense_grad, int32_t side, const int64_t* local_ids, int64_t n,
                                  const float* grad_rows, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- row-sharded step over PEER MEMORY (NVLink / NVSwitch, one-sided) ------------------------------------------
 * SURVEY section 5 "stretch design": every rank maps its peers' table shards and gradient receive buffers into its own
 * address space (CUDA IPC: ncf_ipc_open), so the two bulk exchanges of a step need no collective and no owner-side
 * kernel in the forward direction:
 *   forward   the requester's gather kernel PULLS the raw rows of its distinct ids straight from the owners' HBM with
 *             128-bit loads and applies the (row-local) LayerNorm itself                       ncf_shard_pull_rows
 *   backward  the requester's sorted segment-sum kernel PUSHES every summed gradient row (and its local id) into the
 *             owner's receive buffer with 128-bit stores, fused into the kernel that forms it    ncf_shard_backward(plan)
 *   owner     after a cross-rank barrier (the step's next collective) the owner runs ncf_shard_owner_update on its
 *             receive buffer.
 * The per-step plan lives in DEVICE memory (the host fills a pinned copy and uploads it on the stream): */
#define NCF_MAX_WORLD 16
typedef struct ncf_shard_plan {
  int32_t world, rank;
  int64_t begin[2][NCF_MAX_WORLD + 1];     /* [side][o]: first of MY distinct ids (owner-major order) that owner o holds */
  const float* tab[NCF_MAX_WORLD][4];      /* peers' table shards (NCF_T_* order), pointers valid on THIS device */
  float* push_rows[2][NCF_MAX_WORLD];      /* [side][o]: owner o's receive buffer where my segment starts ([n,128] fp32) */
  int64_t* push_ids[2][NCF_MAX_WORLD];     /* likewise for the local ids of those rows */
} ncf_shard_plan;

/* CUDA IPC plumbing for the plan's peer pointers: the exporter passes (handle of the allocation that contains ptr, offset
 * of ptr inside it) to its peers by any host-side channel; ncf_ipc_open maps it (peer access enabled) and returns
 * base + offset; ncf_ipc_close takes what ncf_ipc_open returned.  handle: 64 bytes (cudaIpcMemHandle_t). */
NCF_API int ncf_ipc_export(const void* ptr, void* handle_out, int64_t* offset_out);
NCF_API int ncf_ipc_open(const void* handle, int64_t offset, void** ptr_out);
NCF_API int ncf_ipc_close(void* ptr);

/* requester: rows [n,128] = [mf_norm(T_mf[id]) | mlp_norm(T_mlp[id])] for its n = plan->begin[side][world] distinct
 * ids of one side, each row read from its owner's shard through plan->tab (local_ids: ncf_shard_route's output for the
 * side, owner-major).  plan: device pointer. */
NCF_API int ncf_shard_pull_rows(const ncf_shard_plan* plan, const float* dense, int32_t side, const int64_t* local_ids,
                        int64_t n, float* rows, void* stream);

/* ncf_shard_backward with the gradient exchange fused in: instead of grad_rows_u / grad_rows_i the summed row of every
 * distinct id is stored at plan->push_rows[side][owner] + (slot - plan->begin[side][owner]) * 128 and its local id at
 * plan->push_ids (route_ws is required: the routing of ncf_shard_route for this batch; local_ids = its output). */
NCF_API int ncf_shard_backward_push(const ncf_run_cfg* cfg, const float* dense, float* dense_grad,
                            const float* rows_u, const float* rows_i, const int64_t* pos_u, const int64_t* pos_i,
                            int64_t N, const float* grad_out, const ncf_shard_plan* plan, const int64_t* local_ids,
                            const void* route_ws, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- input pipeline on the device --------------------------------------------
 * SheetzDataset.__getitem__ + _sample_negative (data_prep.py:134-161, 181-212) + collate_recommender_batch
 * (:230-320) for B positive interactions: writes the key-major id columns user_ids / item_ids [B*S] (row
 * b*S = positive, b*S+1.. = negatives of the same user) and targets [B*S] = [1,0,..,0].
 * cdf [I] = cumulative inverse-popularity weights (float64); hist_off [U+1] / hist_items = sorted CSR of
 * every user's training items (NULL, NULL = no rejection by history). */
NCF_API int ncf_sample_batch(const int64_t* pos_user, const int64_t* pos_item, int64_t B, int32_t S,
                     const double* cdf, int64_t I, const int64_t* hist_off, const int64_t* hist_items,
                     uint64_t seed, uint64_t step, int64_t* user_ids, int64_t* item_ids, float* targets,
                     void* stream);

/* ---- ranking metrics on the device (reference src/utils/metrics.py) ---------------------------------
 * scores / targets [groups, M] row-major (targets 1.0 = relevant).  out [n_k][4] = SUMS over the groups of hit@K, ndcg@K,
 * mrr@K, map@K (calculate_hit_rate :110-134, calculate_ndcg :136-177, calculate_mrr :179-205, calculate_map :207-242);
 * the caller divides by `groups`.  Order inside a group: score descending, equal scores by index (stable). */
NCF_API int ncf_rank_metrics(const float* scores, const float* targets, int64_t groups, int32_t M, const int32_t* host_k_values,
                     int32_t n_k, double* out, void* stream);
/* out[0] = ROC AUC exactly as sklearn.roc_auc_score (metrics.py:244-252: Mann-Whitney U, ties 1/2; NaN for a single class;
 * -1 when the smaller class exceeds small_class_cap: call again with 0), out[1] = accuracy at `threshold` (:267-275),
 * out[2], out[3] = number of positives / negatives.  small_class_cap: upper bound of min(#pos, #neg) if known, else 0. */
NCF_API int64_t ncf_auc_workspace_bytes(int64_t n, int64_t small_class_cap);
NCF_API int ncf_auc(const float* scores, const float* targets, int64_t n, int64_t small_class_cap, float threshold, double* out,
            void* workspace, int64_t workspace_bytes, void* stream);

/* ---- tensor-core self test: one 128-row tcgen05 GEMM tile in the three operand arrangements the
 * towers use (0 forward A.B^T, 1 input-gradient A.B, 2 weight-gradient A^T.B); fp32 in/out. */
NCF_API int ncf_tc_selftest(int32_t mode, int32_t K, int32_t N, const float* A, const float* B, float* D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NCF_B200_H */
