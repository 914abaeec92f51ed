// Internal interface between the C-ABI glue (ncf_abi.cu) and the tower implementations.
#pragma once
#include <algorithm>

#include "ncf_common.cuh"

namespace ncf {

// activations kept between forward and backward, carved out of the caller's workspace
constexpr int MLP_WG_PART_COLS = 448 + 3;    // weight-gradient TMEM columns + three bias-gradient columns
struct TowerWs {
  float *mf_pred, *mlp_pred, *p_saved;     // [N]
  float *xu, *xp;                          // [N,64]  mlp_norm(user row), mlp_norm(item row)
  float *y_pmf, *y_umf;                    // [N,64]  mf_norm(item row), mf_norm(user row): training only
  float *q, *kv, *ctx, *a;                 // [N,64] [N,128] [N,64] [N,64]
  float *r1, *y1, *r2, *y2, *r3, *y3;      // relu outputs / layer outputs of the 3 MLP layers
  // backward scratch (training only)
  float *d_mf, *d_mlp;                     // [N] dL/d mf_pred, dL/d mlp_pred
  float *g64a, *g64b, *g128, *g128b, *g256, *g256b;
  float *dxu, *dxp;                        // aliases set by the backward
  // bf16 tensors of the tcgen05 path (NCF_BF16_TC): saved activations and pre-activation gradients
  void *r1b, *y1b, *r2b, *y2b, *r3b, *dz1b, *dz2b, *dz3b;
  float* wg_partial;                       // per-CTA accumulators of the tcgen05 MLP wgrad kernel: [SMs][MLP_WG_PART_COLS][128]
  float* at_partial;                       // same for the fused attention backward
  void* a_img;                             // attention output as bf16 tile image [ceil(N/128)][128 x 64]
  float *st1, *st2, *st3;                  // LayerNorm (mean, rstd) per row of the three MLP layers
  char* emb;                               // workspace of the fused embedding backward
  int64_t emb_bytes;
  int64_t total;
};

TowerWs carve_tower_ws(void* ws, int64_t N, const ncf_run_cfg& cfg);
int tower_f32_forward(const ncf_run_cfg& cfg, const float* dense, int64_t N, const int64_t* hour, const float* tail1,
                      float* out, TowerWs& w, cudaStream_t st);
// defer_join: the caller calls tower_side_join(st) once it has enqueued the work that does not need the MLP weight gradients
int tower_f32_backward(const ncf_run_cfg& cfg, const float* dense, float* dense_grad, int64_t N, const float* grad_out,
                       TowerWs& w, cudaStream_t st, bool defer_join = false, bool head_done = false);
int tower_side_join(cudaStream_t st);
// True when the attention block runs as the fused tcgen05 kernels (bf16 towers, S = 5): then its row-vector
// interfaces w.xu, w.xp (from K1 / the sharded requester) and da = w.g64a (from the MLP backward) hold bf16
// [N,64] rows instead of fp32 - the same values the operand tiles would be rounded to anyway, half the traffic.
// So do the rows that only the embedding backward reads: the LayerNorm-ed MF rows K1 saves (w.y_pmf, w.y_umf) and the
// attention backward's dxu / dxp (summed per id in fp32 by K6).
bool tower_bf16_rows(const ncf_run_cfg& cfg);
// K1 with selectable row format (ncf_embed.cu); the C-ABI export ncf_gather_ln_gmf_fwd is the fp32 case
int gather_ln_gmf_fwd_rows(bool bf16_rows, const ncf_tables* T, const float* dense, const int64_t* user_ids,
                           const int64_t* item_ids, int64_t N, const int64_t* hour, const float* tmod, float* mf_pred,
                           float* xu, float* xp, float* y_item_mf, float* y_user_mf, void* stream);
// the two halves of each direction (also exported one by one: ncf_attn_fwd / ncf_mlp_fwd / ncf_mlp_bwd / ncf_attn_bwd)
int tower_attn_forward(const ncf_run_cfg& cfg, const float* dense, int64_t N, TowerWs& w, cudaStream_t st);
int tower_mlp_forward(const ncf_run_cfg& cfg, const float* dense, int64_t N, const int64_t* hour, const float* tail1,
                      float* out, TowerWs& w, cudaStream_t st);
int tower_mlp_backward(const ncf_run_cfg& cfg, const float* dense, float* dense_grad, int64_t N, const float* grad_out,
                       TowerWs& w, cudaStream_t st, cudaStream_t side = nullptr, int side_sms = 0, cudaEvent_t fork = nullptr,
                       cudaEvent_t join = nullptr, bool head_done = false);
int tower_attn_backward(const ncf_run_cfg& cfg, const float* dense, float* dense_grad, int64_t N, TowerWs& w, cudaStream_t st);
// tcgen05 MLP tower (ncf_tower_tc.cu): forward from w.a (fills w.mlp_pred, w.p_saved, out; w.y3 stays unused)
int mlp_tc_forward(const ncf_run_cfg& cfg, const float* dense, int64_t N, const int64_t* hour, const float* tail1,
                   float* out, TowerWs& w, cudaStream_t st);
// backward from w.d_mlp (dL/d mlp_pred per row) to da = w.g64a; accumulates the MLP parameter gradients incl. mlp_output.weight
// side != null: the weight-gradient kernel (HBM-bound, independent of everything up to the dense Adam) is launched on that
// stream with side_sms CTAs behind the event `fork` and signals `join`; the caller's next tower kernel leaves those SMs free
int mlp_tc_backward(const ncf_run_cfg& cfg, const float* dense, float* dense_grad, int64_t N, TowerWs& w, cudaStream_t st,
                    cudaStream_t side = nullptr, int side_sms = 0, cudaEvent_t fork = nullptr, cudaEvent_t join = nullptr);
// tcgen05 projections of the attention block (fp32 tensors, bf16 operands)
int tc_proj_forward(int which, const float* X, const float* W, const float* bias, float* Y, int64_t N, cudaStream_t st);
// 64 -> 64 projection whose output is written as a bf16 tile image (the MLP kernels' A operand)
int tc_proj_forward_img(const float* X, const float* W, const float* bias, void* img, int64_t N, cudaStream_t st);
int tc_proj_dgrad(int which, const float* dY, const float* W, float* dX, int64_t N, cudaStream_t st);
int tc_proj_wgrad(int which, const float* Z, const float* X, float* dW, float* db, int64_t N, cudaStream_t st);
// fused attention block on tcgen05 for S = 5 (ncf_attn_tc.cu): forward xu, xp -> a_img; backward da (w.g64a) ->
// dxu (w.g64b), dxp (w.g256) + the attention parameter gradients, recomputing q, k, v and the probabilities
int attn_tc_forward(const ncf_run_cfg& cfg, const float* dense, int64_t N, TowerWs& w, cudaStream_t st);
// reduce_st != null: the reduction of the per-CTA weight-gradient partials runs there, behind `done` recorded on st
int attn_tc_backward(const ncf_run_cfg& cfg, const float* dense, float* dense_grad, int64_t N, TowerWs& w, cudaStream_t st,
                     int leave_sms = 0, cudaStream_t reduce_st = nullptr, cudaEvent_t done = nullptr);
int64_t attn_tc_partial_floats();
// fused embedding backward of both sides with a single radix sort (ncf_embed.cu)
int emb_bwd_both(const ncf_adam_cfg* adam, const ncf_tables* T, const float* dense, float* dense_grad,
                 const int64_t* user_ids, const int64_t* item_ids, int64_t N, const float* d_mf_pred, const float* dxu,
                 const float* dxp, const float* y_item_mf, const float* y_user_mf, void* workspace, int64_t workspace_bytes,
                 cudaStream_t st, bool presorted = false, bool preswept = false, cudaStream_t side_stream = nullptr,
                 bool rows_bf16 = false);
// ncf_set_aux_stream (ncf_abi.cu): per-device auxiliary stream (null = none) + the events that order it
constexpr int NCF_MAX_DEVICES = 64;
struct AuxCtx {
  cudaStream_t stream = nullptr;
  float* loss_host = nullptr;          // ncf_set_loss_readback
  cudaEvent_t loss_event = nullptr;
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};     // 0 fork, 1 sorted (ncf_train_step); 2 fork, 3 join (emb_bwd_both); 4 fork, 5 join, 6 attention backward done (side stream)
  bool side_pending = false;           // the side stream's weight-gradient kernels have not been joined into the caller's stream yet
  cudaStream_t side = nullptr;         // library-owned second stream (created with the first use): MLP weight gradients next to the attention backward
};
AuxCtx* aux_ctx();
int aux_events(AuxCtx* a);
// SMs given to the MLP weight-gradient kernel while the attention backward runs on the others (0 = one after the other);
// only with an auxiliary stream set (the opt-in for concurrency inside a step)
int& wgrad_side_sms();
int emb_sweep_early(const ncf_adam_cfg* adam, const ncf_tables* T, const int64_t* user_ids, const int64_t* item_ids, int64_t N,
                    cudaStream_t st);
int emb_sort_both(const ncf_tables* T, const int64_t* user_ids, const int64_t* item_ids, int64_t N, void* workspace,
                  int64_t workspace_bytes, cudaStream_t st);
// sharded requester: de-duplicated routing of a batch and the per-distinct-id gradient rows (ncf_embed.cu)
int shard_requester_grads(const float* dense, float* dense_grad, const float* rows_u, const float* rows_i, const int64_t* pos_u,
                          const int64_t* pos_i, int64_t N, const float* d_mf, const float* dxu, const float* dxp,
                          const void* route_ws, float* gu, float* gi, void* emb_ws, int64_t emb_ws_bytes, cudaStream_t st,
                          const ncf_shard_plan* plan = nullptr, const int64_t* local_ids = nullptr, bool rows_bf16 = false,
                          const float* y_item_mf = nullptr, const float* y_user_mf = nullptr);
// fused backward of one projection: dX = dY.W and dW += dY^T.X, db += colsum(dY) (which: 0 = 64 cols, 1 = 128)
int tc_proj_backward(int which, const float* dY, const float* X, const float* W, float* dX, float* dW, float* db, int64_t N,
                     cudaStream_t st);
int launch_bce(const float* out, const float* targets, int64_t N, float* loss_out, float* grad_out, cudaStream_t st,
               bool loss_zeroed = false);
int launch_bce_head_bwd(const float* out, const float* targets, int64_t N, float* loss_out, const float* dense, float* dense_grad,
                        TowerWs& w, cudaStream_t st);

}  // namespace ncf
