// extern "C" boundary of libncf_b200.so: argument checking, workspace carving, kernel sequencing.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "ncf_tower.cuh"

namespace ncf {
static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return NCF_ERR_CUDA;
}
}  // namespace ncf

using namespace ncf;

extern "C" int ncf_version(void) { return NCF_ABI_VERSION; }
extern "C" const char* ncf_last_error(void) { return g_err; }
extern "C" int64_t ncf_launch_count(void) { return (int64_t)g_launches; }

// optional auxiliary stream for work that is independent of the main stream's kernels (ncf_train_step forks the
// id sort of the embedding backward onto it); NULL = everything on the stream argument (default)
// The stream and the events that order it against the caller's stream are kept PER DEVICE (the current device of the calling
// thread), created lazily on that device: engines on different GPUs of one process do not share them.
namespace ncf {
static AuxCtx g_aux[NCF_MAX_DEVICES];
AuxCtx* aux_ctx() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= NCF_MAX_DEVICES) dev = 0;
  return &g_aux[dev];
}
int aux_events(AuxCtx* a) {
  for (int i = 0; i < 8; ++i)
    if (!a->ev[i]) NCF_CUDA(cudaEventCreateWithFlags(&a->ev[i], cudaEventDisableTiming));
  if (!a->side) NCF_CUDA(cudaStreamCreateWithFlags(&a->side, cudaStreamNonBlocking));
  return NCF_OK;
}
int& wgrad_side_sms() {
  static int r = getenv("NCF_WGRAD_SMS") ? atoi(getenv("NCF_WGRAD_SMS")) : 28;     // measured: tools/ab_env.sh NCF_WGRAD_SMS
  return r;
}
}
extern "C" int ncf_set_aux_stream(void* stream) {
  aux_ctx()->stream = (cudaStream_t)stream;
  return NCF_OK;
}
extern "C" int ncf_set_sm_reserve(int32_t sms) {
  NCF_REQUIRE(sms >= 0 && sms < num_sms(), "set_sm_reserve: %d outside [0, %d)", sms, num_sms());
  sm_reserve() = sms;
  return NCF_OK;
}
extern "C" int ncf_set_loss_readback(float* host_loss_pinned, void* cuda_event) {
  AuxCtx* a = aux_ctx();
  a->loss_host = host_loss_pinned;
  a->loss_event = (cudaEvent_t)cuda_event;
  return NCF_OK;
}
extern "C" int64_t ncf_dense_numel(void) { return kLayout.total; }
extern "C" int64_t ncf_dense_offset(int32_t id) { return id >= 0 && id < NCF_P_COUNT ? kLayout.off[id] : -1; }
extern "C" int64_t ncf_dense_size(int32_t id) { return id >= 0 && id < NCF_P_COUNT ? kLayout.size[id] : -1; }

static int check_cfg(const ncf_run_cfg* cfg, int64_t N) {
  NCF_REQUIRE(cfg, "null run cfg");
  NCF_REQUIRE(cfg->S >= 1 && cfg->S <= NCF_MAX_S, "S=%d outside [1,%d]", cfg->S, NCF_MAX_S);
  NCF_REQUIRE(N >= 0 && N % cfg->S == 0, "N=%lld is not a multiple of S=%d (architecture.py:276)", (long long)N, cfg->S);
  NCF_REQUIRE(N < ((int64_t)1 << 31), "N too large");
  NCF_REQUIRE(cfg->dropout_p >= 0.f && cfg->dropout_p < 1.f, "dropout_p outside [0,1)");
  if (cfg->precision != NCF_FP32 && cfg->precision != NCF_BF16_TC) {
    set_error("unknown precision %d", cfg->precision);
    return NCF_ERR_UNSUPPORTED;
  }
  return NCF_OK;
}

extern "C" int64_t ncf_workspace_bytes(int64_t N, const ncf_run_cfg* cfg) {
  if (!cfg || N < 0) return -1;
  return carve_tower_ws(nullptr, std::max<int64_t>(N, 1), *cfg).total;
}

extern "C" int ncf_forward(const ncf_run_cfg* cfg, const ncf_tables* T, const float* dense, const int64_t* user_ids,
                           const int64_t* item_ids, int64_t N, const int64_t* hour, const float* tmod,
                           const float* tail1, float* out, void* workspace, int64_t workspace_bytes, void* stream) {
  NCF_TRY(check_cfg(cfg, N));
  NCF_REQUIRE(T && dense && user_ids && item_ids && out && workspace, "forward: null argument");
  NCF_REQUIRE(!hour || (tmod && tail1), "forward: the hour path needs tmod and tail1");
  if (N == 0) return NCF_OK;
  TowerWs w = carve_tower_ws(workspace, N, *cfg);
  if (workspace_bytes < w.total) {
    set_error("forward: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)w.total);
    return NCF_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  NCF_TRY(gather_ln_gmf_fwd_rows(tower_bf16_rows(*cfg), T, dense, user_ids, item_ids, N, hour, tmod, w.mf_pred, w.xu, w.xp,
                                 cfg->training ? w.y_pmf : nullptr, cfg->training ? w.y_umf : nullptr, stream));
  return tower_f32_forward(*cfg, dense, N, hour, tail1, out, w, st);
}

static int backward_impl(const ncf_run_cfg* cfg, const ncf_adam_cfg* adam, const ncf_tables* T, const float* dense,
                         float* dense_grad, const int64_t* user_ids, const int64_t* item_ids, int64_t N,
                         const float* grad_out, void* workspace, int64_t workspace_bytes, void* stream, cudaEvent_t sorted,
                         bool preswept, bool head_done = false);

extern "C" int ncf_backward(const ncf_run_cfg* cfg, const ncf_adam_cfg* adam, const ncf_tables* T, const float* dense,
                            float* dense_grad, const int64_t* user_ids, const int64_t* item_ids, int64_t N,
                            const float* grad_out, void* workspace, int64_t workspace_bytes, void* stream) {
  return backward_impl(cfg, adam, T, dense, dense_grad, user_ids, item_ids, N, grad_out, workspace, workspace_bytes, stream,
                       nullptr, false);
}

// sorted != null: the id sort of the embedding backward already runs on the auxiliary stream and signals this event;
// preswept: so does the dense-equivalent sweep of the rows this batch does not touch (before the event)
static int backward_impl(const ncf_run_cfg* cfg, const ncf_adam_cfg* adam, const ncf_tables* T, const float* dense,
                         float* dense_grad, const int64_t* user_ids, const int64_t* item_ids, int64_t N,
                         const float* grad_out, void* workspace, int64_t workspace_bytes, void* stream, cudaEvent_t sorted,
                         bool preswept, bool head_done) {
  NCF_TRY(check_cfg(cfg, N));
  NCF_REQUIRE(adam && T && dense && dense_grad && user_ids && item_ids && grad_out && workspace, "backward: null argument");
  NCF_REQUIRE(cfg->training, "backward: needs the workspace of a training-mode forward");
  if (N == 0) return NCF_OK;
  TowerWs w = carve_tower_ws(workspace, N, *cfg);
  if (workspace_bytes < w.total) {
    set_error("backward: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)w.total);
    return NCF_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  static const bool defer = !(getenv("NCF_WGRAD_JOIN_LATE") && getenv("NCF_WGRAD_JOIN_LATE")[0] == '0');     // A/B switch
  NCF_TRY(tower_f32_backward(*cfg, dense, dense_grad, N, grad_out, w, st, defer, head_done));
  struct Join {        // the embedding backward does not read the MLP weight gradients: the side stream is joined behind it
    cudaStream_t st;
    ~Join() { tower_side_join(st); }
  } join{st};
  if (adam->emb_mode != NCF_EMB_NONE) {
    if (sorted) NCF_CUDA(cudaStreamWaitEvent(st, sorted, 0));
    NCF_TRY(emb_bwd_both(adam, T, dense, dense_grad, user_ids, item_ids, N, w.d_mf, w.dxu, w.dxp, w.y_pmf, w.y_umf, w.emb, w.emb_bytes, st,
                         sorted != nullptr, preswept, sorted ? aux_ctx()->stream : nullptr, tower_bf16_rows(*cfg)));
    if (adam->emb_mode == NCF_EMB_ADAM_DENSE_EQUIV && !preswept) NCF_TRY(ncf_emb_adam_sweep(adam, T, stream));
  }
  return NCF_OK;
}

// stage entry points on the workspace of an earlier ncf_forward (profiling / unit tests)
static int stage_ws(const ncf_run_cfg* cfg, const float* dense, int64_t N, void* workspace, int64_t workspace_bytes, TowerWs* w) {
  NCF_TRY(check_cfg(cfg, N));
  NCF_REQUIRE(dense && workspace && N > 0, "stage call: null argument or empty batch");
  *w = carve_tower_ws(workspace, N, *cfg);
  if (workspace_bytes < w->total) {
    set_error("stage call: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)w->total);
    return NCF_ERR_WORKSPACE;
  }
  return NCF_OK;
}
extern "C" int ncf_attn_fwd(const ncf_run_cfg* cfg, const float* dense, int64_t N, void* workspace, int64_t workspace_bytes,
                            void* stream) {
  TowerWs w;
  NCF_TRY(stage_ws(cfg, dense, N, workspace, workspace_bytes, &w));
  return tower_attn_forward(*cfg, dense, N, w, (cudaStream_t)stream);
}
extern "C" int ncf_mlp_fwd(const ncf_run_cfg* cfg, const float* dense, int64_t N, float* out, void* workspace,
                           int64_t workspace_bytes, void* stream) {
  TowerWs w;
  NCF_TRY(stage_ws(cfg, dense, N, workspace, workspace_bytes, &w));
  NCF_REQUIRE(out, "mlp_fwd: null out");
  return tower_mlp_forward(*cfg, dense, N, nullptr, nullptr, out, w, (cudaStream_t)stream);
}
extern "C" int ncf_mlp_bwd(const ncf_run_cfg* cfg, const float* dense, float* dense_grad, int64_t N, const float* grad_out,
                           void* workspace, int64_t workspace_bytes, void* stream) {
  TowerWs w;
  NCF_TRY(stage_ws(cfg, dense, N, workspace, workspace_bytes, &w));
  NCF_REQUIRE(dense_grad && grad_out, "mlp_bwd: null argument");
  return tower_mlp_backward(*cfg, dense, dense_grad, N, grad_out, w, (cudaStream_t)stream);
}
extern "C" int ncf_attn_bwd(const ncf_run_cfg* cfg, const float* dense, float* dense_grad, int64_t N, void* workspace,
                            int64_t workspace_bytes, void* stream) {
  TowerWs w;
  NCF_TRY(stage_ws(cfg, dense, N, workspace, workspace_bytes, &w));
  NCF_REQUIRE(dense_grad, "attn_bwd: null argument");
  return tower_attn_backward(*cfg, dense, dense_grad, N, w, (cudaStream_t)stream);
}

extern "C" int ncf_bce_loss(const float* out, const float* targets, int64_t N, float* loss_out, float* grad_out,
                            void* stream) {
  NCF_REQUIRE(out && targets && loss_out && N >= 0, "bce_loss: bad argument");
  return launch_bce(out, targets, N, loss_out, grad_out, (cudaStream_t)stream);
}

extern "C" int ncf_train_step(const ncf_run_cfg* cfg, const ncf_adam_cfg* adam, const ncf_tables* T, float* dense,
                              float* dense_grad, float* dense_m, float* dense_v, const int64_t* user_ids,
                              const int64_t* item_ids, const float* targets, int64_t N, float* out, float* loss_out,
                              void* workspace, int64_t workspace_bytes, void* stream) {
  NCF_TRY(check_cfg(cfg, N));
  NCF_REQUIRE(adam && T && dense && dense_grad && dense_m && dense_v && targets && out && loss_out, "train_step: null argument");
  NCF_REQUIRE(cfg->training, "train_step: cfg->training must be set");
  NCF_REQUIRE(N > 0, "train_step: empty batch");
  cudaStream_t st = (cudaStream_t)stream;
  TowerWs w = carve_tower_ws(workspace, N, *cfg);
  if (workspace_bytes < w.total) {
    set_error("train_step: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)w.total);
    return NCF_ERR_WORKSPACE;
  }
  // The id sort of the embedding backward depends on the ids only: with an auxiliary stream (ncf_set_aux_stream) it
  // is forked off here and runs next to the gather / attention kernels, which leave SM resources free.  So does the
  // dense-equivalent sweep: it updates exactly the rows this batch does NOT name, which nothing else in the step reads
  // or writes, and the tower kernels it overlaps are not HBM-bound.
  // K1 is launched first so that a caller who has just synchronised (a training loop reading its loss every step)
  // gets the main stream going before the dozen launches of the auxiliary stream are issued.
  NCF_REQUIRE(T && user_ids && item_ids && workspace, "train_step: null argument");
  cudaEvent_t sorted = nullptr;
  bool preswept = false;
  AuxCtx* aux = aux_ctx();
  cudaStream_t g_aux_stream = aux->stream;
  const bool fork = g_aux_stream && adam->emb_mode != NCF_EMB_NONE && adam->emb_mode != NCF_EMB_MATERIALIZE;
  cudaEvent_t g_ev_fork = nullptr, g_ev_sorted = nullptr;
  if (fork) {
    NCF_TRY(aux_events(aux));
    g_ev_fork = aux->ev[0];
    g_ev_sorted = aux->ev[1];
    NCF_CUDA(cudaEventRecord(g_ev_fork, st));                  // the previous step's K6 has released the sort buffers
  }
  NCF_CUDA(cudaMemsetAsync(dense_grad, 0, sizeof(float) * kLayout.total, st));                       // optimizer.zero_grad()
  NCF_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));      // here, not in front of the loss kernel: see launch_bce
  NCF_TRY(gather_ln_gmf_fwd_rows(tower_bf16_rows(*cfg), T, dense, user_ids, item_ids, N, nullptr, nullptr, w.mf_pred, w.xu, w.xp,
                                 w.y_pmf, w.y_umf, stream));
  if (fork) {
    NCF_CUDA(cudaStreamWaitEvent(g_aux_stream, g_ev_fork, 0));
    NCF_TRY(emb_sort_both(T, user_ids, item_ids, N, w.emb, w.emb_bytes, g_aux_stream));
    static const bool early_sweep = !(getenv("NCF_EARLY_SWEEP") && getenv("NCF_EARLY_SWEEP")[0] == '0');   // A/B switch
    if (adam->emb_mode == NCF_EMB_ADAM_DENSE_EQUIV && early_sweep) {
      NCF_TRY(emb_sweep_early(adam, T, user_ids, item_ids, N, g_aux_stream));
      preswept = true;
    }
    NCF_CUDA(cudaEventRecord(g_ev_sorted, g_aux_stream));
    sorted = g_ev_sorted;
  }
  NCF_TRY(tower_f32_forward(*cfg, dense, N, nullptr, nullptr, out, w, st));
  // bf16 towers: loss, its gradient and the scalar half of the head's backward in ONE kernel (bce_head_bwd_kernel); otherwise
  // the BCELoss gradient goes into the (not yet used) backward scratch g128b
  static const bool fuse_head = !(getenv("NCF_FUSE_HEAD") && getenv("NCF_FUSE_HEAD")[0] == '0');     // A/B switch
  const bool head_done = fuse_head && cfg->precision == NCF_BF16_TC;
  if (head_done) NCF_TRY(launch_bce_head_bwd(out, targets, N, loss_out, dense, dense_grad, w, st));
  else NCF_TRY(launch_bce(out, targets, N, loss_out, w.g128b, st, true));
  if (aux->loss_host && aux->loss_event) {        // ncf_set_loss_readback: the loss is final here
    NCF_CUDA(cudaMemcpyAsync(aux->loss_host, loss_out, sizeof(float), cudaMemcpyDeviceToHost, st));
    NCF_CUDA(cudaEventRecord(aux->loss_event, st));
  }
  NCF_TRY(backward_impl(cfg, adam, T, dense, dense_grad, user_ids, item_ids, N, w.g128b, workspace, workspace_bytes, stream, sorted,
                        preswept, head_done));
  return ncf_dense_adam(dense, dense_grad, dense_m, dense_v, kLayout.total, adam, stream);
}
