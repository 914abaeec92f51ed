// Row-sharded embedding path: the tables live on their owner GPUs, the towers stay
// data-parallel.  Per step and per side (user / item):
//   requester: bucketize ids by owner -> all-to-all ids                       (ncf_shard_bucketize)
//   owner:     gather + LayerNorm its rows -> [n,128] = [mf_norm(row) | mlp_norm(row)]
//              -> all-to-all rows back                                        (ncf_shard_owner_rows)
//   requester: GMF product + tower forward / backward on its own samples       (ncf_shard_forward,
//              upstream gradients per sample, packed in owner order             ncf_shard_backward)
//              -> all-to-all gradient rows to the owners
//   owner:     sorted-id segment sum of the upstream rows, LayerNorm backward once per unique id,
//              fused Adam                                                      (ncf_shard_owner_update)
// LayerNorm is row-local, so it runs at the owner in both directions; its affine gradients and all
// tower gradients go through one flat all-reduce.
#include <string.h>

#include "ncf_tower.cuh"

namespace ncf {

// [n,128] rows of one side for the owner's local ids
__global__ void __launch_bounds__(256) owner_rows_kernel(const float* __restrict__ t_mf, const float* __restrict__ t_mlp,
                                                          const float* __restrict__ dense,
                                                          const int64_t* __restrict__ local_ids, int64_t n,
                                                          float* __restrict__ rows) {
  const int lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float* tab = half ? t_mlp : t_mf;
  const float4 g = ldg4(dense + (half ? NCF_OFF(NCF_P_MLP_NORM_W) : NCF_OFF(NCF_P_MF_NORM_W)) + 4 * l16);
  const float4 b = ldg4(dense + (half ? NCF_OFF(NCF_P_MLP_NORM_B) : NCF_OFF(NCF_P_MF_NORM_B)) + 4 * l16);
  for (int64_t r = warp; r < n; r += nwarps) {
    const float4 x = ldg4(tab + local_ids[r] * D + 4 * l16);
    const float mean = half_warp_sum(f4_hsum(x)) * (1.0f / 64.0f);
    const float4 d = make_float4(x.x - mean, x.y - mean, x.z - mean, x.w - mean);
    const float rstd = rsqrtf(half_warp_sum(f4_dot(d, d)) * (1.0f / 64.0f) + LN_EPS);
    st4(rows + r * 2 * D + half * D + 4 * l16,
        make_float4(fmaf(d.x * rstd, g.x, b.x), fmaf(d.y * rstd, g.y, b.y), fmaf(d.z * rstd, g.z, b.z),
                    fmaf(d.w * rstd, g.w, b.w)));
  }
}

// One-sided forward: the requester reads the RAW rows of its distinct ids straight out of the owners'
// shards (peer-mapped pointers in the plan, 128-bit loads over NVLink for remote owners) and applies the row-local
// LayerNorm itself: gather, LayerNorm and the "collective" are one kernel; no id exchange, no owner-side work.
__global__ void __launch_bounds__(256) pull_rows_kernel(const ncf_shard_plan* __restrict__ plan, const float* __restrict__ dense,
                                                         int side, const int64_t* __restrict__ local_ids, int64_t n,
                                                         float* __restrict__ rows) {
  __shared__ int64_t s_begin[NCF_MAX_WORLD + 1];
  __shared__ const float* s_tab[NCF_MAX_WORLD][2];
  const int world = plan->world;
  if (threadIdx.x <= world) s_begin[threadIdx.x] = plan->begin[side][threadIdx.x];
  if (threadIdx.x < world) {
    s_tab[threadIdx.x][0] = plan->tab[threadIdx.x][side];
    s_tab[threadIdx.x][1] = plan->tab[threadIdx.x][2 + side];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float4 g = ldg4(dense + (half ? NCF_OFF(NCF_P_MLP_NORM_W) : NCF_OFF(NCF_P_MF_NORM_W)) + 4 * l16);
  const float4 b = ldg4(dense + (half ? NCF_OFF(NCF_P_MLP_NORM_B) : NCF_OFF(NCF_P_MF_NORM_B)) + 4 * l16);
  // a warp takes 4 consecutive rows per trip: 4 independent 128-bit loads per lane in flight (remote latency ~2 us)
  for (int64_t r0 = warp * 4; r0 < n; r0 += nwarps * 4) {
    float4 x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t r = min(r0 + j, n - 1);
      int o = 0;
      while (o + 1 < world && r >= s_begin[o + 1]) ++o;
      const float* tab = s_tab[o][half];
      // plain (coherent) loads: the owner rewrites these rows every step, so the read-only path must not cache them
      x[j] = *reinterpret_cast<const float4*>(tab + local_ids[r] * D + 4 * l16);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t r = r0 + j;
      const float mean = half_warp_sum(f4_hsum(x[j])) * (1.0f / 64.0f);
      const float4 d = make_float4(x[j].x - mean, x[j].y - mean, x[j].z - mean, x[j].w - mean);
      const float rstd = rsqrtf(half_warp_sum(f4_dot(d, d)) * (1.0f / 64.0f) + LN_EPS);
      if (r < n)
        st4(rows + r * 2 * D + half * D + 4 * l16,
            make_float4(fmaf(d.x * rstd, g.x, b.x), fmaf(d.y * rstd, g.y, b.y), fmaf(d.z * rstd, g.z, b.z),
                        fmaf(d.w * rstd, g.w, b.w)));
    }
    // The ids this rank will send gradient rows for go into the owners' receive buffers NOW (the backward's push stores the
    // same words again next to the rows): behind a barrier the owner sorts them next to the towers (ncf_shard_owner_sort)
    // instead of between the push and its update.
    if (lane < 4 && r0 + lane < n) {
      const int64_t r = r0 + lane;
      int o = 0;
      while (o + 1 < world && r >= s_begin[o + 1]) ++o;
      plan->push_ids[side][o][r - s_begin[o]] = local_ids[r];
    }
  }
}

// requester: sample n reads its user rows at rows_u[pos_u[n]] and its item rows at rows_i[pos_i[n]].  Training: the two
// LayerNorm-ed MF rows of every sample are kept per SAMPLE (y_item_mf, y_user_mf, like K1 does), so that the requester's
// segment sum of the backward reads everything by sample row (the lean phase-1 kernel) instead of chasing pos -> row twice.
__global__ void __launch_bounds__(256) gmf_from_rows_kernel(const float* __restrict__ rows_u, const float* __restrict__ rows_i,
                                                            const int64_t* __restrict__ pos_u, const int64_t* __restrict__ pos_i,
                                                            const float* __restrict__ dense, int64_t N,
                                                            float* __restrict__ mf_pred, float* __restrict__ xu,
                                                            float* __restrict__ xp, float* __restrict__ y_item_mf,
                                                            float* __restrict__ y_user_mf, bool bf16_rows) {
  const int lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float4 w_out = ldg4(dense + NCF_OFF(NCF_P_MF_OUT_W) + 4 * l16);
  const float b_out = __ldg(dense + NCF_OFF(NCF_P_MF_OUT_B));
  const float* rows = half ? rows_i : rows_u;
  const int64_t* pos = half ? pos_i : pos_u;
  float* ykeep = half ? y_item_mf : y_user_mf;
  // four samples per trip: eight independent 128-bit loads per lane in flight
  for (int64_t n0 = warp * 4; n0 < N; n0 += nwarps * 4) {
    float4 y_mf[4], y_ml[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float* src = rows + pos[min(n0 + j, N - 1)] * 2 * D;
      y_mf[j] = ldg4(src + 4 * l16);
      y_ml[j] = ldg4(src + D + 4 * l16);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + j;
      const float4 other = make_float4(__shfl_xor_sync(0xffffffffu, y_mf[j].x, 16), __shfl_xor_sync(0xffffffffu, y_mf[j].y, 16),
                                       __shfl_xor_sync(0xffffffffu, y_mf[j].z, 16), __shfl_xor_sync(0xffffffffu, y_mf[j].w, 16));
      const float dot = half_warp_sum(f4_dot(f4_mul(y_mf[j], other), w_out));
      if (n < N) {
        if (lane == 0) mf_pred[n] = dot + b_out;
        st_row4(half ? xp : xu, n, 4 * l16, y_ml[j], bf16_rows);
        if (ykeep) st_row4(ykeep, n, 4 * l16, y_mf[j], bf16_rows);
      }
    }
  }
}

// requester backward: upstream gradient rows, one per EXCHANGED row (= run of adjacent samples that share it,
// ncf_shard_bucketize_runs), written at the owner-order position:
//   gu[p] = [ sum_j d_mf[j] * w_mf * y_item_mf[j] | sum_j dxu[j] ]  over the samples j of the run with pos_u[j] == p
//   gi[p] likewise with the user rows, and d mf_output.weight += sum_n d_mf[n] * y_user_mf[n] * y_item_mf[n].
// Half warp = (first sample of a run, side); the samples of a run are adjacent, so the sum is a short loop.
__global__ void __launch_bounds__(256) pack_grads_kernel(const float* __restrict__ rows_u, const float* __restrict__ rows_i,
                                                         const int64_t* __restrict__ pos_u, const int64_t* __restrict__ pos_i,
                                                         const float* __restrict__ dense, const float* __restrict__ d_mf,
                                                         const float* __restrict__ dxu, const float* __restrict__ dxp,
                                                         int64_t N, float* __restrict__ gu, float* __restrict__ gi,
                                                         float* __restrict__ dense_grad, bool rows_bf16) {
  __shared__ float s_red[8][D];
  const int lane = threadIdx.x & 31, warpi = threadIdx.x >> 5, half = lane >> 4, l16 = lane & 15;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + warpi;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float4 w_out = ldg4(dense + NCF_OFF(NCF_P_MF_OUT_W) + 4 * l16);
  const int64_t* pos_mine = half ? pos_i : pos_u;
  const int64_t* pos_other = half ? pos_u : pos_i;
  const float* rows_mine = half ? rows_i : rows_u;
  const float* rows_other = half ? rows_u : rows_i;
  const float* dx = half ? dxp : dxu;
  float* dst = half ? gi : gu;
  float4 dw = make_float4(0, 0, 0, 0);
  for (int64_t n = warp; n < N; n += nwarps) {
    const int64_t p = pos_mine[n];
    if (n > 0 && pos_mine[n - 1] == p) continue;            // not the first sample of its run (uniform per half warp)
    const float4 y_mine = ldg4(rows_mine + p * 2 * D + 4 * l16);
    float4 a_mf = make_float4(0, 0, 0, 0), a_ml = a_mf;
    for (int64_t j = n; j < N && pos_mine[j] == p; ++j) {
      const float4 y_other = ldg4(rows_other + pos_other[j] * 2 * D + 4 * l16);
      const float g = d_mf[j];
      const float4 t = make_float4(g * y_other.x, g * y_other.y, g * y_other.z, g * y_other.w);
      a_mf = f4_add(a_mf, f4_mul(t, w_out));
      a_ml = f4_add(a_ml, ld_row4(dx, j, 4 * l16, rows_bf16));
      if (half == 0) dw = f4_add(dw, f4_mul(t, y_mine));
    }
    st4(dst + p * 2 * D + 4 * l16, a_mf);
    st4(dst + p * 2 * D + D + 4 * l16, a_ml);
  }
  if (half == 0) {
    s_red[warpi][4 * l16 + 0] = dw.x; s_red[warpi][4 * l16 + 1] = dw.y;
    s_red[warpi][4 * l16 + 2] = dw.z; s_red[warpi][4 * l16 + 3] = dw.w;
  }
  __syncthreads();
  if (threadIdx.x < D) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += s_red[w][threadIdx.x];
    atomicAdd(dense_grad + NCF_OFF(NCF_P_MF_OUT_W) + threadIdx.x, s);
  }
}

}  // namespace ncf

using namespace ncf;

extern "C" int ncf_shard_owner_rows(const ncf_tables* T, const float* dense, int32_t side, const int64_t* local_ids,
                                    int64_t n, float* rows, void* stream) {
  NCF_REQUIRE(T && dense && (side == 0 || side == 1) && n >= 0, "shard_owner_rows: bad argument");
  if (n == 0) return NCF_OK;
  NCF_REQUIRE(local_ids && rows, "shard_owner_rows: null buffer");
  const int grid = (int)std::min<int64_t>((n + 7) / 8, (int64_t)num_sms() * 8);
  owner_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(T->w[side], T->w[2 + side], dense, local_ids, n, rows);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

extern "C" int ncf_shard_pull_rows(const ncf_shard_plan* plan, const float* dense, int32_t side, const int64_t* local_ids,
                                   int64_t n, float* rows, void* stream) {
  NCF_REQUIRE(plan && dense && (side == 0 || side == 1) && n >= 0, "shard_pull_rows: bad argument");
  if (n == 0) return NCF_OK;
  NCF_REQUIRE(local_ids && rows, "shard_pull_rows: null buffer");
  const int grid = (int)std::min<int64_t>((n + 31) / 32, (int64_t)num_sms() * 8);
  pull_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(plan, dense, side, local_ids, n, rows);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

// ---- CUDA IPC (peer-mapped table shards and receive buffers of the one-sided step) ---------------------------------
extern "C" int ncf_ipc_export(const void* ptr, void* handle_out, int64_t* offset_out) {
  NCF_REQUIRE(ptr && handle_out && offset_out, "ipc_export: null argument");
  // base of the allocation that contains ptr (torch sub-allocates its cudaMalloc segments): driver entry point fetched at
  // run time, so the library carries no link-time dependency on libcuda
  typedef int (*GetRange)(unsigned long long*, size_t*, unsigned long long);
  static GetRange get_range = nullptr;
  if (!get_range) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    NCF_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr));
    NCF_REQUIRE(fn && qr == cudaDriverEntryPointSuccess, "ipc_export: cuMemGetAddressRange is not available");
    get_range = reinterpret_cast<GetRange>(fn);
  }
  unsigned long long base = 0;
  size_t size = 0;
  const int rc = get_range(&base, &size, (unsigned long long)(uintptr_t)ptr);
  NCF_REQUIRE(rc == 0 && base, "ipc_export: cuMemGetAddressRange failed (%d)", rc);
  cudaIpcMemHandle_t h;
  NCF_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>((uintptr_t)base)));
  memcpy(handle_out, &h, sizeof(h));
  *offset_out = (int64_t)((unsigned long long)(uintptr_t)ptr - base);
  return NCF_OK;
}
namespace {
struct IpcMapping { void* base; int refs; char handle[sizeof(cudaIpcMemHandle_t)]; };
static IpcMapping g_ipc[256];
static int g_ipc_n = 0;
}
extern "C" int ncf_ipc_open(const void* handle, int64_t offset, void** ptr_out) {
  NCF_REQUIRE(handle && ptr_out && offset >= 0, "ipc_open: bad argument");
  // one mapping per exporting allocation (cudaIpcOpenMemHandle may be called once per handle and process)
  for (int i = 0; i < g_ipc_n; ++i)
    if (g_ipc[i].refs > 0 && memcmp(g_ipc[i].handle, handle, sizeof(cudaIpcMemHandle_t)) == 0) {
      ++g_ipc[i].refs;
      *ptr_out = static_cast<char*>(g_ipc[i].base) + offset;
      return NCF_OK;
    }
  int slot = -1;                              // a closed mapping's entry is reused
  for (int i = 0; i < g_ipc_n && slot < 0; ++i)
    if (g_ipc[i].refs == 0) slot = i;
  NCF_REQUIRE(slot >= 0 || g_ipc_n < 256, "ipc_open: too many mappings");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* base = nullptr;
  NCF_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  IpcMapping& m = g_ipc[slot >= 0 ? slot : g_ipc_n++];
  m.base = base;
  m.refs = 1;
  memcpy(m.handle, handle, sizeof(h));
  *ptr_out = static_cast<char*>(base) + offset;
  return NCF_OK;
}
extern "C" int ncf_ipc_close(void* ptr) {
  if (!ptr) return NCF_OK;
  // the mapping that contains ptr: bases are distinct virtual ranges, pick the closest base at or below ptr
  int best = -1;
  for (int i = 0; i < g_ipc_n; ++i)
    if (g_ipc[i].refs > 0 && g_ipc[i].base <= ptr && (best < 0 || g_ipc[i].base > g_ipc[best].base)) best = i;
  NCF_REQUIRE(best >= 0, "ipc_close: unknown pointer");
  if (--g_ipc[best].refs == 0) NCF_CUDA(cudaIpcCloseMemHandle(g_ipc[best].base));
  return NCF_OK;
}

static int check_shard_cfg(const ncf_run_cfg* cfg, int64_t N) {
  NCF_REQUIRE(cfg, "null run cfg");
  NCF_REQUIRE(cfg->S >= 1 && cfg->S <= NCF_MAX_S && N >= 0 && N % cfg->S == 0, "bad S / N");
  NCF_REQUIRE(cfg->precision == NCF_FP32 || cfg->precision == NCF_BF16_TC, "sharded path: unknown precision");
  return NCF_OK;
}

extern "C" int ncf_shard_forward(const ncf_run_cfg* cfg, const float* dense, const float* rows_u, const float* rows_i,
                                 const int64_t* pos_u, const int64_t* pos_i, int64_t N, float* out, void* workspace,
                                 int64_t workspace_bytes, void* stream) {
  NCF_TRY(check_shard_cfg(cfg, N));
  if (N == 0) return NCF_OK;
  NCF_REQUIRE(dense && rows_u && rows_i && pos_u && pos_i && out && workspace, "shard_forward: null argument");
  TowerWs w = carve_tower_ws(workspace, N, *cfg);
  if (workspace_bytes < w.total) {
    set_error("shard_forward: workspace %lld < %lld", (long long)workspace_bytes, (long long)w.total);
    return NCF_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)std::min<int64_t>((N + 31) / 32, (int64_t)num_sms() * 8);
  gmf_from_rows_kernel<<<grid, 256, 0, st>>>(rows_u, rows_i, pos_u, pos_i, dense, N, w.mf_pred, w.xu, w.xp,
                                             cfg->training ? w.y_pmf : nullptr, cfg->training ? w.y_umf : nullptr, tower_bf16_rows(*cfg));
  NCF_LAUNCH_CHECK();
  return tower_f32_forward(*cfg, dense, N, nullptr, nullptr, out, w, st);
}

static int shard_backward_impl(const ncf_run_cfg* cfg, const float* dense, float* dense_grad, const float* rows_u,
                               const float* rows_i, const int64_t* pos_u, const int64_t* pos_i, int64_t N,
                               const float* grad_out, float* grad_rows_u, float* grad_rows_i, const void* route_ws,
                               void* workspace, int64_t workspace_bytes, void* stream, const ncf_shard_plan* plan,
                               const int64_t* local_ids);

extern "C" int ncf_shard_backward(const ncf_run_cfg* cfg, const float* dense, float* dense_grad, const float* rows_u,
                                  const float* rows_i, const int64_t* pos_u, const int64_t* pos_i, int64_t N,
                                  const float* grad_out, float* grad_rows_u, float* grad_rows_i, const void* route_ws,
                                  void* workspace, int64_t workspace_bytes, void* stream) {
  NCF_REQUIRE(grad_rows_u && grad_rows_i, "shard_backward: null gradient row buffer");
  return shard_backward_impl(cfg, dense, dense_grad, rows_u, rows_i, pos_u, pos_i, N, grad_out, grad_rows_u, grad_rows_i, route_ws,
                             workspace, workspace_bytes, stream, nullptr, nullptr);
}

extern "C" int ncf_shard_backward_push(const ncf_run_cfg* cfg, const float* dense, float* dense_grad, const float* rows_u,
                                       const float* rows_i, const int64_t* pos_u, const int64_t* pos_i, int64_t N,
                                       const float* grad_out, const ncf_shard_plan* plan, const int64_t* local_ids,
                                       const void* route_ws, void* workspace, int64_t workspace_bytes, void* stream) {
  NCF_REQUIRE(plan && local_ids && route_ws, "shard_backward_push: needs the plan, the routed local ids and the route workspace");
  return shard_backward_impl(cfg, dense, dense_grad, rows_u, rows_i, pos_u, pos_i, N, grad_out, nullptr, nullptr, route_ws, workspace,
                             workspace_bytes, stream, plan, local_ids);
}

static int shard_backward_impl(const ncf_run_cfg* cfg, const float* dense, float* dense_grad, const float* rows_u,
                               const float* rows_i, const int64_t* pos_u, const int64_t* pos_i, int64_t N,
                               const float* grad_out, float* grad_rows_u, float* grad_rows_i, const void* route_ws,
                               void* workspace, int64_t workspace_bytes, void* stream, const ncf_shard_plan* plan,
                               const int64_t* local_ids) {
  NCF_TRY(check_shard_cfg(cfg, N));
  if (N == 0) return NCF_OK;
  NCF_REQUIRE(dense && dense_grad && rows_u && rows_i && pos_u && pos_i && grad_out && workspace,
              "shard_backward: null argument");
  NCF_REQUIRE(cfg->training, "shard_backward: needs a training-mode forward");
  TowerWs w = carve_tower_ws(workspace, N, *cfg);
  if (workspace_bytes < w.total) {
    set_error("shard_backward: workspace %lld < %lld", (long long)workspace_bytes, (long long)w.total);
    return NCF_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  NCF_TRY(tower_f32_backward(*cfg, dense, dense_grad, N, grad_out, w, st, true));
  struct Join {        // the requester's segment sum + push do not read the MLP weight gradients: joined behind them
    cudaStream_t st;
    ~Join() { tower_side_join(st); }
  } join{st};
  if (route_ws)     // ids routed by ncf_shard_route: samples that share a row are scattered -> sorted segment sum
    return shard_requester_grads(dense, dense_grad, rows_u, rows_i, pos_u, pos_i, N, w.d_mf, w.dxu, w.dxp, route_ws, grad_rows_u,
                                 grad_rows_i, w.emb, w.emb_bytes, st, plan, local_ids, tower_bf16_rows(*cfg), w.y_pmf, w.y_umf);
  const int grid = (int)std::min<int64_t>((N + 7) / 8, (int64_t)num_sms() * 8);
  pack_grads_kernel<<<grid, 256, 0, st>>>(rows_u, rows_i, pos_u, pos_i, dense, w.d_mf, w.dxu, w.dxp, N, grad_rows_u,
                                          grad_rows_i, dense_grad, tower_bf16_rows(*cfg));
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}
