// Embedding path of AdvancedNCF on sm_100a: fused dual-tower gather + LayerNorm + GMF (K1), the
// sorted-id fused embedding backward + Adam (K6), the dense-equivalent Adam sweep, the temporal
// encoding kernels and the dropout-mask dump.  All HBM-bound: rows are 256 B (64 fp32), moved as
// 128-bit vectors by half warps (16 lanes x 16 B = one row), id blocks staged into shared memory
// by cp.async.bulk (TMA bulk copy) with an mbarrier, reductions by warp shuffles.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <stdlib.h>

#include <type_traits>

#include "ncf_tower.cuh"

namespace ncf {

// ---- mbarrier / bulk-copy PTX ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t a = smem_u32(bar);
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  }
}

// LayerNorm of a 64-float row held as one float4 per lane of a half warp.  Returns the
// normalised values (x - mean) * rstd; rstd through the reference.
__device__ __forceinline__ float4 ln_normalise(float4 x, float& rstd) {
  const float mean = half_warp_sum(f4_hsum(x)) * (1.0f / 64.0f);
  const float4 d = make_float4(x.x - mean, x.y - mean, x.z - mean, x.w - mean);
  const float var = half_warp_sum(f4_dot(d, d)) * (1.0f / 64.0f);
  rstd = rsqrtf(var + LN_EPS);
  return make_float4(d.x * rstd, d.y * rstd, d.z * rstd, d.w * rstd);
}
__device__ __forceinline__ float4 affine(float4 xh, float4 g, float4 b) {
  return make_float4(fmaf(xh.x, g.x, b.x), fmaf(xh.y, g.y, b.y), fmaf(xh.z, g.z, b.z), fmaf(xh.w, g.w, b.w));
}

// =============================================================================================
// K1  fused dual-tower gather + LayerNorm + GMF product   (architecture.py:286-287, 305-312)
// =============================================================================================
constexpr int K1_THREADS = 256;
constexpr int K1_TILE = 128;  // sample rows per CTA iteration (id block = 2 x 1 KiB)

template <bool HOUR>
__global__ void __launch_bounds__(K1_THREADS) gather_ln_gmf_fwd_kernel(
    const float* __restrict__ t_umf, const float* __restrict__ t_pmf, const float* __restrict__ t_umlp,
    const float* __restrict__ t_pmlp, const float* __restrict__ dense, const int64_t* __restrict__ user_ids,
    const int64_t* __restrict__ item_ids, int64_t N, const int64_t* __restrict__ hour,
    const float* __restrict__ tmod, float* __restrict__ mf_pred, float* __restrict__ xu, float* __restrict__ xp,
    float* __restrict__ y_item_mf, float* __restrict__ y_user_mf, bool bf16_rows, int64_t rows_user, int64_t rows_item,
    int32_t* __restrict__ status) {
  __shared__ __align__(16) int64_t s_ids[2][2][K1_TILE];
  __shared__ __align__(8) uint64_t s_bar[2];
  pdl_launch_dependents();      // the attention forward that follows may start its prologue (ncf_common.cuh, PDL)

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, l16 = lane & 15;
  const int64_t num_tiles = (N + K1_TILE - 1) / K1_TILE;
  const bool bulk_ok = ((reinterpret_cast<uintptr_t>(user_ids) | reinterpret_cast<uintptr_t>(item_ids)) & 15) == 0;

  if (threadIdx.x == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // per-lane constants: this half's tables and LayerNorm affine slices
  const float* tab_mf = half ? t_pmf : t_umf;
  const float* tab_mlp = half ? t_pmlp : t_umlp;
  const int64_t tab_rows = half ? rows_item : rows_user;
  bool bad = false, bad_hour = false;      // ids outside the tables are clamped and reported (ncf_b200.h "Id validation")
  const float4 g_mf = ldg4(dense + NCF_OFF(NCF_P_MF_NORM_W) + 4 * l16);
  const float4 b_mf = ldg4(dense + NCF_OFF(NCF_P_MF_NORM_B) + 4 * l16);
  const float4 g_ml = ldg4(dense + NCF_OFF(NCF_P_MLP_NORM_W) + 4 * l16);
  const float4 b_ml = ldg4(dense + NCF_OFF(NCF_P_MLP_NORM_B) + 4 * l16);
  const float4 w_out = ldg4(dense + NCF_OFF(NCF_P_MF_OUT_W) + 4 * l16);
  const float b_out = __ldg(dense + NCF_OFF(NCF_P_MF_OUT_B));

  auto stage = [&](int64_t tile, int buf) {  // whole CTA calls; thread 0 issues the bulk copies
    const int64_t n0 = tile * K1_TILE;
    const int rows = (int)min((int64_t)K1_TILE, N - n0);
    if (bulk_ok && (rows & 1) == 0) {
      if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&s_bar[buf], 2u * rows * 8u);
        bulk_g2s(&s_ids[buf][0][0], user_ids + n0, rows * 8u, &s_bar[buf]);
        bulk_g2s(&s_ids[buf][1][0], item_ids + n0, rows * 8u, &s_bar[buf]);
      }
    } else {  // unaligned / odd tail: plain loads, then complete the same barrier phase
      for (int i = threadIdx.x; i < 2 * rows; i += K1_THREADS) {
        const int side = i >= rows, r = side ? i - rows : i;
        s_ids[buf][side][r] = side ? item_ids[n0 + r] : user_ids[n0 + r];
      }
      __syncthreads();
      if (threadIdx.x == 0) mbar_arrive_expect_tx(&s_bar[buf], 0);
    }
  };

  uint32_t phase[2] = {0, 0};
  int buf = 0;
  int64_t tile = blockIdx.x;
  if (tile < num_tiles) stage(tile, 0);
  for (; tile < num_tiles; tile += gridDim.x, buf ^= 1) {
    const int64_t next = tile + gridDim.x;
    if (next < num_tiles) stage(next, buf ^ 1);  // prefetch the next id block while this one is consumed
    mbar_wait(&s_bar[buf], phase[buf]);
    phase[buf] ^= 1;

    const int64_t n0 = tile * K1_TILE;
    const int rows = (int)min((int64_t)K1_TILE, N - n0);
    // each warp owns rows warp*16 .. warp*16+15 of the tile, four at a time (8 x 128-bit loads in flight)
#pragma unroll 1
    for (int r0 = warp * (K1_TILE / 8); r0 < (warp + 1) * (K1_TILE / 8); r0 += 4) {
      float4 a[4], b[4];
      int64_t id[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = min(r0 + j, rows - 1);
        const int64_t raw = s_ids[buf][half][r];
        bad |= bad_id(raw, tab_rows);
        id[j] = clamp_id(raw, tab_rows);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[j] = ldg4(tab_mf + id[j] * D + 4 * l16);
        b[j] = ldg4(tab_mlp + id[j] * D + 4 * l16);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = r0 + j;
        float rs;
        float4 y_mf = affine(ln_normalise(a[j], rs), g_mf, b_mf);
        float4 y_ml = affine(ln_normalise(b[j], rs), g_ml, b_ml);
        if (HOUR) {  // forward_simple hour path: item rows scaled by 1 + 0.3 * proj(hour_embed[h])
          const int64_t hraw = hour[n0 + min(r, rows - 1)];
          bad_hour |= bad_id(hraw, 24);
          const int64_t h = clamp_id(hraw, 24);
          const float4 t = ldg4(tmod + h * D + 4 * l16);
          if (half) {
            y_mf = f4_mul(y_mf, t);
            y_ml = f4_mul(y_ml, t);
          }
        }
        const float4 other = make_float4(__shfl_xor_sync(0xffffffffu, y_mf.x, 16), __shfl_xor_sync(0xffffffffu, y_mf.y, 16),
                                         __shfl_xor_sync(0xffffffffu, y_mf.z, 16), __shfl_xor_sync(0xffffffffu, y_mf.w, 16));
        const float dot = half_warp_sum(f4_dot(f4_mul(y_mf, other), w_out));
        if (r < rows) {
          const int64_t n = n0 + r;
          if (lane == 0) mf_pred[n] = dot + b_out;
          st_row4(half ? xp : xu, n, 4 * l16, y_ml, bf16_rows);
          float* ykeep = half ? y_item_mf : y_user_mf;                      // mf_norm rows kept for the backward
          if (ykeep) st_row4(ykeep, n, 4 * l16, y_mf, bf16_rows);
        }
      }
    }
    __syncthreads();  // everyone is done with s_ids[buf] before it is refilled two tiles later
  }
  if (bad) flag_status(status, half ? NCF_STATUS_BAD_ITEM_ID : NCF_STATUS_BAD_USER_ID);
  if (HOUR && bad_hour) flag_status(status, NCF_STATUS_BAD_HOUR);
}

// K1 for the reference's collate layout (data_prep.py:286-303): the rows of an interaction - the positive and its four
// negatives - are consecutive and carry the SAME user id, so of the 20 row gathers + LayerNorms the per-sample kernel above
// spends on five samples only 12 are distinct (2 user rows, 10 item rows).  A warp takes five consecutive samples at a time:
// step 0 half 0 = the user's rows, half 1 = item 0; steps 1, 2 = items 1..4, two per step; the user's mf_norm row crosses to
// the other half once.  Any five samples whose user ids differ (and the tail of the batch) go through the per-sample mapping
// inside the same kernel, so the result never depends on the grouping: same values as the kernel above, 40 % fewer LayerNorms
// on the layout every training batch has.  Measured gain at config[2]: 84 -> 79 us (the 20 half-row stores per interaction and
// the DRAM writes behind them bound the kernel, not the arithmetic).
constexpr int K1G = 5;
constexpr int K1G_TILE = 160;      // 32 groups of five sample rows per CTA iteration: 4 per warp

template <int GPT, int MINB>      // groups in flight per warp trip, CTAs per SM the register budget is set for
__global__ void __launch_bounds__(K1_THREADS, MINB) gather_ln_gmf_fwd_grouped_kernel(
    const float* __restrict__ t_umf, const float* __restrict__ t_pmf, const float* __restrict__ t_umlp,
    const float* __restrict__ t_pmlp, const float* __restrict__ dense, const int64_t* __restrict__ user_ids,
    const int64_t* __restrict__ item_ids, int64_t N, float* __restrict__ mf_pred, float* __restrict__ xu, float* __restrict__ xp,
    float* __restrict__ y_item_mf, float* __restrict__ y_user_mf, bool bf16_rows, int64_t rows_user, int64_t rows_item,
    int32_t* __restrict__ status) {
  __shared__ __align__(16) int64_t s_ids[2][2][K1G_TILE];
  __shared__ __align__(8) uint64_t s_bar[2];
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, l16 = lane & 15;
  const int64_t num_tiles = (N + K1G_TILE - 1) / K1G_TILE;
  const bool bulk_ok = ((reinterpret_cast<uintptr_t>(user_ids) | reinterpret_cast<uintptr_t>(item_ids)) & 15) == 0;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  bool bad_u = false, bad_i = false;
  const float4 g_mf = ldg4(dense + NCF_OFF(NCF_P_MF_NORM_W) + 4 * l16);
  const float4 b_mf = ldg4(dense + NCF_OFF(NCF_P_MF_NORM_B) + 4 * l16);
  const float4 g_ml = ldg4(dense + NCF_OFF(NCF_P_MLP_NORM_W) + 4 * l16);
  const float4 b_ml = ldg4(dense + NCF_OFF(NCF_P_MLP_NORM_B) + 4 * l16);
  const float4 w_out = ldg4(dense + NCF_OFF(NCF_P_MF_OUT_W) + 4 * l16);
  const float b_out = __ldg(dense + NCF_OFF(NCF_P_MF_OUT_B));

  auto stage = [&](int64_t tile, int buf) {
    const int64_t n0 = tile * K1G_TILE;
    const int rows = (int)min((int64_t)K1G_TILE, N - n0);
    if (bulk_ok && (rows & 1) == 0) {
      if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&s_bar[buf], 2u * rows * 8u);
        bulk_g2s(&s_ids[buf][0][0], user_ids + n0, rows * 8u, &s_bar[buf]);
        bulk_g2s(&s_ids[buf][1][0], item_ids + n0, rows * 8u, &s_bar[buf]);
      }
    } else {
      for (int i = threadIdx.x; i < 2 * rows; i += K1_THREADS) {
        const int side = i >= rows, r = side ? i - rows : i;
        s_ids[buf][side][r] = side ? item_ids[n0 + r] : user_ids[n0 + r];
      }
      __syncthreads();
      if (threadIdx.x == 0) mbar_arrive_expect_tx(&s_bar[buf], 0);
    }
  };
  auto xhalf = [](float4 v) {
    return make_float4(__shfl_xor_sync(0xffffffffu, v.x, 16), __shfl_xor_sync(0xffffffffu, v.y, 16),
                       __shfl_xor_sync(0xffffffffu, v.z, 16), __shfl_xor_sync(0xffffffffu, v.w, 16));
  };

  uint32_t phase[2] = {0, 0};
  int buf = 0;
  int64_t tile = blockIdx.x;
  if (tile < num_tiles) stage(tile, 0);
  for (; tile < num_tiles; tile += gridDim.x, buf ^= 1) {
    const int64_t next = tile + gridDim.x;
    if (next < num_tiles) stage(next, buf ^ 1);
    mbar_wait(&s_bar[buf], phase[buf]);
    phase[buf] ^= 1;
    const int64_t n0 = tile * K1G_TILE;
    const int rows = (int)min((int64_t)K1G_TILE, N - n0);
    const int64_t* su = s_ids[buf][0];
    const int64_t* si = s_ids[buf][1];
    // this warp's four groups, GPT per trip: the row loads of all of them are issued before the first LayerNorm
#pragma unroll 1
    for (int gp = 0; gp < 4 / GPT; ++gp) {
      float4 a[GPT][3], b[GPT][3];
      bool fast[GPT];
      int r0[GPT];
#pragma unroll
      for (int q = 0; q < GPT; ++q) {
        r0[q] = (warp * 4 + gp * GPT + q) * K1G;
        fast[q] = r0[q] + K1G <= rows;
        if (fast[q]) {
          const int64_t u0 = su[r0[q]];
#pragma unroll
          for (int j = 1; j < K1G; ++j) fast[q] = fast[q] && su[r0[q] + j] == u0;
        }
        if (fast[q]) {
          bad_u |= bad_id(su[r0[q]], rows_user);
          const int64_t uid = clamp_id(su[r0[q]], rows_user);
#pragma unroll
          for (int st = 0; st < 3; ++st) {
            const bool user = st == 0 && half == 0;
            const int j = st == 0 ? 0 : 2 * st - 1 + half;      // the item this half holds in step st
            const int64_t raw = si[r0[q] + j];
            if (!user) bad_i |= bad_id(raw, rows_item);
            const int64_t id = user ? uid : clamp_id(raw, rows_item);
            a[q][st] = ldg4((user ? t_umf : t_pmf) + id * D + 4 * l16);
            b[q][st] = ldg4((user ? t_umlp : t_pmlp) + id * D + 4 * l16);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < GPT; ++q) {
        if (r0[q] >= rows) continue;                         // warp-uniform
        const int64_t n = n0 + r0[q];
        if (fast[q]) {
          float rs;
          const float4 ymf0 = affine(ln_normalise(a[q][0], rs), g_mf, b_mf);
          const float4 yml0 = affine(ln_normalise(b[q][0], rs), g_ml, b_ml);
          const float4 crossed = xhalf(ymf0);
          const float4 yu = half ? crossed : ymf0;           // the user's mf_norm row, in both halves
          {
            const float dot = half_warp_sum(f4_dot(f4_mul(ymf0, yu), w_out));
            if (half) {                                      // item 0
              if (l16 == 0) mf_pred[n] = dot + b_out;
              st_row4(xp, n, 4 * l16, yml0, bf16_rows);
              if (y_item_mf) st_row4(y_item_mf, n, 4 * l16, ymf0, bf16_rows);
            }
          }
          // the user's rows, one copy per sample: half 0 writes mlp_norm (xu), half 1 mf_norm (kept for the backward)
#pragma unroll
          for (int j = 0; j < K1G; ++j) {
            if (!half) st_row4(xu, n + j, 4 * l16, yml0, bf16_rows);
            else if (y_user_mf) st_row4(y_user_mf, n + j, 4 * l16, yu, bf16_rows);
          }
#pragma unroll
          for (int st = 1; st < 3; ++st) {                   // items 1..4, two per step
            const float4 ym = affine(ln_normalise(a[q][st], rs), g_mf, b_mf);
            const float4 yl = affine(ln_normalise(b[q][st], rs), g_ml, b_ml);
            const float dot = half_warp_sum(f4_dot(f4_mul(ym, yu), w_out));
            const int j = 2 * st - 1 + half;
            if (l16 == 0) mf_pred[n + j] = dot + b_out;
            st_row4(xp, n + j, 4 * l16, yl, bf16_rows);
            if (y_item_mf) st_row4(y_item_mf, n + j, 4 * l16, ym, bf16_rows);
          }
        } else {
          // per-sample mapping (different users inside the five rows, or the tail of the batch)
          const int cntj = min(K1G, rows - r0[q]);
          for (int j = 0; j < cntj; ++j) {
            const int64_t raw = half ? si[r0[q] + j] : su[r0[q] + j];
            const int64_t tab_rows = half ? rows_item : rows_user;
            if (bad_id(raw, tab_rows)) { if (half) bad_i = true; else bad_u = true; }
            const int64_t id = clamp_id(raw, tab_rows);
            float rs;
            const float4 ym = affine(ln_normalise(ldg4((half ? t_pmf : t_umf) + id * D + 4 * l16), rs), g_mf, b_mf);
            const float4 yl = affine(ln_normalise(ldg4((half ? t_pmlp : t_umlp) + id * D + 4 * l16), rs), g_ml, b_ml);
            const float dot = half_warp_sum(f4_dot(f4_mul(ym, xhalf(ym)), w_out));
            if (lane == 0) mf_pred[n + j] = dot + b_out;
            st_row4(half ? xp : xu, n + j, 4 * l16, yl, bf16_rows);
            float* ykeep = half ? y_item_mf : y_user_mf;
            if (ykeep) st_row4(ykeep, n + j, 4 * l16, ym, bf16_rows);
          }
        }
      }
    }
    __syncthreads();
  }
  if (bad_u) flag_status(status, NCF_STATUS_BAD_USER_ID);
  if (bad_i) flag_status(status, NCF_STATUS_BAD_ITEM_ID);
}

// LN'd rows of one side (get_user_embeddings / get_product_embeddings, architecture.py:383-407)
__global__ void __launch_bounds__(256) gather_ln_kernel(const float* __restrict__ t_mf, const float* __restrict__ t_mlp,
                                                         const float* __restrict__ dense,
                                                         const int64_t* __restrict__ ids, int64_t n,
                                                         float* __restrict__ mf_out, float* __restrict__ mlp_out,
                                                         int64_t rows, int32_t* __restrict__ status, int status_word) {
  const int lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float* tab = half ? t_mlp : t_mf;
  float* out = half ? mlp_out : mf_out;
  const float4 g = ldg4(dense + (half ? NCF_OFF(NCF_P_MLP_NORM_W) : NCF_OFF(NCF_P_MF_NORM_W)) + 4 * l16);
  const float4 b = ldg4(dense + (half ? NCF_OFF(NCF_P_MLP_NORM_B) : NCF_OFF(NCF_P_MF_NORM_B)) + 4 * l16);
  for (int64_t r = warp; r < n; r += nwarps) {
    const int64_t raw = ids[r];
    if (bad_id(raw, rows)) flag_status(status, status_word);
    const int64_t id = clamp_id(raw, rows);
    float rs;
    const float4 y = affine(ln_normalise(ldg4(tab + id * D + 4 * l16), rs), g, b);
    if (out) st4(out + r * D + 4 * l16, y);
  }
}

// =============================================================================================
// K6  sorted-id fused embedding backward + Adam
//
// For one side (user or item) the sample rows are sorted by id, so all rows that hit the same table
// row are adjacent.  LayerNorm backward is linear in its upstream gradient for a fixed input row, so
// the upstream gradients of a run are summed FIRST and LayerNorm-backward + Adam run once per unique
// id: the per-sample row gradients are never written to memory.
//   upstream(MF row)  = sum_n d_mf_pred[n] * w_mf * LN(other side's MF row of sample n)
//   upstream(MLP row) = sum_n d_x[n]
// A warp owns a chunk of EB_CHUNK sorted positions; half 0 works on the MF tower, half 1 on the MLP
// tower.  Phase 1 streams the samples and leaves one upstream sum per run piece (512 B per unique
// id); phase 2 adds the pieces of a run in a fixed order and applies LayerNorm backward + Adam
// (deterministic, no float atomics on table rows).
// =============================================================================================
constexpr int EB_CHUNK = 32;
constexpr int EB_THREADS = 256;

struct AdamScalars {
  float lr_bc1;       // lr / (1 - beta1^t)
  float inv_sqrt_bc2; // 1 / sqrt(1 - beta2^t)
  float beta1, beta2, eps, wd;
};
inline AdamScalars adam_scalars(const ncf_adam_cfg& a) {
  AdamScalars s;
  const double bc1 = 1.0 - pow((double)a.beta1, (double)a.step);
  const double bc2 = 1.0 - pow((double)a.beta2, (double)a.step);
  s.lr_bc1 = (float)((double)a.lr / bc1);
  s.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  s.beta1 = a.beta1;
  s.beta2 = a.beta2;
  s.eps = a.eps;
  s.wd = a.weight_decay;
  return s;
}
// torch.optim.Adam single-tensor update order
__device__ __forceinline__ void adam_update(float& w, float& m, float& v, float g, const AdamScalars& s) {
  g = fmaf(s.wd, w, g);
  m = m + (g - m) * (1.0f - s.beta1);
  v = fmaf(s.beta2, v, (1.0f - s.beta2) * g * g);
  const float denom = fmaf(sqrtf(v), s.inv_sqrt_bc2, s.eps);
  w = w - s.lr_bc1 * (m / denom);
}
__device__ __forceinline__ void adam_update4(float4& w, float4& m, float4& v, float4 g, const AdamScalars& s) {
  adam_update(w.x, m.x, v.x, g.x, s);
  adam_update(w.y, m.y, v.y, g.y, s);
  adam_update(w.z, m.z, v.z, g.z, s);
  adam_update(w.w, m.w, v.w, g.w, s);
}

struct EmbBwdArgs {
  // this side's two tables (mf, mlp) and their optimizer state / dense-grad targets
  float* w[2];
  float* m[2];
  float* v[2];
  float* g[2];
  uint8_t* touched;
  uint8_t touched_val;            // what phase 2 stores for an updated row: 1 = mark for the sweep that follows, 0 = the
                                  // sweep already ran on the pre-marked flags (ncf_train_step, auxiliary stream): clear
  const float* other_mf;          // other side's MF table (for the GMF product), used when other_y == null
  const float* other_y;           // [N,64] LN'd MF row of the other side saved by the forward (or null)
  const float* own_y;             // [N,64] LN'd MF row of THIS side saved by the forward (or null): d mf_output.weight
  const float* upstream;          // sharded owner path: [N,128] ready-made upstream rows [mf | mlp] (or null)
  // sharded requester path: the LN'd rows live in the exchanged row buffers [n_unique,128] = [mf | mlp]
  const float* other_rows;        // other side's exchanged rows (or null)
  const int64_t* other_pos;       // [N] exchanged-row index of each sample on the other side
  const float* own_rows;          // this side's exchanged rows (d mf_output.weight)
  const int64_t* own_pos;
  float* out_rows;                // phase 2 output mode: [n_unique,128] summed upstream row per unique id, no update
  const int32_t* out_slot;        // [N] unique-id index of every sorted position
  const ncf_shard_plan* push;     // output mode over peer memory: rows (and ids) go to the owners' receive buffers
  const int64_t* push_local;      // [n_unique] local id of every distinct id of this side (ncf_shard_route)
  int32_t push_side;
  const int64_t* other_ids;       // other side's ids, original sample order
  int64_t other_nrows;            // rows of the other side's table (ids are clamped into it)
  const uint32_t* sorted_ids;     // this side's ids, sorted
  const int32_t* perm;            // sample row of each sorted position
  const float* d_mf_pred;         // [N]
  const uint32_t* other_sorted;   // optional: other side's id per SORTED position (pre-gathered, coalesced)
  const float* dmf_sorted;        // optional: d_mf_pred per sorted position
  const float* d_x;               // [N,64] gradient wrt this side's LN'd MLP row
  int32_t boundary_only;          // phase 2 behind the fused kernel: only the runs that leave their chunk are still to be applied
  int32_t rows_bf16;              // the per-sample [N,64] row arrays (d_x, other_y, own_y) hold bf16 rows (bf16 towers, S = 5)
  const float* dense;
  float* dense_grad;
  float* acc_buf;                 // [N][2][64] upstream sum of the run piece that starts at each sorted position
  int32_t* chunk_counter;         // phase 2: next chunk to hand out (zeroed before the launch)
  int64_t N;
  uint32_t id_off;                // sorted keys hold id + id_off (combined two-side sort)
  int32_t mode;                   // ncf_emb_mode
  int32_t accumulate_wmf;         // only one side adds d mf_output.weight
  AdamScalars adam;
};

__device__ __forceinline__ void block_flush(float* s_red, float4 val, float* dst, int lane, int warp, int nwarps,
                                            bool active_half, int l16) {
  // reduce one float4-per-lane accumulator (16 lanes x 4 = 64 columns, per half) over the block's warps
  __syncthreads();
  float* mine = s_red + (warp * 32 + lane) * 4;
  mine[0] = val.x; mine[1] = val.y; mine[2] = val.z; mine[3] = val.w;
  __syncthreads();
  if (warp == 0) {
    float4 t = make_float4(0, 0, 0, 0);
    for (int w = 0; w < nwarps; ++w) {
      const float* p = s_red + (w * 32 + lane) * 4;
      t = f4_add(t, make_float4(p[0], p[1], p[2], p[3]));
    }
    if (active_half && dst) {
      atomicAdd(dst + 4 * l16 + 0, t.x);
      atomicAdd(dst + 4 * l16 + 1, t.y);
      atomicAdd(dst + 4 * l16 + 2, t.z);
      atomicAdd(dst + 4 * l16 + 3, t.w);
    }
  }
}

// Phase 1, lean variant for the single-GPU step (both LayerNorm-ed MF rows were saved by K1): every per-sample input
// is a [N,64] row (or the scalar d_mf) indexed by the sample row, so a position costs three shuffles (row, d_mf, id),
// one 128-bit load per half (two for the side that also forms d mf_output.weight) and
// the multiply-adds - none of the source-selection logic of the general kernel below.
template <bool WMF, bool BF>
__global__ void __launch_bounds__(EB_THREADS, 3) emb_bwd_phase1_lean_kernel(EmbBwdArgs A) {
  __shared__ float s_red[(EB_THREADS / 32) * 32 * 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = EB_THREADS / 32;
  const int half = lane >> 4, l16 = lane & 15;
  const int64_t nchunks = (A.N + EB_CHUNK - 1) / EB_CHUNK;
  const int64_t gw = (int64_t)blockIdx.x * nwarps + warp, gstride = (int64_t)gridDim.x * nwarps;
  const float4 w_out = ldg4(A.dense + NCF_OFF(NCF_P_MF_OUT_W) + 4 * l16);
  const float* src = half ? A.d_x : A.other_y;
  const float* own = A.own_y;
  float4 dwout = make_float4(0, 0, 0, 0);
  // positions in flight per lane: bf16 rows are 8 B per lane, so twice as many keep the same bytes in flight
  constexpr int U = BF ? 8 : 4;
  using Raw = typename std::conditional<BF, uint2, float4>::type;
  auto load = [&](const float* base, int64_t row) -> Raw {
    if constexpr (BF) return __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(base) + row * 64 + 4 * l16));
    else return __ldg(reinterpret_cast<const float4*>(base + row * 64 + 4 * l16));
  };
  auto widen = [](Raw q) -> float4 {
    if constexpr (BF) return make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u), __uint_as_float(q.y << 16),
                                         __uint_as_float(q.y & 0xffff0000u));
    else return q;
  };

  for (int64_t c = gw; c < nchunks; c += gstride) {
    const int64_t p0 = c * EB_CHUNK;
    const int cnt = (int)min((int64_t)EB_CHUNK, A.N - p0);
    const uint32_t my_id = lane < cnt ? A.sorted_ids[p0 + lane] : 0xffffffffu;
    const int32_t my_row = lane < cnt ? A.perm[p0 + lane] : 0;
    const float my_dmf = lane < cnt ? __ldg(A.d_mf_pred + my_row) : 0.f;      // 1.3 MB array, L2-resident
    float4 acc = make_float4(0, 0, 0, 0);
    int piece_first = 0;
    uint32_t id_prev = __shfl_sync(0xffffffffu, my_id, 0);
    for (int k0 = 0; k0 < cnt; k0 += U) {
      Raw x[U], sf[U];
      float dmf[U];
      uint32_t idk[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int kk = min(k0 + u, cnt - 1);
        const int64_t srow = __shfl_sync(0xffffffffu, my_row, kk);
        dmf[u] = __shfl_sync(0xffffffffu, my_dmf, kk);
        idk[u] = __shfl_sync(0xffffffffu, my_id, kk);
        x[u] = load(src, srow);
        if (WMF) sf[u] = half ? Raw{} : load(own, srow);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k0 + u < cnt) {                                   // warp-uniform
          if (idk[u] != id_prev) {                            // a new run starts: flush the piece
            st4(A.acc_buf + ((p0 + piece_first) * 2 + half) * D + 4 * l16, acc);
            acc = make_float4(0, 0, 0, 0);
            piece_first = k0 + u;
            id_prev = idk[u];
          }
          const float4 xv = widen(x[u]);
          if (half) {
            acc = f4_add(acc, xv);
          } else {
            const float4 t = make_float4(dmf[u] * xv.x, dmf[u] * xv.y, dmf[u] * xv.z, dmf[u] * xv.w);
            acc = f4_add(acc, f4_mul(t, w_out));
            if (WMF) dwout = f4_add(dwout, f4_mul(t, widen(sf[u])));
          }
        }
      }
    }
    st4(A.acc_buf + ((p0 + piece_first) * 2 + half) * D + 4 * l16, acc);
  }
  if (WMF) block_flush(s_red, dwout, A.dense_grad ? A.dense_grad + NCF_OFF(NCF_P_MF_OUT_W) : nullptr, lane, warp, nwarps,
                       half == 0, l16);
}

// Single-pass variant of the lean phase 1 (Adam modes, single-GPU step): a run of equal ids that lies inside the warp's
// chunk of 32 sorted positions is applied RIGHT HERE (LayerNorm backward + Adam, the arithmetic of phase 2), so its 512-byte
// sum never travels through acc_buf and its ids are not read a second time; only the pieces of runs that cross a chunk border
// (the chunk's first piece if the run came in from the previous chunk, its last piece if the run goes on) are left in acc_buf
// for phase 2, which then runs with boundary_only.  Same order of additions as the two-phase path: bit-identical tables.
// The state rows (w, m, v of both towers) of every run start are prefetched into L2 when the chunk is picked up.
template <bool WMF, bool BF>
__global__ void __launch_bounds__(EB_THREADS, 2) emb_bwd_fused_kernel(EmbBwdArgs A) {
  __shared__ float s_red[(EB_THREADS / 32) * 32 * 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = EB_THREADS / 32;
  const int half = lane >> 4, l16 = lane & 15;
  const int64_t nchunks = (A.N + EB_CHUNK - 1) / EB_CHUNK;
  const int64_t gw = (int64_t)blockIdx.x * nwarps + warp, gstride = (int64_t)gridDim.x * nwarps;
  const float4 w_out = ldg4(A.dense + NCF_OFF(NCF_P_MF_OUT_W) + 4 * l16);
  const float4 gamma = ldg4(A.dense + (half ? NCF_OFF(NCF_P_MLP_NORM_W) : NCF_OFF(NCF_P_MF_NORM_W)) + 4 * l16);
  const float* src = half ? A.d_x : A.other_y;
  const float* own = A.own_y;
  float* tw = A.w[half];
  float* tm = A.m[half];
  float* tv = A.v[half];
  float4 dwout = make_float4(0, 0, 0, 0), dgamma = dwout, dbeta = dwout;
  constexpr int U = BF ? 8 : 4;
  using Raw = typename std::conditional<BF, uint2, float4>::type;
  auto load = [&](const float* base, int64_t row) -> Raw {
    if constexpr (BF) return __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(base) + row * 64 + 4 * l16));
    else return __ldg(reinterpret_cast<const float4*>(base + row * 64 + 4 * l16));
  };
  auto widen = [](Raw q) -> float4 {
    if constexpr (BF) return make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u), __uint_as_float(q.y << 16),
                                         __uint_as_float(q.y & 0xffff0000u));
    else return q;
  };

  for (int64_t c = gw; c < nchunks; c += gstride) {
    const int64_t p0 = c * EB_CHUNK;
    const int cnt = (int)min((int64_t)EB_CHUNK, A.N - p0);
    const uint32_t my_id = lane < cnt ? A.sorted_ids[p0 + lane] : 0xffffffffu;
    const int32_t my_row = lane < cnt ? A.perm[p0 + lane] : 0;
    const float my_dmf = lane < cnt ? __ldg(A.d_mf_pred + my_row) : 0.f;
    // neighbours across the chunk borders: lane 0 looks back, lane 1 looks ahead
    uint32_t nb = 0xfffffffeu;
    if (lane == 0 && p0 > 0) nb = A.sorted_ids[p0 - 1];
    if (lane == 1 && p0 + cnt < A.N) nb = A.sorted_ids[p0 + cnt];
    const uint32_t id_first = __shfl_sync(0xffffffffu, my_id, 0), id_last = __shfl_sync(0xffffffffu, my_id, cnt - 1);
    const bool cont_in = __shfl_sync(0xffffffffu, nb, 0) == id_first;
    const bool cont_out = __shfl_sync(0xffffffffu, nb, 1) == id_last;
    {   // L2 prefetch of the state rows of every run that starts here (each row = two 128-byte lines per tower and array)
      const uint32_t prev = __shfl_up_sync(0xffffffffu, my_id, 1);
      if (lane < cnt && (lane == 0 ? !cont_in : prev != my_id)) {
        const int64_t o = (int64_t)(my_id - A.id_off) * D;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(A.w[t] + o + 32 * q));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(A.m[t] + o + 32 * q));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(A.v[t] + o + 32 * q));
          }
        }
      }
    }
    float4 acc = make_float4(0, 0, 0, 0);
    int piece_first = 0;
    uint32_t id_prev = id_first;
    // a piece [piece_first, end) of id `id` is complete: apply it (complete = the whole run lies in this chunk)
    auto finish = [&](uint32_t id, bool complete) {
      if (!complete) {
        st4(A.acc_buf + ((p0 + piece_first) * 2 + half) * D + 4 * l16, acc);
        return;
      }
      const int64_t o = (int64_t)(id - A.id_off) * D + 4 * l16;
      const float4 wrow = ld4(tw + o);
      float4 mm = ld4(tm + o), vv = ld4(tv + o);
      float rstd;
      const float4 xhat = ln_normalise(wrow, rstd);
      dgamma = f4_add(dgamma, f4_mul(acc, xhat));
      dbeta = f4_add(dbeta, acc);
      const float4 dyg = f4_mul(acc, gamma);
      const float m1 = half_warp_sum(f4_hsum(dyg)) * (1.0f / 64.0f);
      const float m2 = half_warp_sum(f4_dot(dyg, xhat)) * (1.0f / 64.0f);
      const float4 graw = make_float4(rstd * (dyg.x - m1 - xhat.x * m2), rstd * (dyg.y - m1 - xhat.y * m2),
                                      rstd * (dyg.z - m1 - xhat.z * m2), rstd * (dyg.w - m1 - xhat.w * m2));
      float4 wn = wrow;
      adam_update4(wn, mm, vv, graw, A.adam);
      st4(tw + o, wn);
      st4(tm + o, mm);
      st4(tv + o, vv);
      if (A.touched && lane == 0) A.touched[id - A.id_off] = A.touched_val;
    };
    for (int k0 = 0; k0 < cnt; k0 += U) {
      Raw x[U], sf[U];
      float dmf[U];
      uint32_t idk[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int kk = min(k0 + u, cnt - 1);
        const int64_t srow = __shfl_sync(0xffffffffu, my_row, kk);
        dmf[u] = __shfl_sync(0xffffffffu, my_dmf, kk);
        idk[u] = __shfl_sync(0xffffffffu, my_id, kk);
        x[u] = load(src, srow);
        if (WMF) sf[u] = half ? Raw{} : load(own, srow);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k0 + u < cnt) {                                   // warp-uniform
          if (idk[u] != id_prev) {                            // a new run starts: the piece before it ends inside the chunk
            finish(id_prev, !(piece_first == 0 && cont_in));
            acc = make_float4(0, 0, 0, 0);
            piece_first = k0 + u;
            id_prev = idk[u];
          }
          const float4 xv = widen(x[u]);
          if (half) {
            acc = f4_add(acc, xv);
          } else {
            const float4 t = make_float4(dmf[u] * xv.x, dmf[u] * xv.y, dmf[u] * xv.z, dmf[u] * xv.w);
            acc = f4_add(acc, f4_mul(t, w_out));
            if (WMF) dwout = f4_add(dwout, f4_mul(t, widen(sf[u])));
          }
        }
      }
    }
    finish(id_prev, !(piece_first == 0 && cont_in) && !cont_out);
  }
  float* dg = A.dense_grad;
  if (WMF) block_flush(s_red, dwout, dg ? dg + NCF_OFF(NCF_P_MF_OUT_W) : nullptr, lane, warp, nwarps, half == 0, l16);
  block_flush(s_red, dgamma, dg ? dg + NCF_OFF(NCF_P_MF_NORM_W) : nullptr, lane, warp, nwarps, half == 0, l16);
  block_flush(s_red, dgamma, dg ? dg + NCF_OFF(NCF_P_MLP_NORM_W) : nullptr, lane, warp, nwarps, half == 1, l16);
  block_flush(s_red, dbeta, dg ? dg + NCF_OFF(NCF_P_MF_NORM_B) : nullptr, lane, warp, nwarps, half == 0, l16);
  block_flush(s_red, dbeta, dg ? dg + NCF_OFF(NCF_P_MLP_NORM_B) : nullptr, lane, warp, nwarps, half == 1, l16);
}

// Phase 1 - streaming segment sum.  A warp owns a chunk of 32 sorted positions and adds up the
// upstream gradients of every run piece inside it (a piece = a run of equal ids cut at chunk borders).
// All loads depend only on the ids, so four sample rows are in flight per lane and nothing waits on a
// table row.  The sum of a piece is written to acc_buf[first sorted position of the piece].
__global__ void __launch_bounds__(EB_THREADS, 3) emb_bwd_phase1_kernel(EmbBwdArgs A) {
  __shared__ float s_red[(EB_THREADS / 32) * 32 * 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = EB_THREADS / 32;
  const int half = lane >> 4, l16 = lane & 15;
  const int64_t nchunks = (A.N + EB_CHUNK - 1) / EB_CHUNK;
  const int64_t gw = (int64_t)blockIdx.x * nwarps + warp, gstride = (int64_t)gridDim.x * nwarps;
  const bool wmf = A.accumulate_wmf != 0;

  const float4 g_mf = ldg4(A.dense + NCF_OFF(NCF_P_MF_NORM_W) + 4 * l16);
  const float4 b_mf = ldg4(A.dense + NCF_OFF(NCF_P_MF_NORM_B) + 4 * l16);
  const float4 w_out = ldg4(A.dense + NCF_OFF(NCF_P_MF_OUT_W) + 4 * l16);
  float4 dwout = make_float4(0, 0, 0, 0);

  for (int64_t c = gw; c < nchunks; c += gstride) {
    const int64_t p0 = c * EB_CHUNK;
    const int cnt = (int)min((int64_t)EB_CHUNK, A.N - p0);
    const uint32_t my_id = lane < cnt ? A.sorted_ids[p0 + lane] : 0xffffffffu;
    const int32_t my_row = lane < cnt ? A.perm[p0 + lane] : 0;
    int64_t my_other = 0;
    float my_dmf = 0.f;
    int64_t my_own = 0;
    if (lane < cnt && !A.upstream) {
      if (A.other_rows) {
        my_other = A.other_pos[my_row];
        my_own = A.own_pos[my_row];
        my_dmf = A.d_mf_pred[my_row];
      } else if (A.other_sorted) {
        my_other = A.other_sorted[p0 + lane];
        my_dmf = A.dmf_sorted[p0 + lane];
      } else {
        if (!A.other_y) my_other = clamp_id(A.other_ids[my_row], A.other_nrows);
        my_dmf = A.d_mf_pred[my_row];
      }
    }

    float4 acc = make_float4(0, 0, 0, 0);
    int piece_first = 0;
    uint32_t id_prev = __shfl_sync(0xffffffffu, my_id, 0);
    for (int k0 = 0; k0 < cnt; k0 += 4) {
      float4 x[4], sf[4];
      float dmf[4];
      uint32_t idk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int kk = min(k0 + u, cnt - 1);
        const int32_t row = __shfl_sync(0xffffffffu, my_row, kk);
        const int64_t oid = __shfl_sync(0xffffffffu, my_other, kk);
        dmf[u] = __shfl_sync(0xffffffffu, my_dmf, kk);
        idk[u] = __shfl_sync(0xffffffffu, my_id, kk);
        const bool bf = A.rows_bf16 != 0;
        if (A.upstream) x[u] = ldg4(A.upstream + (int64_t)row * 2 * D + half * D + 4 * l16);
        else if (half) x[u] = ld_row4(A.d_x, row, 4 * l16, bf);
        else if (A.other_rows) x[u] = ldg4(A.other_rows + oid * 2 * D + 4 * l16);
        else if (A.other_y) x[u] = ld_row4(A.other_y, row, 4 * l16, bf);
        else x[u] = ldg4(A.other_mf + oid * D + 4 * l16);
        sf[u] = !wmf ? make_float4(0, 0, 0, 0)                    // own MF row (d mf_output.weight): saved LN'd row or table row
                : A.own_rows ? ldg4(A.own_rows + __shfl_sync(0xffffffffu, my_own, kk) * 2 * D + 4 * l16)
                : A.own_y ? ld_row4(A.own_y, row, 4 * l16, bf)
                          : ld4(A.w[0] + (int64_t)(idk[u] - A.id_off) * D + 4 * l16);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (k0 + u < cnt) {                                   // warp-uniform
          if (idk[u] != id_prev) {                            // a new run starts: flush the piece
            st4(A.acc_buf + ((p0 + piece_first) * 2 + half) * D + 4 * l16, acc);
            acc = make_float4(0, 0, 0, 0);
            piece_first = k0 + u;
            id_prev = idk[u];
          }
          // the LayerNorms use full-warp shuffles: every lane runs them, only half 0 keeps the result
          float rs;
          float4 yo = x[u];
          const bool ready = A.other_y || A.upstream || A.other_rows;     // the rows are LayerNorm-ed already
          if (!ready) yo = affine(ln_normalise(x[u], rs), g_mf, b_mf);   // LN of the other side's MF row
          float4 ys = make_float4(0, 0, 0, 0);
          if (wmf) ys = (A.own_y || A.own_rows) ? sf[u] : affine(ln_normalise(sf[u], rs), g_mf, b_mf);
          if (half || A.upstream) {
            acc = f4_add(acc, x[u]);
          } else {
            const float4 t = make_float4(dmf[u] * yo.x, dmf[u] * yo.y, dmf[u] * yo.z, dmf[u] * yo.w);
            acc = f4_add(acc, f4_mul(t, w_out));
            dwout = f4_add(dwout, f4_mul(t, ys));
          }
        }
      }
    }
    st4(A.acc_buf + ((p0 + piece_first) * 2 + half) * D + 4 * l16, acc);
  }
  if (wmf) block_flush(s_red, dwout, A.dense_grad ? A.dense_grad + NCF_OFF(NCF_P_MF_OUT_W) : nullptr, lane, warp, nwarps,
                       half == 0, l16);
}

// Phase 2 - apply.  For every run START in the warp's chunk: add the run's pieces (its own plus the
// chunk-start pieces it continues into, fixed order), LayerNorm backward once, fused Adam.  The table
// row, its moments and the first piece are independent loads issued together.
__global__ void __launch_bounds__(EB_THREADS, 3) emb_bwd_phase2_kernel(EmbBwdArgs A) {
  __shared__ float s_red[(EB_THREADS / 32) * 32 * 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = EB_THREADS / 32;
  const int half = lane >> 4, l16 = lane & 15;
  const int64_t nchunks = (A.N + EB_CHUNK - 1) / EB_CHUNK;
  const int64_t gw = (int64_t)blockIdx.x * nwarps + warp, gstride = (int64_t)gridDim.x * nwarps;
  const float4 gamma = ldg4(A.dense + (half ? NCF_OFF(NCF_P_MLP_NORM_W) : NCF_OFF(NCF_P_MF_NORM_W)) + 4 * l16);
  float4 dgamma = make_float4(0, 0, 0, 0), dbeta = dgamma;
  const bool adam = A.mode != NCF_EMB_MATERIALIZE;

  // Chunks are handed out dynamically (one atomic per chunk): chunks that start long runs cost far more than the
  // rest, and a static assignment leaves most warps of a block waiting for one at the final reduction.
  (void)gw; (void)gstride;
  for (;;) {
    int64_t c = 0;
    if (lane == 0) c = atomicAdd(A.chunk_counter, 1);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (c >= nchunks) break;
    const int64_t p0 = c * EB_CHUNK;
    const int cnt = (int)min((int64_t)EB_CHUNK, A.N - p0);
    const uint32_t my_id = lane < cnt ? A.sorted_ids[p0 + lane] : 0xffffffffu;
    const uint32_t before = (p0 + lane > 0 && lane < cnt) ? A.sorted_ids[p0 + lane - 1] : 0xfffffffeu;
    uint32_t starts = __ballot_sync(0xffffffffu, lane < cnt && (p0 + lane == 0 || before != my_id));
    if (A.boundary_only) {      // behind emb_bwd_fused_kernel: only the chunk's last run is left, and only if it goes on
      uint32_t ahead = 0xfffffffeu;
      if (lane == 0 && p0 + cnt < A.N) ahead = A.sorted_ids[p0 + cnt];
      ahead = __shfl_sync(0xffffffffu, ahead, 0);
      if (!starts || ahead != __shfl_sync(0xffffffffu, my_id, cnt - 1)) continue;
      starts = 1u << (31 - __clz(starts));
    }
    // software pipeline over the run starts: the loads of the NEXT run are in flight while this one is applied
    struct Loads { float4 w, m, v, a; };
    auto issue = [&](int i, uint32_t id) {
      Loads L;
      const int64_t o = (int64_t)(id - A.id_off) * D + 4 * l16;
      L.m = make_float4(0, 0, 0, 0);
      L.v = L.m;
      L.w = L.m;
      const bool output_mode = A.out_rows || A.push;
      if (!output_mode) L.w = ld4(A.w[half] + o);
      if (adam && !output_mode) {
        L.m = ld4(A.m[half] + o);
        L.v = ld4(A.v[half] + o);
      }
      L.a = ld4(A.acc_buf + ((p0 + i) * 2 + half) * D + 4 * l16);
      return L;
    };
    Loads nxt{};
    int ni = -1;
    uint32_t nid = 0;
    if (starts) {
      ni = __ffs(starts) - 1;
      starts &= starts - 1;
      nid = __shfl_sync(0xffffffffu, my_id, ni);
      nxt = issue(ni, nid);
    }
    while (ni >= 0) {
      const int i = ni;
      const uint32_t id = nid;
      const Loads cur = nxt;
      if (starts) {
        ni = __ffs(starts) - 1;
        starts &= starts - 1;
        nid = __shfl_sync(0xffffffffu, my_id, ni);
        nxt = issue(ni, nid);
      } else {
        ni = -1;
      }
      const int64_t o = (int64_t)(id - A.id_off) * D + 4 * l16;
      const float4 wrow = cur.w;
      float4 mm = cur.m, vv = cur.v;
      float4 acc = cur.a;
      // does the run leave this chunk?  (its last in-chunk element is the chunk's last element)
      const uint32_t last_id = __shfl_sync(0xffffffffu, my_id, cnt - 1);
      if (last_id == id) {
        // the run may continue into the following chunks: walk their first pieces four at a time with speculative,
        // independent loads and add those that still belong to the run, in order (deterministic sum)
        int64_t cs = p0 + EB_CHUNK;
        bool more = cs < A.N;
        while (more) {
          uint32_t idn[4];
          float4 a[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int64_t pos = cs + (int64_t)u * EB_CHUNK;
            const bool ok = pos < A.N;
            idn[u] = ok ? A.sorted_ids[pos] : ~id;
            a[u] = ok ? ld4(A.acc_buf + (pos * 2 + half) * D + 4 * l16) : make_float4(0, 0, 0, 0);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            more = more && idn[u] == id;
            if (more) acc = f4_add(acc, a[u]);
          }
          cs += 4 * EB_CHUNK;
        }
      }
      if (A.push) {            // sharded requester, one-sided: the summed upstream row goes straight into its OWNER's
        // receive buffer over NVLink (128-bit stores to the peer-mapped pointer), at this rank's segment of that buffer
        const int64_t slot = A.out_slot[p0 + i];
        const int world = A.push->world;
        int o = 0;
        while (o + 1 < world && slot >= __ldg(&A.push->begin[A.push_side][o + 1])) ++o;
        const int64_t r = slot - __ldg(&A.push->begin[A.push_side][o]);
        st4(A.push->push_rows[A.push_side][o] + r * 2 * D + half * D + 4 * l16, acc);
        if (lane == 0) A.push->push_ids[A.push_side][o][r] = A.push_local[slot];
        continue;
      }
      if (A.out_rows) {        // sharded requester: the summed upstream row of this unique id goes to the exchange buffer
        st4(A.out_rows + (int64_t)A.out_slot[p0 + i] * 2 * D + half * D + 4 * l16, acc);
        continue;
      }
      float rstd;
      const float4 xhat = ln_normalise(wrow, rstd);
      dgamma = f4_add(dgamma, f4_mul(acc, xhat));
      dbeta = f4_add(dbeta, acc);
      const float4 dyg = f4_mul(acc, gamma);
      const float m1 = half_warp_sum(f4_hsum(dyg)) * (1.0f / 64.0f);
      const float m2 = half_warp_sum(f4_dot(dyg, xhat)) * (1.0f / 64.0f);
      const float4 graw = make_float4(rstd * (dyg.x - m1 - xhat.x * m2), rstd * (dyg.y - m1 - xhat.y * m2),
                                      rstd * (dyg.z - m1 - xhat.z * m2), rstd * (dyg.w - m1 - xhat.w * m2));
      if (!adam) {
        st4(A.g[half] + o, graw);
      } else {
        float4 wn = wrow;
        adam_update4(wn, mm, vv, graw, A.adam);
        st4(A.w[half] + o, wn);
        st4(A.m[half] + o, mm);
        st4(A.v[half] + o, vv);
        if (A.touched && lane == 0) A.touched[id - A.id_off] = A.touched_val;
      }
    }
  }
  float* dg = A.dense_grad;
  block_flush(s_red, dgamma, dg ? dg + NCF_OFF(NCF_P_MF_NORM_W) : nullptr, lane, warp, nwarps, half == 0, l16);
  block_flush(s_red, dgamma, dg ? dg + NCF_OFF(NCF_P_MLP_NORM_W) : nullptr, lane, warp, nwarps, half == 1, l16);
  block_flush(s_red, dbeta, dg ? dg + NCF_OFF(NCF_P_MF_NORM_B) : nullptr, lane, warp, nwarps, half == 0, l16);
  block_flush(s_red, dbeta, dg ? dg + NCF_OFF(NCF_P_MLP_NORM_B) : nullptr, lane, warp, nwarps, half == 1, l16);
}

__global__ void ids_to_keys_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t rows, uint32_t* __restrict__ keys,
                                   int32_t* __restrict__ vals) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    keys[i] = (uint32_t)clamp_id(ids[i], rows);
    vals[i] = (int32_t)i;
  }
}

// every row the step did not touch: g = wd * w (what the reference's dense Adam does, trainer.py:285)
__global__ void __launch_bounds__(256) emb_adam_sweep_kernel(float* __restrict__ w0, float* __restrict__ m0,
                                                              float* __restrict__ v0, float* __restrict__ w1,
                                                              float* __restrict__ m1, float* __restrict__ v1,
                                                              uint8_t* __restrict__ touched, int64_t rows,
                                                              AdamScalars s, bool keep_flags) {
  const int lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float* w = half ? w1 : w0;
  float* m = half ? m1 : m0;
  float* v = half ? v1 : v0;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const uint8_t t = touched ? touched[r] : 0;
    if (t) {
      __syncwarp();
      if (lane == 0 && !keep_flags) touched[r] = 0;
      continue;
    }
    const int64_t o = r * D + 4 * l16;
    float4 ww = ld4(w + o), mm = ld4(m + o), vv = ld4(v + o);
    adam_update4(ww, mm, vv, make_float4(0, 0, 0, 0), s);
    st4(w + o, ww);
    st4(m + o, mm);
    st4(v + o, vv);
  }
}

// flags of the rows a batch will update, straight from its ids (benign write races: every writer stores 1)
__global__ void mark_touched_kernel(const int64_t* __restrict__ user_ids, const int64_t* __restrict__ item_ids, int64_t n,
                                    uint8_t* __restrict__ touched_user, uint8_t* __restrict__ touched_item, int64_t rows_user,
                                    int64_t rows_item) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    touched_user[clamp_id(user_ids[i], rows_user)] = 1;
    touched_item[clamp_id(item_ids[i], rows_item)] = 1;
  }
}

// ncf_check_ids: sticky flags for ids / hours outside their range
__global__ void check_ids_kernel(const int64_t* __restrict__ user_ids, const int64_t* __restrict__ item_ids, int64_t n,
                                 int64_t rows_user, int64_t rows_item, const int64_t* __restrict__ hour,
                                 int32_t* __restrict__ status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    if (user_ids && bad_id(user_ids[i], rows_user)) flag_status(status, NCF_STATUS_BAD_USER_ID);
    if (item_ids && bad_id(item_ids[i], rows_item)) flag_status(status, NCF_STATUS_BAD_ITEM_ID);
    if (hour && bad_id(hour[i], 24)) flag_status(status, NCF_STATUS_BAD_HOUR);
  }
}

__global__ void dense_adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                                  float* __restrict__ v, int64_t n, AdamScalars s) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float ww = w[i], mm = m[i], vv = v[i];
    adam_update(ww, mm, vv, g[i], s);
    w[i] = ww;
    m[i] = mm;
    v[i] = vv;
  }
}

// TemporalEncoding.forward (architecture.py:86-94)
__global__ void temporal_fwd_kernel(const float* __restrict__ he, const float* __restrict__ de,
                                    const float* __restrict__ me, const float* __restrict__ pe,
                                    const int64_t* __restrict__ hour, const int64_t* __restrict__ day,
                                    const int64_t* __restrict__ month, const int64_t* __restrict__ days_since,
                                    int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one float4 (4 of 32 columns) per thread
  const int64_t r = i >> 3;
  const int c = (int)(i & 7) * 4;
  if (r >= n) return;
  int64_t ds = days_since[r] % 365;
  if (ds < 0) ds += 365;  // python modulo
  // nn.Embedding would raise on an index outside its table; here the index is clamped (the Python wrapper validates)
  float4 a = ldg4(he + clamp_id(hour[r], 24) * TDIM + c);
  a = f4_add(a, ldg4(de + clamp_id(day[r], 7) * TDIM + c));
  a = f4_add(a, ldg4(me + clamp_id(month[r], 12) * TDIM + c));
  a = f4_add(a, ldg4(pe + ds * TDIM + c));
  st4(out + r * TDIM + c, a);
}

// forward_simple hour tables: one block per hour (24), 256 threads
__global__ void temporal_tables_kernel(const float* __restrict__ he, const float* __restrict__ proj_w,
                                       const float* __restrict__ proj_b, const float* __restrict__ dense,
                                       float* __restrict__ tmod, float* __restrict__ tail1) {
  __shared__ float e[TDIM];
  const int h = blockIdx.x, t = threadIdx.x;
  if (t < TDIM) e[t] = he[h * TDIM + t];
  __syncthreads();
  if (t < D) {
    float s = proj_b[t];
    for (int k = 0; k < TDIM; ++k) s = fmaf(e[k], proj_w[t * TDIM + k], s);   // F.linear order: sum_k x_k W[t,k] + b
    tmod[h * D + t] = 1.0f + 0.3f * s;
  }
  if (t < H1) {
    const float* w0 = dense + NCF_OFF(NCF_P_MLP0_W) + (int64_t)t * K0 + D;
    float s = 0.f;
    for (int k = 0; k < TDIM; ++k) s = fmaf(e[k], w0[k], s);
    tail1[h * H1 + t] = s;
  }
}

__global__ void dropout_mask_kernel(DropoutRng rng, int64_t numel, uint8_t* __restrict__ keep) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < numel) keep[i] = rng.keep((uint64_t)i) ? 1 : 0;
}

}  // namespace ncf

using namespace ncf;

// ---- host entry points ------------------------------------------------------------------------
// xu / xp rows as fp32 (the C-ABI contract) or bf16 (internal: input of the tcgen05 attention block)
namespace ncf {
int gather_ln_gmf_fwd_rows(bool bf16_rows, const ncf_tables* T, const float* dense, const int64_t* user_ids,
                                     const int64_t* item_ids, int64_t N, const int64_t* hour, const float* tmod,
                                     float* mf_pred, float* xu, float* xp, float* y_item_mf, float* y_user_mf, void* stream) {
  NCF_REQUIRE(T && dense && user_ids && item_ids && mf_pred && xu && xp, "gather_ln_gmf_fwd: null argument");
  NCF_REQUIRE(N >= 0, "gather_ln_gmf_fwd: N < 0");
  NCF_REQUIRE(!hour || tmod, "gather_ln_gmf_fwd: hour needs tmod");
  if (N == 0) return NCF_OK;
  const int64_t tiles = (N + K1_TILE - 1) / K1_TILE;
  const int grid = (int)std::min<int64_t>(tiles, (int64_t)num_sms() * 6);
  cudaStream_t st = (cudaStream_t)stream;
  static const bool grouped = !(getenv("NCF_K1_GROUPED") && getenv("NCF_K1_GROUPED")[0] == '0');     // A/B switch
  if (!hour && grouped) {
    const int64_t gtiles = (N + K1G_TILE - 1) / K1G_TILE;
    const int ggrid = (int)std::min<int64_t>(gtiles, (int64_t)num_sms() * 6);
    // one group per trip at 80 registers (3 CTAs per SM): 79 us at config[2]; two groups per trip need 128 registers
    // (81 us) or spill at 80 (97 us); the per-sample kernel: 84 us
    static const int variant = getenv("NCF_K1_VARIANT") ? atoi(getenv("NCF_K1_VARIANT")) : 2;     // A/B switch
    auto kern = variant == 1 ? gather_ln_gmf_fwd_grouped_kernel<2, 2> : gather_ln_gmf_fwd_grouped_kernel<1, 3>;
    kern<<<ggrid, K1_THREADS, 0, st>>>(T->w[0], T->w[1], T->w[2], T->w[3], dense, user_ids, item_ids, N, mf_pred, xu, xp, y_item_mf,
                                       y_user_mf, bf16_rows, T->rows_user, T->rows_item, T->status);
    NCF_LAUNCH_CHECK();
    return NCF_OK;
  }
  if (hour)
    gather_ln_gmf_fwd_kernel<true><<<grid, K1_THREADS, 0, st>>>(T->w[0], T->w[1], T->w[2], T->w[3], dense, user_ids,
                                                                item_ids, N, hour, tmod, mf_pred, xu, xp, y_item_mf, y_user_mf, bf16_rows,
                                                                T->rows_user, T->rows_item, T->status);
  else
    gather_ln_gmf_fwd_kernel<false><<<grid, K1_THREADS, 0, st>>>(T->w[0], T->w[1], T->w[2], T->w[3], dense, user_ids,
                                                                 item_ids, N, nullptr, nullptr, mf_pred, xu, xp, y_item_mf, y_user_mf, bf16_rows,
                                                                 T->rows_user, T->rows_item, T->status);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}
}  // namespace ncf

extern "C" int ncf_gather_ln_gmf_fwd(const ncf_tables* T, const float* dense, const int64_t* user_ids,
                                     const int64_t* item_ids, int64_t N, const int64_t* hour, const float* tmod,
                                     float* mf_pred, float* xu, float* xp, float* y_item_mf, float* y_user_mf,
                                     void* stream) {
  return gather_ln_gmf_fwd_rows(false, T, dense, user_ids, item_ids, N, hour, tmod, mf_pred, xu, xp, y_item_mf, y_user_mf, stream);
}

extern "C" int ncf_gather_ln_gmf_fwd_bf16(const ncf_tables* T, const float* dense, const int64_t* user_ids,
                                          const int64_t* item_ids, int64_t N, const int64_t* hour, const float* tmod,
                                          float* mf_pred, void* xu_bf16, void* xp_bf16, void* y_item_mf_bf16,
                                          void* y_user_mf_bf16, void* stream) {
  return gather_ln_gmf_fwd_rows(true, T, dense, user_ids, item_ids, N, hour, tmod, mf_pred, static_cast<float*>(xu_bf16),
                                static_cast<float*>(xp_bf16), static_cast<float*>(y_item_mf_bf16),
                                static_cast<float*>(y_user_mf_bf16), stream);
}

extern "C" int ncf_gather_ln(const ncf_tables* T, const float* dense, int32_t side, const int64_t* ids, int64_t n,
                             float* mf_out, float* mlp_out, void* stream) {
  NCF_REQUIRE(T && dense && ids && (side == 0 || side == 1), "gather_ln: bad argument");
  if (n == 0) return NCF_OK;
  const int grid = (int)std::min<int64_t>((n + 7) / 8, (int64_t)num_sms() * 8);
  gather_ln_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(T->w[side], T->w[2 + side], dense, ids, n, mf_out, mlp_out,
                                                            side ? T->rows_item : T->rows_user, T->status,
                                                            side ? NCF_STATUS_BAD_ITEM_ID : NCF_STATUS_BAD_USER_ID);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

static int bits_for(int64_t rows) {
  int b = 1;
  while (b < 32 && ((int64_t)1 << b) < rows) ++b;
  return b;
}

struct EmbWs {
  uint32_t *keys_in, *keys_out;
  int32_t *vals_in, *vals_out;
  uint32_t* other_sorted;
  float* dmf_sorted;
  float *acc_buf, *acc_buf2;      // second buffer: the two sides running side by side (emb_bwd_both with a side stream)
  int32_t* counters;        // dynamic chunk schedulers of the two phase-2 launches
  void* cub_tmp;
  size_t cub_bytes;
  int64_t total;
};
// sized for the combined two-side sort (2N keys); the one-side entry points use the first half
static EmbWs carve_emb_ws(void* ws, int64_t N) {
  EmbWs w;
  Carver c(ws);
  w.keys_in = c.take<uint32_t>(2 * N);
  w.keys_out = c.take<uint32_t>(2 * N);
  w.vals_in = c.take<int32_t>(2 * N);
  w.vals_out = c.take<int32_t>(2 * N);
  w.other_sorted = c.take<uint32_t>(2 * N);
  w.dmf_sorted = c.take<float>(2 * N);
  w.acc_buf = c.take<float>(N * 2 * D);
  w.counters = c.take<int32_t>(4);
  w.cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, w.cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)std::max<int64_t>(2 * N, 1), 0, 32);
  w.cub_tmp = c.take<char>((int64_t)w.cub_bytes);
  w.acc_buf2 = c.take<float>(N * 2 * D);
  w.total = align_up(c.used, 256);
  return w;
}

extern "C" int64_t ncf_emb_bwd_workspace_bytes(int64_t N) { return carve_emb_ws(nullptr, std::max<int64_t>(N, 1)).total; }

// shared by the single-GPU backward (ids = this side's global ids, per-sample inputs) and the sharded
// owner update (ids = local row ids, `upstream` = received gradient rows)
// keys = clamped ids, one radix sort (key, position): the id-only half of run_emb_bwd.  The sharded owner runs it ahead of
// time (ncf_shard_owner_sort) on the ids the requesters write next to their pull
static int emb_sort_one(const ncf_tables* T, int32_t side, const int64_t* ids, int64_t N, void* workspace,
                        int64_t workspace_bytes, cudaStream_t st) {
  NCF_REQUIRE(side == 0 || side == 1, "emb_bwd: side must be 0 or 1");
  NCF_REQUIRE(N < ((int64_t)1 << 31), "emb_bwd: N too large");
  if (N == 0) return NCF_OK;
  EmbWs w = carve_emb_ws(workspace, N);
  if (workspace_bytes < w.total) {
    set_error("emb_bwd: workspace %lld < %lld", (long long)workspace_bytes, (long long)w.total);
    return NCF_ERR_WORKSPACE;
  }
  const int64_t rows = side ? T->rows_item : T->rows_user;
  NCF_REQUIRE(rows > 0 && rows < ((int64_t)1 << 32), "emb_bwd: table rows out of range");
  ids_to_keys_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(ids, N, rows, w.keys_in, w.vals_in);
  NCF_LAUNCH_CHECK();
  size_t tmp = w.cub_bytes;
  NCF_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tmp, w.keys_in, w.keys_out, w.vals_in, w.vals_out, (int)N, 0,
                                           bits_for(rows), st));
  return NCF_OK;
}

static int run_emb_bwd(const ncf_adam_cfg* adam, const ncf_tables* T, const float* dense, float* dense_grad,
                       int32_t side, const int64_t* ids, const int64_t* other_ids, int64_t N, const float* d_mf_pred,
                       const float* d_x, const float* other_y_mf, const float* upstream, void* workspace,
                       int64_t workspace_bytes, void* stream, bool presorted = false) {
  NCF_REQUIRE(side == 0 || side == 1, "emb_bwd: side must be 0 or 1");
  NCF_REQUIRE(N < ((int64_t)1 << 31), "emb_bwd: N too large");
  if (N == 0 || adam->emb_mode == NCF_EMB_NONE) return NCF_OK;
  EmbWs w = carve_emb_ws(workspace, N);
  if (workspace_bytes < w.total) {
    set_error("emb_bwd: workspace %lld < %lld", (long long)workspace_bytes, (long long)w.total);
    return NCF_ERR_WORKSPACE;
  }
  const int64_t rows = side ? T->rows_item : T->rows_user;
  NCF_REQUIRE(rows > 0 && rows < ((int64_t)1 << 32), "emb_bwd: table rows out of range");
  if (adam->emb_mode == NCF_EMB_MATERIALIZE)
    NCF_REQUIRE(T->g[side] && T->g[2 + side], "emb_bwd: materialize mode needs tables->g");
  else
    NCF_REQUIRE(T->m[side] && T->v[side] && T->m[2 + side] && T->v[2 + side], "emb_bwd: Adam mode needs m and v");
  cudaStream_t st = (cudaStream_t)stream;
  if (!presorted) NCF_TRY(emb_sort_one(T, side, ids, N, workspace, workspace_bytes, st));
  EmbBwdArgs A;
  A.w[0] = T->w[side];
  A.w[1] = T->w[2 + side];
  A.m[0] = T->m[side];
  A.m[1] = T->m[2 + side];
  A.v[0] = T->v[side];
  A.v[1] = T->v[2 + side];
  A.g[0] = T->g[side];
  A.g[1] = T->g[2 + side];
  A.touched = adam->emb_mode == NCF_EMB_ADAM_DENSE_EQUIV ? T->touched[side] : nullptr;
  A.touched_val = 1;
  A.other_mf = T->w[side ? 0 : 1];
  A.other_y = other_y_mf;
  A.own_y = nullptr;
  A.other_rows = A.own_rows = nullptr;
  A.other_pos = A.own_pos = nullptr;
  A.out_rows = nullptr;
  A.push = nullptr;
  A.push_local = nullptr;
  A.push_side = 0;
  A.out_slot = nullptr;
  A.upstream = upstream;
  A.other_ids = other_ids;
  A.other_nrows = side ? T->rows_user : T->rows_item;
  A.sorted_ids = w.keys_out;
  A.perm = w.vals_out;
  A.d_mf_pred = d_mf_pred;
  A.d_x = d_x;
  A.rows_bf16 = 0;
  A.boundary_only = 0;
  A.dense = dense;
  A.dense_grad = dense_grad;
  A.acc_buf = w.acc_buf;
  A.other_sorted = nullptr;
  A.dmf_sorted = nullptr;
  A.N = N;
  A.id_off = 0;
  A.mode = adam->emb_mode;
  A.accumulate_wmf = (side == 0 && !upstream) ? 1 : 0;
  A.adam = adam_scalars(*adam);
  if (adam->emb_mode == NCF_EMB_ADAM_DENSE_EQUIV) NCF_REQUIRE(A.touched, "dense-equivalent mode needs tables->touched");
  const int64_t nchunks = (N + EB_CHUNK - 1) / EB_CHUNK;
  const int wpb = EB_THREADS / 32;
  const int grid = (int)std::min<int64_t>((nchunks + wpb - 1) / wpb, (int64_t)num_sms() * 8);
  A.chunk_counter = w.counters;
  NCF_CUDA(cudaMemsetAsync(w.counters, 0, 4 * sizeof(int32_t), st));
  emb_bwd_phase1_kernel<<<grid, EB_THREADS, 0, st>>>(A);
  NCF_LAUNCH_CHECK();
  emb_bwd_phase2_kernel<<<std::min(grid, num_sms() * 3), EB_THREADS, 0, st>>>(A);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

__global__ void ids_to_keys2_kernel(const int64_t* __restrict__ user_ids, const int64_t* __restrict__ item_ids, int64_t n,
                                    uint32_t item_off, int64_t rows_item, uint32_t* __restrict__ keys,
                                    int32_t* __restrict__ vals) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    // clamped: a bad user id must not land in the item region of the combined key space (and vice versa)
    keys[i] = (uint32_t)clamp_id(user_ids[i], (int64_t)item_off);
    keys[n + i] = (uint32_t)clamp_id(item_ids[i], rows_item) + item_off;
    vals[i] = (int32_t)i;
    vals[n + i] = (int32_t)i;
  }
}

// per sorted position: the other side's id and d_mf_pred of the sample, so phase 1 reads them coalesced
__global__ void gather_sorted_kernel(const int32_t* __restrict__ perm, const int64_t* __restrict__ user_ids,
                                     const int64_t* __restrict__ item_ids, const float* __restrict__ d_mf, int64_t n,
                                     uint32_t* __restrict__ other_sorted, float* __restrict__ dmf_sorted) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < 2 * n) {
    const int32_t row = perm[p];
    other_sorted[p] = (uint32_t)(p < n ? item_ids[row] : user_ids[row]);
    dmf_sorted[p] = d_mf[row];
  }
}

namespace ncf {
// Both sides of the single-GPU backward with ONE radix sort: keys = user id | item id + rows_user, so the
// sorted array holds the users' runs in [0,N) and the items' runs in [N,2N).  Item side first (it gathers the
// user MF rows before the user side overwrites them), user side second (reads the saved item rows).
// the id-only part of emb_bwd_both: keys = user id | item id + rows_user, one radix sort.  It depends on nothing the
// forward or backward computes, so ncf_train_step may run it on an auxiliary stream next to the forward.
int emb_sort_both(const ncf_tables* T, const int64_t* user_ids, const int64_t* item_ids, int64_t N, void* workspace,
                  int64_t workspace_bytes, cudaStream_t st) {
  if (N == 0) return NCF_OK;
  NCF_REQUIRE(2 * N < ((int64_t)1 << 31), "emb_bwd: N too large");
  NCF_REQUIRE(T->rows_user + T->rows_item < ((int64_t)1 << 32), "emb_bwd: too many table rows for 32-bit keys");
  EmbWs w = carve_emb_ws(workspace, N);
  if (workspace_bytes < w.total) {
    set_error("emb_bwd: workspace %lld < %lld", (long long)workspace_bytes, (long long)w.total);
    return NCF_ERR_WORKSPACE;
  }
  ids_to_keys2_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(user_ids, item_ids, N, (uint32_t)T->rows_user, T->rows_item, w.keys_in, w.vals_in);
  NCF_LAUNCH_CHECK();
  size_t tmp = w.cub_bytes;
  NCF_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tmp, w.keys_in, w.keys_out, w.vals_in, w.vals_out, (int)(2 * N), 0,
                                           bits_for(T->rows_user + T->rows_item), st));
  return NCF_OK;
}

int emb_bwd_both(const ncf_adam_cfg* adam, const ncf_tables* T, const float* dense, float* dense_grad,
                 const int64_t* user_ids, const int64_t* item_ids, int64_t N, const float* d_mf_pred, const float* dxu,
                 const float* dxp, const float* y_item_mf, const float* y_user_mf, void* workspace, int64_t workspace_bytes,
                 cudaStream_t st, bool presorted, bool preswept, cudaStream_t side_stream, bool rows_bf16) {
  if (N == 0 || adam->emb_mode == NCF_EMB_NONE) return NCF_OK;
  NCF_REQUIRE(2 * N < ((int64_t)1 << 31), "emb_bwd: N too large");
  NCF_REQUIRE(T->rows_user + T->rows_item < ((int64_t)1 << 32), "emb_bwd: too many table rows for 32-bit keys");
  EmbWs w = carve_emb_ws(workspace, N);
  if (workspace_bytes < w.total) {
    set_error("emb_bwd: workspace %lld < %lld", (long long)workspace_bytes, (long long)w.total);
    return NCF_ERR_WORKSPACE;
  }
  if (adam->emb_mode == NCF_EMB_MATERIALIZE) {
    NCF_REQUIRE(T->g[0] && T->g[1] && T->g[2] && T->g[3], "emb_bwd: materialize mode needs tables->g");
  } else {
    for (int k = 0; k < 4; ++k) NCF_REQUIRE(T->m[k] && T->v[k], "emb_bwd: Adam mode needs m and v");
  }
  if (adam->emb_mode == NCF_EMB_ADAM_DENSE_EQUIV) NCF_REQUIRE(T->touched[0] && T->touched[1], "dense-equivalent mode needs tables->touched");
  if (!presorted) NCF_TRY(emb_sort_both(T, user_ids, item_ids, N, workspace, workspace_bytes, st));
  const bool lean = y_item_mf && y_user_mf;      // K1 saved both LayerNorm-ed MF rows: phase 1 reads everything by sample row
  if (!lean) {
    gather_sorted_kernel<<<(unsigned)((2 * N + 255) / 256), 256, 0, st>>>(w.vals_out, user_ids, item_ids, d_mf_pred, N,
                                                                          w.other_sorted, w.dmf_sorted);
    NCF_LAUNCH_CHECK();
  }
  const int64_t nchunks = (N + EB_CHUNK - 1) / EB_CHUNK;
  const int wpb = EB_THREADS / 32;
  const int grid = (int)std::min<int64_t>((nchunks + wpb - 1) / wpb, (int64_t)num_sms() * 8);
  NCF_CUDA(cudaMemsetAsync(w.counters, 0, 4 * sizeof(int32_t), st));
  // With the forward's saved rows (the lean phase 1) the two sides share nothing but the read-only inputs and the
  // atomically flushed dense gradients: given a side stream, the item side runs there next to the user side, each
  // with its own segment-sum buffer.  Phase 2 is latency-bound (long runs of popular ids), so the overlap pays.
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  static const bool two_ok = !(getenv("NCF_K6_TWO_STREAMS") && getenv("NCF_K6_TWO_STREAMS")[0] == '0');   // A/B switch
  const bool two_streams = two_ok && side_stream && side_stream != st && lean;
  if (two_streams) {
    AuxCtx* aux = aux_ctx();
    NCF_TRY(aux_events(aux));
    ev_fork = aux->ev[2];
    ev_join = aux->ev[3];
    NCF_CUDA(cudaEventRecord(ev_fork, st));
    NCF_CUDA(cudaStreamWaitEvent(side_stream, ev_fork, 0));
  }
  for (int side = 1; side >= 0; --side) {
    const cudaStream_t sst = (two_streams && side == 1) ? side_stream : st;      // this side's stream
    EmbBwdArgs A;
    A.w[0] = T->w[side];
    A.w[1] = T->w[2 + side];
    A.m[0] = T->m[side];
    A.m[1] = T->m[2 + side];
    A.v[0] = T->v[side];
    A.v[1] = T->v[2 + side];
    A.g[0] = T->g[side];
    A.g[1] = T->g[2 + side];
    A.touched = adam->emb_mode == NCF_EMB_ADAM_DENSE_EQUIV ? T->touched[side] : nullptr;
    A.touched_val = preswept ? 0 : 1;
    A.other_mf = T->w[side ? 0 : 1];
    A.other_y = side ? y_user_mf : y_item_mf;       // null: gather the other side's table row and LayerNorm it again
    A.own_y = side ? nullptr : y_user_mf;
    A.other_rows = A.own_rows = nullptr;
    A.other_pos = A.own_pos = nullptr;
    A.out_rows = nullptr;
    A.push = nullptr;
    A.push_local = nullptr;
    A.push_side = 0;
  A.push = nullptr;
  A.push_local = nullptr;
  A.push_side = 0;
    A.out_slot = nullptr;
    A.upstream = nullptr;
    A.other_ids = side ? user_ids : item_ids;
    A.other_nrows = side ? T->rows_user : T->rows_item;
    A.sorted_ids = w.keys_out + (side ? N : 0);
    A.perm = w.vals_out + (side ? N : 0);
    A.d_mf_pred = d_mf_pred;
    A.other_sorted = w.other_sorted + (side ? N : 0);
    A.dmf_sorted = w.dmf_sorted + (side ? N : 0);
    A.d_x = side ? dxp : dxu;
    A.rows_bf16 = rows_bf16 ? 1 : 0;
    A.boundary_only = 0;
    A.dense = dense;
    A.dense_grad = dense_grad;
    A.acc_buf = (two_streams && side == 1) ? w.acc_buf2 : w.acc_buf;
    A.N = N;
    A.id_off = side ? (uint32_t)T->rows_user : 0u;
    A.mode = adam->emb_mode;
    A.accumulate_wmf = side == 0 ? 1 : 0;
    A.adam = adam_scalars(*adam);
    A.chunk_counter = w.counters + side;
    // measured SLOWER than the two phases (c2 step 1.101 -> 1.169 ms, c3shard 1.213 -> 1.346 ms): every run's w / m / v loads sit
    // in series with its LayerNorm backward + Adam inside one warp, at 16 warps per SM (128 registers); opt-in A/B switch
    static const bool fused_ok = getenv("NCF_K6_FUSED") && getenv("NCF_K6_FUSED")[0] == '1';
    if (lean && fused_ok && adam->emb_mode != NCF_EMB_MATERIALIZE) {      // single pass + phase 2 for the runs that cross chunk borders
      const int fgrid = (int)std::min<int64_t>((nchunks + wpb - 1) / wpb, (int64_t)num_sms() * 2);
      if (A.accumulate_wmf) {
        if (rows_bf16) emb_bwd_fused_kernel<true, true><<<fgrid, EB_THREADS, 0, sst>>>(A);
        else emb_bwd_fused_kernel<true, false><<<fgrid, EB_THREADS, 0, sst>>>(A);
      } else {
        if (rows_bf16) emb_bwd_fused_kernel<false, true><<<fgrid, EB_THREADS, 0, sst>>>(A);
        else emb_bwd_fused_kernel<false, false><<<fgrid, EB_THREADS, 0, sst>>>(A);
      }
      A.boundary_only = 1;
    } else if (lean) {      // two phases
      if (A.accumulate_wmf) {
        if (rows_bf16) emb_bwd_phase1_lean_kernel<true, true><<<grid, EB_THREADS, 0, sst>>>(A);
        else emb_bwd_phase1_lean_kernel<true, false><<<grid, EB_THREADS, 0, sst>>>(A);
      } else {
        if (rows_bf16) emb_bwd_phase1_lean_kernel<false, true><<<grid, EB_THREADS, 0, sst>>>(A);
        else emb_bwd_phase1_lean_kernel<false, false><<<grid, EB_THREADS, 0, sst>>>(A);
      }
    } else {
      emb_bwd_phase1_kernel<<<grid, EB_THREADS, 0, sst>>>(A);
    }
    NCF_LAUNCH_CHECK();
    emb_bwd_phase2_kernel<<<std::min(grid, num_sms() * 3), EB_THREADS, 0, sst>>>(A);     // resident blocks pull chunks
    NCF_LAUNCH_CHECK();
  }
  if (two_streams) {
    NCF_CUDA(cudaEventRecord(ev_join, side_stream));
    NCF_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
  }
  return NCF_OK;
}
}  // namespace ncf

extern "C" int ncf_emb_bwd_adam_both(const ncf_adam_cfg* adam, const ncf_tables* T, const float* dense, float* dense_grad,
                                     const int64_t* user_ids, const int64_t* item_ids, int64_t N, const float* d_mf_pred,
                                     const float* d_xu, const float* d_xp, const float* y_item_mf, const float* y_user_mf,
                                     void* workspace, int64_t workspace_bytes, void* stream) {
  NCF_REQUIRE(adam && T && dense && user_ids && item_ids && d_mf_pred && d_xu && d_xp && y_item_mf && workspace,
              "emb_bwd_adam_both: null argument");
  return emb_bwd_both(adam, T, dense, dense_grad, user_ids, item_ids, N, d_mf_pred, d_xu, d_xp, y_item_mf, y_user_mf,
                      workspace, workspace_bytes, (cudaStream_t)stream, false, false, ncf::aux_ctx()->stream, false);
}

extern "C" int ncf_emb_bwd_adam_both_bf16(const ncf_adam_cfg* adam, const ncf_tables* T, const float* dense, float* dense_grad,
                                          const int64_t* user_ids, const int64_t* item_ids, int64_t N, const float* d_mf_pred,
                                          const void* d_xu_bf16, const void* d_xp_bf16, const void* y_item_mf_bf16,
                                          const void* y_user_mf_bf16, void* workspace, int64_t workspace_bytes, void* stream) {
  NCF_REQUIRE(adam && T && dense && user_ids && item_ids && d_mf_pred && d_xu_bf16 && d_xp_bf16 && y_item_mf_bf16 &&
                  y_user_mf_bf16 && workspace, "emb_bwd_adam_both_bf16: null argument");
  return emb_bwd_both(adam, T, dense, dense_grad, user_ids, item_ids, N, d_mf_pred, static_cast<const float*>(d_xu_bf16),
                      static_cast<const float*>(d_xp_bf16), static_cast<const float*>(y_item_mf_bf16),
                      static_cast<const float*>(y_user_mf_bf16), workspace, workspace_bytes, (cudaStream_t)stream, false, false,
                      ncf::aux_ctx()->stream, true);
}

extern "C" int ncf_emb_bwd_adam(const ncf_adam_cfg* adam, const ncf_tables* T, const float* dense, float* dense_grad,
                                int32_t side, const int64_t* user_ids, const int64_t* item_ids, int64_t N,
                                const float* d_mf_pred, const float* d_x, const float* other_y_mf, void* workspace,
                                int64_t workspace_bytes, void* stream) {
  NCF_REQUIRE(adam && T && dense && user_ids && item_ids && d_mf_pred && d_x, "emb_bwd_adam: null argument");
  return run_emb_bwd(adam, T, dense, dense_grad, side, side ? item_ids : user_ids, side ? user_ids : item_ids, N,
                     d_mf_pred, d_x, other_y_mf, nullptr, workspace, workspace_bytes, stream);
}

// =============================================================================================
// Sharded requester: full de-duplication of the ids a rank asks for.  ONE radix sort of both sides
// (sorted ids are owner-major already: owner = id / block is monotonic in the id): every distinct id is exchanged
// once, pos[n] points each sample at its row, and the backward sums the per-sample upstream gradients per distinct
// id with the K6 segment-sum kernels before they travel (deterministic order, no atomics on the rows).
// =============================================================================================
namespace ncf {
struct RouteWs {
  uint32_t *keys_in, *keys_out;
  int32_t *vals_in, *vals_out, *slot, *incl;   // keys_out / vals_out / slot stay valid for the backward
  void* cub_tmp;
  size_t cub_bytes;
  int64_t total;
};
static RouteWs carve_route_ws(void* ws, int64_t N) {
  RouteWs w;
  Carver c(ws);
  w.keys_out = c.take<uint32_t>(2 * N);
  w.vals_out = c.take<int32_t>(2 * N);
  w.slot = c.take<int32_t>(2 * N);
  w.incl = c.take<int32_t>(2 * N);
  w.keys_in = c.take<uint32_t>(2 * N);
  w.vals_in = c.take<int32_t>(2 * N);
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)std::max<int64_t>(2 * N, 1), 0, 32);
  cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (int32_t*)nullptr, (int32_t*)nullptr, (int)std::max<int64_t>(2 * N, 1));
  w.cub_bytes = std::max(sort_bytes, scan_bytes);
  w.cub_tmp = c.take<char>((int64_t)w.cub_bytes);
  w.total = align_up(c.used, 256);
  return w;
}

__global__ void route_flags_kernel(const uint32_t* __restrict__ skeys, int64_t N, int32_t* __restrict__ flag) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < 2 * N) flag[p] = (p == 0 || p == N || skeys[p] != skeys[p - 1]) ? 1 : 0;
}
// incl = inclusive sum of the head flags over both sides; slot = index of the distinct id inside its side
__global__ void route_finish_kernel(const uint32_t* __restrict__ skeys, const int32_t* __restrict__ perm,
                                    const int32_t* __restrict__ incl, int64_t N, uint32_t item_off, int64_t block_u,
                                    int64_t block_i, int world, int32_t* __restrict__ slot,
                                    unsigned long long* __restrict__ counts, int64_t* __restrict__ local_ids,
                                    int64_t* __restrict__ pos) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = p < 2 * N;
  int owner = -1;                   // counter index of a run head (side * world + owner rank), -1 for everyone else
  if (valid) {
    const int side = p >= N;
    const int32_t s = incl[p] - 1 - (side ? __ldg(incl + N - 1) : 0);
    slot[p] = s;
    pos[(int64_t)side * N + perm[p]] = s;
    const bool head = p == 0 || p == N || skeys[p] != skeys[p - 1];
    if (head) {
      const int64_t id = (int64_t)skeys[p] - (side ? item_off : 0u);
      const int64_t block = side ? block_i : block_u;
      local_ids[(int64_t)side * N + s] = id % block;
      owner = side * world + (int)(id / block);
    }
  }
  // Sorted keys are owner-major: the heads of a warp name one or two owners.  One atomic per owner and warp instead of one
  // per distinct id - 100M x 10M ids are almost all distinct, and 655k atomics on 2 * world words took 236 us (12 us now).
  const unsigned peers = __match_any_sync(0xffffffffu, owner);
  if (owner >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(counts + owner, (unsigned long long)__popc(peers));
}

int64_t shard_route_ws_bytes(int64_t N) { return carve_route_ws(nullptr, std::max<int64_t>(N, 1)).total; }

int shard_route(const int64_t* user_ids, const int64_t* item_ids, int64_t N, int64_t rows_user, int64_t rows_item, int32_t world,
                int64_t* counts, int64_t* local_ids, int64_t* pos, void* route_ws, int64_t route_ws_bytes, cudaStream_t st) {
  NCF_REQUIRE(user_ids && item_ids && counts && local_ids && pos && route_ws, "shard_route: null argument");
  NCF_REQUIRE(world >= 1 && world <= 1024 && rows_user >= 1 && rows_item >= 1, "shard_route: bad world / rows");
  NCF_REQUIRE(N >= 0 && 2 * N < ((int64_t)1 << 31), "shard_route: bad N");
  NCF_REQUIRE(rows_user + rows_item < ((int64_t)1 << 32), "shard_route: too many table rows for 32-bit keys");
  NCF_CUDA(cudaMemsetAsync(counts, 0, sizeof(int64_t) * 2 * world, st));
  if (N == 0) return NCF_OK;
  RouteWs w = carve_route_ws(route_ws, N);
  if (route_ws_bytes < w.total) {
    set_error("shard_route: workspace %lld < %lld", (long long)route_ws_bytes, (long long)w.total);
    return NCF_ERR_WORKSPACE;
  }
  ids_to_keys2_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(user_ids, item_ids, N, (uint32_t)rows_user, rows_item, w.keys_in, w.vals_in);
  NCF_LAUNCH_CHECK();
  size_t tmp = w.cub_bytes;
  NCF_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tmp, w.keys_in, w.keys_out, w.vals_in, w.vals_out, (int)(2 * N), 0,
                                           bits_for(rows_user + rows_item), st));
  const unsigned grid = (unsigned)((2 * N + 255) / 256);
  route_flags_kernel<<<grid, 256, 0, st>>>(w.keys_out, N, w.slot);
  NCF_LAUNCH_CHECK();
  tmp = w.cub_bytes;
  NCF_CUDA(cub::DeviceScan::InclusiveSum(w.cub_tmp, tmp, w.slot, w.incl, (int)(2 * N), st));
  route_finish_kernel<<<grid, 256, 0, st>>>(w.keys_out, w.vals_out, w.incl, N, (uint32_t)rows_user, (rows_user + world - 1) / world,
                                            (rows_item + world - 1) / world, world, w.slot,
                                            reinterpret_cast<unsigned long long*>(counts), local_ids, pos);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

// gu[slot] / gi[slot] = sum over the samples of each distinct id of [ d_mf * w_mf * y_other_mf | d_x ]; also adds
// d mf_output.weight.  route_ws = the buffer shard_route filled for the same batch.
int shard_requester_grads(const float* dense, float* dense_grad, const float* rows_u, const float* rows_i, const int64_t* pos_u,
                          const int64_t* pos_i, int64_t N, const float* d_mf, const float* dxu, const float* dxp,
                          const void* route_ws, float* gu, float* gi, void* emb_ws, int64_t emb_ws_bytes, cudaStream_t st,
                          const ncf_shard_plan* plan, const int64_t* local_ids, bool rows_bf16, const float* y_item_mf,
                          const float* y_user_mf) {
  if (N == 0) return NCF_OK;
  // the forward kept both LayerNorm-ed MF rows per sample (ncf_shard_forward): phase 1 reads everything by sample row
  static const bool lean_ok = !(getenv("NCF_SHARD_LEAN") && getenv("NCF_SHARD_LEAN")[0] == '0');     // A/B switch
  const bool lean = lean_ok && y_item_mf && y_user_mf;
  RouteWs r = carve_route_ws(const_cast<void*>(route_ws), N);
  EmbWs w = carve_emb_ws(emb_ws, N);
  if (emb_ws_bytes < w.total) {
    set_error("shard_requester_grads: workspace %lld < %lld", (long long)emb_ws_bytes, (long long)w.total);
    return NCF_ERR_WORKSPACE;
  }
  NCF_CUDA(cudaMemsetAsync(w.counters, 0, 4 * sizeof(int32_t), st));
  const int64_t nchunks = (N + EB_CHUNK - 1) / EB_CHUNK;
  const int wpb = EB_THREADS / 32;
  const int grid = (int)std::min<int64_t>((nchunks + wpb - 1) / wpb, (int64_t)num_sms() * 8);
  // the two sides share read-only inputs only (own segment-sum buffer each): with an auxiliary stream (ncf_set_aux_stream)
  // the item side runs next to the user side, like the single-GPU embedding backward
  AuxCtx* aux = aux_ctx();
  const bool two_streams = aux->stream && aux->stream != st;
  if (two_streams) {
    NCF_TRY(aux_events(aux));
    NCF_CUDA(cudaEventRecord(aux->ev[2], st));
    NCF_CUDA(cudaStreamWaitEvent(aux->stream, aux->ev[2], 0));
  }
  for (int side = 0; side < 2; ++side) {
    const cudaStream_t sst = (two_streams && side == 1) ? aux->stream : st;
    EmbBwdArgs A{};
    A.sorted_ids = r.keys_out + (side ? N : 0);
    A.perm = r.vals_out + (side ? N : 0);
    A.d_mf_pred = d_mf;
    A.d_x = side ? dxp : dxu;
    A.rows_bf16 = rows_bf16 ? 1 : 0;
    A.other_rows = side ? rows_u : rows_i;
    A.other_pos = side ? pos_u : pos_i;
    A.own_rows = side ? rows_i : rows_u;
    A.own_pos = side ? pos_i : pos_u;
    A.out_rows = side ? gi : gu;
    A.out_slot = r.slot + (side ? N : 0);
    A.push = plan;
    A.push_local = local_ids ? local_ids + (side ? N : 0) : nullptr;
    A.push_side = side;
    A.dense = dense;
    A.dense_grad = dense_grad;
    A.acc_buf = (two_streams && side == 1) ? w.acc_buf2 : w.acc_buf;
    A.chunk_counter = w.counters + side;
    A.N = N;
    A.mode = NCF_EMB_ADAM_SPARSE;
    A.accumulate_wmf = side == 0 ? 1 : 0;
    if (lean) {
      A.other_y = side ? y_user_mf : y_item_mf;
      A.own_y = side ? nullptr : y_user_mf;
      if (A.accumulate_wmf) {
        if (rows_bf16) emb_bwd_phase1_lean_kernel<true, true><<<grid, EB_THREADS, 0, sst>>>(A);
        else emb_bwd_phase1_lean_kernel<true, false><<<grid, EB_THREADS, 0, sst>>>(A);
      } else {
        if (rows_bf16) emb_bwd_phase1_lean_kernel<false, true><<<grid, EB_THREADS, 0, sst>>>(A);
        else emb_bwd_phase1_lean_kernel<false, false><<<grid, EB_THREADS, 0, sst>>>(A);
      }
    } else {
      emb_bwd_phase1_kernel<<<grid, EB_THREADS, 0, sst>>>(A);
    }
    NCF_LAUNCH_CHECK();
    A.dense_grad = nullptr;                  // output mode: no LayerNorm-affine gradients here (the owners add them)
    emb_bwd_phase2_kernel<<<std::min(grid, num_sms() * 3), EB_THREADS, 0, sst>>>(A);
    NCF_LAUNCH_CHECK();
  }
  if (two_streams) {
    NCF_CUDA(cudaEventRecord(aux->ev[3], aux->stream));
    NCF_CUDA(cudaStreamWaitEvent(st, aux->ev[3], 0));
  }
  return NCF_OK;
}
}  // namespace ncf

extern "C" int64_t ncf_shard_route_workspace_bytes(int64_t N) { return ncf::shard_route_ws_bytes(N); }
extern "C" int ncf_shard_route(const int64_t* user_ids, const int64_t* item_ids, int64_t N, int64_t rows_user, int64_t rows_item,
                               int32_t world, int64_t* counts, int64_t* local_ids, int64_t* pos, void* route_ws,
                               int64_t route_ws_bytes, void* stream) {
  return ncf::shard_route(user_ids, item_ids, N, rows_user, rows_item, world, counts, local_ids, pos, route_ws, route_ws_bytes,
                          (cudaStream_t)stream);
}

extern "C" int ncf_shard_owner_sort(const ncf_tables* T, int32_t side, const int64_t* local_ids, int64_t n, void* workspace,
                                    int64_t workspace_bytes, void* stream) {
  NCF_REQUIRE(T && n >= 0, "shard_owner_sort: bad argument");
  if (n == 0) return NCF_OK;
  NCF_REQUIRE(local_ids && workspace, "shard_owner_sort: null buffer");
  return emb_sort_one(T, side, local_ids, n, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int ncf_shard_owner_update_sorted(const ncf_adam_cfg* adam, const ncf_tables* T, const float* dense,
                                             float* dense_grad, int32_t side, const int64_t* local_ids, int64_t n,
                                             const float* grad_rows, void* workspace, int64_t workspace_bytes, void* stream) {
  NCF_REQUIRE(adam && T && dense && n >= 0, "shard_owner_update_sorted: bad argument");
  if (n == 0) return NCF_OK;
  NCF_REQUIRE(local_ids && grad_rows && workspace, "shard_owner_update_sorted: null buffer");
  return run_emb_bwd(adam, T, dense, dense_grad, side, local_ids, nullptr, n, nullptr, nullptr, nullptr, grad_rows,
                     workspace, workspace_bytes, stream, true);
}

extern "C" int ncf_shard_owner_update(const ncf_adam_cfg* adam, const ncf_tables* T, const float* dense,
                                      float* dense_grad, int32_t side, const int64_t* local_ids, int64_t n,
                                      const float* grad_rows, void* workspace, int64_t workspace_bytes, void* stream) {
  NCF_REQUIRE(adam && T && dense && n >= 0, "shard_owner_update: bad argument");
  if (n == 0) return NCF_OK;
  NCF_REQUIRE(local_ids && grad_rows && workspace, "shard_owner_update: null buffer");
  return run_emb_bwd(adam, T, dense, dense_grad, side, local_ids, nullptr, n, nullptr, nullptr, nullptr, grad_rows,
                     workspace, workspace_bytes, stream);
}

namespace ncf {
static int launch_sweep(const ncf_adam_cfg* adam, const ncf_tables* T, bool keep_flags, cudaStream_t st) {
  const AdamScalars s = adam_scalars(*adam);
  for (int side = 0; side < 2; ++side) {
    const int64_t rows = side ? T->rows_item : T->rows_user;
    if (rows == 0) continue;
    NCF_REQUIRE(T->m[side] && T->v[side] && T->m[2 + side] && T->v[2 + side], "emb_adam_sweep: needs m and v");
    const int grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)num_sms() * 8);
    emb_adam_sweep_kernel<<<grid, 256, 0, st>>>(T->w[side], T->m[side], T->v[side], T->w[2 + side], T->m[2 + side],
                                                T->v[2 + side], T->touched[side], rows, s, keep_flags);
    NCF_LAUNCH_CHECK();
  }
  return NCF_OK;
}

// The sweep depends on WHICH rows a batch updates, not on anything the step computes, and it writes only rows the step
// neither reads nor writes: flags are set from the ids, the sweep leaves them alone, and emb_bwd_both(preswept) clears
// the flag of every row it updates.  Same per-row arithmetic as sweep-after-K6, so the tables come out bit-identical.
int emb_sweep_early(const ncf_adam_cfg* adam, const ncf_tables* T, const int64_t* user_ids, const int64_t* item_ids, int64_t N,
                    cudaStream_t st) {
  NCF_REQUIRE(T->touched[0] && T->touched[1], "dense-equivalent mode needs tables->touched");
  if (N > 0) {
    mark_touched_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(user_ids, item_ids, N, T->touched[0], T->touched[1], T->rows_user,
                                                                     T->rows_item);
    NCF_LAUNCH_CHECK();
  }
  return launch_sweep(adam, T, true, st);
}
}  // namespace ncf

extern "C" int ncf_emb_adam_sweep(const ncf_adam_cfg* adam, const ncf_tables* T, void* stream) {
  NCF_REQUIRE(adam && T, "emb_adam_sweep: null argument");
  return ncf::launch_sweep(adam, T, false, (cudaStream_t)stream);
}

extern "C" int ncf_dense_adam(float* w, const float* g, float* m, float* v, int64_t n, const ncf_adam_cfg* adam,
                              void* stream) {
  NCF_REQUIRE(w && g && m && v && adam && n >= 0, "dense_adam: bad argument");
  if (n == 0) return NCF_OK;
  dense_adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, g, m, v, n, adam_scalars(*adam));
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

extern "C" int ncf_check_ids(const int64_t* user_ids, const int64_t* item_ids, int64_t N, int64_t rows_user, int64_t rows_item,
                             const int64_t* hour, int32_t* status, void* stream) {
  NCF_REQUIRE(status && N >= 0 && (user_ids || item_ids || hour), "check_ids: bad argument");
  if (N == 0) return NCF_OK;
  check_ids_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(user_ids, item_ids, N, rows_user, rows_item,
                                                                                  hour, status);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

extern "C" int ncf_temporal_fwd(const float* he, const float* de, const float* me, const float* pe,
                                const int64_t* hour, const int64_t* day, const int64_t* month,
                                const int64_t* days_since, int64_t n, float* out, void* stream) {
  NCF_REQUIRE(he && de && me && pe && hour && day && month && days_since && out, "temporal_fwd: null argument");
  if (n == 0) return NCF_OK;
  const int64_t threads = n * 8;
  temporal_fwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(he, de, me, pe, hour, day,
                                                                                          month, days_since, n, out);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

extern "C" int ncf_temporal_tables(const float* he, const float* proj_w, const float* proj_b, const float* dense,
                                   float* tmod, float* tail1, void* stream) {
  NCF_REQUIRE(he && proj_w && proj_b && dense && tmod && tail1, "temporal_tables: null argument");
  temporal_tables_kernel<<<24, 256, 0, (cudaStream_t)stream>>>(he, proj_w, proj_b, dense, tmod, tail1);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

extern "C" int ncf_dropout_mask(const ncf_run_cfg* cfg, int32_t site, int64_t numel, uint8_t* keep, void* stream) {
  NCF_REQUIRE(cfg && keep && site >= 0 && site < 4 && numel >= 0, "dropout_mask: bad argument");
  if (numel == 0) return NCF_OK;
  dropout_mask_kernel<<<(unsigned)((numel + 255) / 256), 256, 0, (cudaStream_t)stream>>>(make_rng(*cfg, site), numel, keep);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}
