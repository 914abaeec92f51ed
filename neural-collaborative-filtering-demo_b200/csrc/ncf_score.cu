// Full-catalogue scoring + deterministic top-k (app.py:43-77), the eval-mode item fold behind it,
// and the row-wise shard bucketizer.
//
// Eval mode attends over ONE key, so softmax == 1 and the MLP tower depends on the item only
// (architecture.py:275-276, 315-323):  logit(u,i) = LN_mf(U_mf[u]) . P_hat[i] + g[i].
// ncf_item_fold computes P_hat/g once per catalogue with the regular tower kernels; ncf_score_topk
// streams P_hat against tiles of users and keeps, per user, the k best (score desc, index asc)
// in shared memory: candidates that beat the current k-th best are appended to a buffer that is
// bitonic-merged into the running list when it fills up.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "ncf_tower.cuh"

namespace ncf {

__global__ void iota_kernel(int64_t* p, int64_t n, int64_t start) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = start + i;
}

__global__ void item_fold_kernel(const float* __restrict__ ymf, const float* __restrict__ mlp_pred,
                                 const float* __restrict__ dense, float* __restrict__ p_hat, float* __restrict__ g,
                                 int64_t I) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one float4 per thread
  const int64_t i = t >> 4;
  const int c = (int)(t & 15) * 4;
  if (i >= I) return;
  const float a = __ldg(dense + NCF_OFF(NCF_P_FINAL_W));
  const float4 w = ldg4(dense + NCF_OFF(NCF_P_MF_OUT_W) + c);
  const float4 y = ld4(ymf + i * D + c);
  st4(p_hat + i * D + c, make_float4(a * y.x * w.x, a * y.y * w.y, a * y.z * w.z, a * y.w * w.w));
  if (c == 0) {
    const float cc = __ldg(dense + NCF_OFF(NCF_P_FINAL_W) + 1);
    g[i] = fmaf(a, __ldg(dense + NCF_OFF(NCF_P_MF_OUT_B)), fmaf(cc, mlp_pred[i], __ldg(dense + NCF_OFF(NCF_P_FINAL_B))));
  }
}

// ---- top-k -------------------------------------------------------------------------------------
// Register-tiled fp32 scoring: a CTA owns 32 users and streams 128-item tiles of P_hat through shared
// memory (cp.async, double buffered).  Warp w holds users 4w..4w+3, lane l items l, l+32, l+64, l+96 of the
// tile: 16 accumulators per thread, every shared-memory word feeds 4 FMAs.  Each dot product is summed in
// ascending-k order with fmaf, so a score does not depend on the tiling.  Scores above the user's
// running k-th best are appended to that user's candidate buffer (shared memory), which is
// bitonic-merged into the running list when it could overflow.
constexpr int TK_THREADS = 256;
constexpr int TK_UT = 32;       // users per CTA
constexpr int TK_IT = 128;      // items per tile
constexpr int TK_LD = 68;       // padded row length (floats): 16-byte aligned, conflict-free float4 reads
constexpr int TK_KMAX = 128;    // running list length (k <= 128)
constexpr int TK_BUF = 512;     // list + candidate buffer, power of two for the bitonic network
constexpr uint32_t TKS_U = 0;
constexpr uint32_t TKS_P = TKS_U + TK_UT * TK_LD * 4;
constexpr uint32_t TKS_G = TKS_P + 2 * TK_IT * TK_LD * 4;
constexpr uint32_t TKS_BUF = TKS_G + 2 * TK_IT * 4;
constexpr uint32_t TKS_CNT = TKS_BUF + TK_UT * TK_BUF * 8;
constexpr uint32_t TKS_THR = TKS_CNT + TK_UT * 4;
constexpr uint32_t TKS_LTHR = TKS_THR + TK_UT * 8;      // per-user logit pre-filter (float)
constexpr uint32_t TKS_TOTAL = TKS_LTHR + TK_UT * 4;

// larger key == better: fp32 score bits (scores are positive) then lower item index
__device__ __forceinline__ unsigned long long tk_key(float score, uint32_t idx) {
  return ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(0xffffffffu - idx);
}

// descending bitonic sort of n (power of two) keys in shared memory by the whole CTA
__device__ void bitonic_sort_desc(unsigned long long* a, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long x = a[i], y = a[l];
          const bool desc = (i & k) == 0;
          if (desc ? x < y : x > y) {
            a[i] = y;
            a[l] = x;
          }
        }
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src, uint32_t src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(TK_THREADS, 1) score_topk_kernel(const float* __restrict__ t_umf,
                                                                   const float* __restrict__ dense,
                                                                   const float* __restrict__ p_hat, const float* __restrict__ g,
                                                                   const int64_t* __restrict__ user_ids, int64_t n_users,
                                                                   int64_t I, int nsplit,
                                                                   unsigned long long* __restrict__ part /*[n_users][nsplit][KMAX]*/,
                                                                   int64_t rows_user, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* s_u = reinterpret_cast<float*>(smem + TKS_U);
  float* s_p = reinterpret_cast<float*>(smem + TKS_P);
  float* s_g = reinterpret_cast<float*>(smem + TKS_G);
  unsigned long long* s_buf = reinterpret_cast<unsigned long long*>(smem + TKS_BUF);
  int* s_cnt = reinterpret_cast<int*>(smem + TKS_CNT);
  unsigned long long* s_thr = reinterpret_cast<unsigned long long*>(smem + TKS_THR);
  float* s_lthr = reinterpret_cast<float*>(smem + TKS_LTHR);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t u0 = (int64_t)blockIdx.x * TK_UT;
  const int split = blockIdx.y;
  const int nu = (int)min((int64_t)TK_UT, n_users - u0);

  // LN_mf of the tile's user rows: one warp per user (4 rounds), two-pass variance like nn.LayerNorm
  for (int uu = warp; uu < TK_UT; uu += TK_THREADS / 32) {
    float x0 = 0.f, x1 = 0.f;
    if (uu < nu) {
      const int64_t uid = user_ids ? user_ids[u0 + uu] : u0 + uu;
      if (bad_id(uid, rows_user)) flag_status(status, NCF_STATUS_BAD_USER_ID);
      const float* row = t_umf + clamp_id(uid, rows_user) * D;
      x0 = row[lane];
      x1 = row[lane + 32];
    }
    if (!dense) {      // raw query rows (ncf_dot_topk): no LayerNorm
      s_u[uu * TK_LD + lane] = x0;
      s_u[uu * TK_LD + lane + 32] = x1;
      continue;
    }
    const float mean = warp_sum(x0 + x1) * (1.0f / D);
    const float d0 = x0 - mean, d1 = x1 - mean;
    const float rstd = rsqrtf(warp_sum(d0 * d0 + d1 * d1) * (1.0f / D) + LN_EPS);
    s_u[uu * TK_LD + lane] = fmaf(d0 * rstd, __ldg(dense + NCF_OFF(NCF_P_MF_NORM_W) + lane), __ldg(dense + NCF_OFF(NCF_P_MF_NORM_B) + lane));
    s_u[uu * TK_LD + lane + 32] =
        fmaf(d1 * rstd, __ldg(dense + NCF_OFF(NCF_P_MF_NORM_W) + lane + 32), __ldg(dense + NCF_OFF(NCF_P_MF_NORM_B) + lane + 32));
  }
  for (int i = tid; i < TK_UT * TK_BUF; i += TK_THREADS) s_buf[i] = 0ull;
  if (tid < TK_UT) {
    s_cnt[tid] = TK_KMAX;   // slots [0,KMAX) hold the running list (zeros = empty)
    s_thr[tid] = 0ull;
    s_lthr[tid] = -INFINITY;
  }

  const int64_t per = (I + nsplit - 1) / nsplit;
  const int64_t i_begin = split * per, i_end = min(I, i_begin + per);
  const int64_t ntile = i_begin < i_end ? (i_end - i_begin + TK_IT - 1) / TK_IT : 0;

  auto stage = [&](int64_t t, int buf) {      // cp.async one 128 x 64 tile of P_hat (+ g) into buffer `buf`
    const int64_t base = i_begin + t * TK_IT;
    float* dst = s_p + buf * TK_IT * TK_LD;
    for (int c = tid; c < TK_IT * 16; c += TK_THREADS) {
      const int r = c >> 4, q = c & 15;
      const int64_t i = base + r;
      const bool ok = i < i_end;
      cp_async16(dst + r * TK_LD + 4 * q, p_hat + (ok ? i : i_begin) * D + 4 * q, ok ? 16u : 0u);
    }
    if (tid < TK_IT) {
      const int64_t i = base + tid;
      s_g[buf * TK_IT + tid] = i < i_end ? __ldg(g + i) : 0.f;
    }
    cp_async_commit();
  };

  if (ntile > 0) stage(0, 0);
  for (int64_t t = 0; t < ntile; ++t) {
    const int buf = (int)(t & 1);
    if (t + 1 < ntile) {
      stage(t + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* pt = s_p + buf * TK_IT * TK_LD;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 4
    for (int k4 = 0; k4 < D / 4; ++k4) {
      float4 uv[4], pv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) uv[a] = *reinterpret_cast<const float4*>(s_u + (warp * 4 + a) * TK_LD + 4 * k4);
#pragma unroll
      for (int b = 0; b < 4; ++b) pv[b] = *reinterpret_cast<const float4*>(pt + (lane + 32 * b) * TK_LD + 4 * k4);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          acc[a][b] = fmaf(uv[a].x, pv[b].x, acc[a][b]);
          acc[a][b] = fmaf(uv[a].y, pv[b].y, acc[a][b]);
          acc[a][b] = fmaf(uv[a].z, pv[b].z, acc[a][b]);
          acc[a][b] = fmaf(uv[a].w, pv[b].w, acc[a][b]);
        }
    }
    const int64_t base = i_begin + t * TK_IT;
    // sigmoid is monotone: a logit below the user's pre-filter (the logit of the running k-th best score minus
    // a rounding margin) cannot enter the list, so the exact score / key is only formed for the few survivors
    float lthr[4];
    unsigned long long kthr[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      lthr[a] = s_lthr[warp * 4 + a];
      kthr[a] = s_thr[warp * 4 + a];
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t i = base + lane + 32 * b;
      if (i < i_end) {
        const float gi = s_g[buf * TK_IT + lane + 32 * b];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int uu = warp * 4 + a;
          const float z = acc[a][b] + gi;
          if (uu < nu && z >= lthr[a]) {
            const float sc = 1.0f / (1.0f + expf(-z));
            const unsigned long long key = tk_key(sc, (uint32_t)i);
            if (key > kthr[a]) {
              const int pos = atomicAdd(&s_cnt[uu], 1);
              s_buf[uu * TK_BUF + pos] = key;
            }
          }
        }
      }
    }
    __syncthreads();
    // merge any list whose buffer might overflow during the next tile (<= 128 new candidates per user)
    for (int uu = 0; uu < nu; ++uu) {
      if (s_cnt[uu] > TK_BUF - TK_IT) {   // uniform: s_cnt read after the barrier
        bitonic_sort_desc(s_buf + uu * TK_BUF, TK_BUF);
        for (int k = TK_KMAX + tid; k < TK_BUF; k += TK_THREADS) s_buf[uu * TK_BUF + k] = 0ull;
        if (tid == 0) {
          s_cnt[uu] = TK_KMAX;
          const unsigned long long kth = s_buf[uu * TK_BUF + TK_KMAX - 1];
          s_thr[uu] = kth;
          const float sk = __uint_as_float((uint32_t)(kth >> 32));
          float lt = -INFINITY;
          if (kth != 0ull && sk > 0.f && sk < 1.f) {
            lt = logf(sk / (1.0f - sk));
            lt -= 1e-4f + 1e-4f * fabsf(lt);
          }
          s_lthr[uu] = lt;
        }
        __syncthreads();
      }
    }
  }
  __syncthreads();
  for (int uu = 0; uu < nu; ++uu) {
    bitonic_sort_desc(s_buf + uu * TK_BUF, TK_BUF);
    unsigned long long* dst = part + ((u0 + uu) * nsplit + split) * TK_KMAX;
    for (int k = tid; k < TK_KMAX; k += TK_THREADS) dst[k] = s_buf[uu * TK_BUF + k];
  }
}

// Small catalogues (below ~1M items): 4 users per CTA, one P_hat row per thread - less work per CTA, so the grid
// fills the machine and the per-user merges do not dominate.
constexpr int TKS_UT = 4;
__global__ void __launch_bounds__(TK_THREADS) score_topk_small_kernel(const float* __restrict__ t_umf,
                                                                const float* __restrict__ dense,
                                                                const float* __restrict__ p_hat, const float* __restrict__ g,
                                                                const int64_t* __restrict__ user_ids, int64_t n_users,
                                                                int64_t I, int nsplit,
                                                                unsigned long long* __restrict__ part /*[n_users][nsplit][KMAX]*/,
                                                                int64_t rows_user, int32_t* __restrict__ status) {
  __shared__ __align__(16) float s_u[TKS_UT][D];
  __shared__ unsigned long long s_buf[TKS_UT][TK_BUF];
  __shared__ int s_cnt[TKS_UT];
  __shared__ unsigned long long s_thr[TKS_UT];
  const int64_t u0 = (int64_t)blockIdx.x * TKS_UT;
  const int split = blockIdx.y;
  const int nu = (int)min((int64_t)TKS_UT, n_users - u0);

  // LN_mf of the tile's user rows (64 threads per user; two-pass variance like nn.LayerNorm)
  {
    const int uu = threadIdx.x >> 6, c = threadIdx.x & 63;
    __shared__ float s_tmp[TKS_UT][D];
    float x = 0.f;
    if (uu < nu) {
      const int64_t uid = user_ids ? user_ids[u0 + uu] : u0 + uu;
      if (bad_id(uid, rows_user)) flag_status(status, NCF_STATUS_BAD_USER_ID);
      x = t_umf[clamp_id(uid, rows_user) * D + c];
    }
    s_tmp[uu][c] = x;
    __syncthreads();
    if (!dense) {      // raw query rows (ncf_dot_topk): no LayerNorm
      s_u[uu][c] = x;
    } else {
      float mean = 0.f;
      for (int k = 0; k < D; ++k) mean += s_tmp[uu][k];
      mean *= (1.0f / D);
      float var = 0.f;
      for (int k = 0; k < D; ++k) var = fmaf(s_tmp[uu][k] - mean, s_tmp[uu][k] - mean, var);
      const float rstd = rsqrtf(var * (1.0f / D) + LN_EPS);
      s_u[uu][c] = fmaf((x - mean) * rstd, __ldg(dense + NCF_OFF(NCF_P_MF_NORM_W) + c), __ldg(dense + NCF_OFF(NCF_P_MF_NORM_B) + c));
    }
  }
  for (int i = threadIdx.x; i < TKS_UT * TK_BUF; i += TK_THREADS) (&s_buf[0][0])[i] = 0ull;
  if (threadIdx.x < TKS_UT) {
    s_cnt[threadIdx.x] = TK_KMAX;   // slots [0,KMAX) hold the running list (zeros = empty)
    s_thr[threadIdx.x] = 0ull;
  }
  __syncthreads();

  const int64_t per = (I + nsplit - 1) / nsplit;
  const int64_t i_begin = split * per, i_end = min(I, i_begin + per);
  for (int64_t base = i_begin; base < i_end; base += TK_THREADS) {
    const int64_t i = base + threadIdx.x;
    if (i < i_end) {
      float acc[TKS_UT];
#pragma unroll
      for (int uu = 0; uu < TKS_UT; ++uu) acc[uu] = 0.f;
      const float* row = p_hat + i * D;
#pragma unroll
      for (int q = 0; q < D / 4; ++q) {
        const float4 p = ldg4(row + 4 * q);
#pragma unroll
        for (int uu = 0; uu < TKS_UT; ++uu) {
          const float4 uv = *reinterpret_cast<const float4*>(&s_u[uu][4 * q]);
          acc[uu] = fmaf(uv.x, p.x, acc[uu]);
          acc[uu] = fmaf(uv.y, p.y, acc[uu]);
          acc[uu] = fmaf(uv.z, p.z, acc[uu]);
          acc[uu] = fmaf(uv.w, p.w, acc[uu]);
        }
      }
      const float gi = __ldg(g + i);
#pragma unroll
      for (int uu = 0; uu < TKS_UT; ++uu) {
        if (uu < nu) {
          const float s = 1.0f / (1.0f + expf(-(acc[uu] + gi)));
          const unsigned long long key = tk_key(s, (uint32_t)i);
          if (key > s_thr[uu]) {
            const int pos = atomicAdd(&s_cnt[uu], 1);
            s_buf[uu][pos] = key;
          }
        }
      }
    }
    __syncthreads();
    // merge any list whose buffer might overflow during the next step
    for (int uu = 0; uu < nu; ++uu) {
      if (s_cnt[uu] > TK_BUF - TK_THREADS) {   // uniform: s_cnt read after the barrier
        bitonic_sort_desc(s_buf[uu], TK_BUF);
        for (int k = TK_KMAX + threadIdx.x; k < TK_BUF; k += TK_THREADS) s_buf[uu][k] = 0ull;
        if (threadIdx.x == 0) {
          s_cnt[uu] = TK_KMAX;
          s_thr[uu] = s_buf[uu][TK_KMAX - 1];
        }
        __syncthreads();
      }
    }
  }
  for (int uu = 0; uu < nu; ++uu) {
    bitonic_sort_desc(s_buf[uu], TK_BUF);
    unsigned long long* dst = part + ((u0 + uu) * nsplit + split) * TK_KMAX;
    for (int k = threadIdx.x; k < TK_KMAX; k += TK_THREADS) dst[k] = s_buf[uu][k];
  }
}

// merge the per-split lists of one user and emit (index, score)
constexpr int TM_MAX = 4096;
__global__ void __launch_bounds__(256) topk_merge_kernel(const unsigned long long* __restrict__ part, int nsplit, int k,
                                                         int64_t* __restrict__ idx_out, float* __restrict__ score_out) {
  __shared__ unsigned long long s[TM_MAX];
  const int64_t u = blockIdx.x;
  const int n = nsplit * TK_KMAX;
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) s[i] = i < n ? part[u * n + i] : 0ull;
  bitonic_sort_desc(s, n2);
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const unsigned long long key = s[i];
    idx_out[u * k + i] = key == 0ull ? -1 : (int64_t)(0xffffffffu - (uint32_t)(key & 0xffffffffull));
    score_out[u * k + i] = __uint_as_float((uint32_t)(key >> 32));
  }
}

int launch_topk_merge(const unsigned long long* part, int nsplit, int k, int64_t n_users, int64_t* idx, float* score,
                      cudaStream_t st) {
  topk_merge_kernel<<<(unsigned)n_users, 256, 0, st>>>(part, nsplit, k, idx, score);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

// ---- shard bucketize ---------------------------------------------------------------------------
__global__ void owner_keys_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t block, uint32_t* __restrict__ keys,
                                  int32_t* __restrict__ vals) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    keys[i] = (uint32_t)(ids[i] / block);
    vals[i] = (int32_t)i;
  }
}
__global__ void bucket_finish_kernel(const int64_t* __restrict__ ids, const uint32_t* __restrict__ skeys,
                                     const int32_t* __restrict__ svals, int64_t n, int64_t block, int world,
                                     int64_t* __restrict__ counts, int64_t* __restrict__ order,
                                     int64_t* __restrict__ local_ids) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int64_t src = svals[i];
    order[i] = src;
    local_ids[i] = ids[src] % block;
  }
  if (i < world) {   // counts[w] = upper_bound(w) - lower_bound(w) in the sorted owner keys
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (skeys[mid] < (uint32_t)i) lo = mid + 1; else hi = mid; }
    const int64_t lb = lo;
    hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (skeys[mid] <= (uint32_t)i) lo = mid + 1; else hi = mid; }
    counts[i] = lo - lb;
  }
}

// ---- bucketize with adjacent-run compression (ncf_shard_bucketize_runs) --------------------------------
// Training batches repeat the user id over the S rows of an interaction (data_prep.py:286-303): equal ADJACENT
// ids are served by one exchanged row.  A sample is a run head when its id differs from its predecessor's;
// heads are bucketed by owner (stable), the other samples go to an extra bucket `world` that is never sent.
__global__ void run_keys_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t block, uint32_t world,
                                uint32_t* __restrict__ keys, int32_t* __restrict__ vals, int32_t* __restrict__ head_idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const bool head = i == 0 || ids[i] != ids[i - 1];
    keys[i] = head ? (uint32_t)(ids[i] / block) : world;
    vals[i] = (int32_t)i;
    head_idx[i] = head ? (int32_t)i : 0;      // inclusive max-scan -> index of the run head of every sample
  }
}
__global__ void run_finish_kernel(const int64_t* __restrict__ ids, const uint32_t* __restrict__ skeys,
                                  const int32_t* __restrict__ svals, int64_t n, int64_t block, int world,
                                  int64_t* __restrict__ counts, int64_t* __restrict__ local_ids,
                                  int32_t* __restrict__ slot_of_head) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && skeys[i] < (uint32_t)world) {
    const int64_t src = svals[i];
    local_ids[i] = ids[src] % block;
    slot_of_head[src] = (int32_t)i;
  }
  if (i < world) {   // counts[w] = upper_bound(w) - lower_bound(w) in the sorted keys
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (skeys[mid] < (uint32_t)i) lo = mid + 1; else hi = mid; }
    const int64_t lb = lo;
    hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (skeys[mid] <= (uint32_t)i) lo = mid + 1; else hi = mid; }
    counts[i] = lo - lb;
  }
}
__global__ void run_pos_kernel(const int32_t* __restrict__ head_idx, const int32_t* __restrict__ slot_of_head, int64_t n,
                               int64_t* __restrict__ pos) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pos[i] = slot_of_head[head_idx[i]];
}

}  // namespace ncf

using namespace ncf;

static ncf_run_cfg eval_cfg() {
  ncf_run_cfg c{};
  c.S = 1;
  c.training = 0;
  c.precision = NCF_FP32;
  return c;
}

constexpr int64_t FOLD_CHUNK = 1 << 20;      // items folded per pass (bounds the tower workspace)

struct FoldWs {
  int64_t* ids;
  float* ymf;
  float* out;
  void* tower;
  int64_t tower_bytes, total;
};
static FoldWs carve_fold(void* ws, int64_t chunk) {
  FoldWs f;
  Carver c(ws);
  f.ids = c.take<int64_t>(chunk);
  f.ymf = c.take<float>(chunk * D);
  f.out = c.take<float>(chunk);
  f.tower_bytes = carve_tower_ws(nullptr, chunk, eval_cfg()).total;
  f.tower = c.take<char>(f.tower_bytes);
  f.total = align_up(c.used, 256);
  return f;
}

extern "C" int64_t ncf_item_fold_workspace_bytes(int64_t I) {
  return carve_fold(nullptr, std::min<int64_t>(std::max<int64_t>(I, 1), FOLD_CHUNK)).total;
}

extern "C" int ncf_item_fold(const ncf_tables* T, const float* dense, float* p_hat, float* g, void* workspace,
                             int64_t workspace_bytes, void* stream) {
  NCF_REQUIRE(T && dense && p_hat && g && workspace, "item_fold: null argument");
  const int64_t I = T->rows_item;
  NCF_REQUIRE(I > 0 && I < ((int64_t)1 << 32) - 1, "item_fold: bad item count");
  const int64_t chunk = std::min<int64_t>(I, FOLD_CHUNK);
  FoldWs f = carve_fold(workspace, chunk);
  if (workspace_bytes < f.total) {
    set_error("item_fold: workspace %lld < %lld", (long long)workspace_bytes, (long long)f.total);
    return NCF_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const ncf_run_cfg cfg = eval_cfg();
  for (int64_t start = 0; start < I; start += chunk) {
    const int64_t n = std::min<int64_t>(chunk, I - start);
    TowerWs w = carve_tower_ws(f.tower, n, cfg);
    iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(f.ids, n, start);
    NCF_LAUNCH_CHECK();
    NCF_TRY(ncf_gather_ln(T, dense, 1, f.ids, n, f.ymf, w.xp, stream));
    NCF_CUDA(cudaMemsetAsync(w.mf_pred, 0, sizeof(float) * n, st));
    NCF_TRY(tower_f32_forward(cfg, dense, n, nullptr, nullptr, f.out, w, st));
    item_fold_kernel<<<(unsigned)((n * 16 + 255) / 256), 256, 0, st>>>(f.ymf, w.mlp_pred, dense, p_hat + start * D, g + start, n);
    NCF_LAUNCH_CHECK();
  }
  return NCF_OK;
}

constexpr int64_t TK_SMALL_CATALOGUE = 1 << 20;
static int topk_splits(int64_t n_users, int64_t I) {
  if (I < TK_SMALL_CATALOGUE) {
    const int64_t tiles = (n_users + TKS_UT - 1) / TKS_UT;
    int64_t want = std::max<int64_t>(1, (2 * (int64_t)num_sms() + tiles - 1) / tiles);
    want = std::min<int64_t>(std::min<int64_t>(want, std::max<int64_t>(1, I / (4 * TK_THREADS))), TM_MAX / TK_KMAX);
    return (int)want;
  }
  const int64_t tiles = (n_users + TK_UT - 1) / TK_UT;
  int64_t want = std::max<int64_t>(1, (2 * (int64_t)num_sms() + tiles - 1) / tiles);
  const int64_t max_by_items = std::max<int64_t>(1, I / (8 * TK_IT));
  want = std::min<int64_t>(std::min<int64_t>(want, max_by_items), TM_MAX / TK_KMAX);
  return (int)want;
}

extern "C" int64_t ncf_score_topk_workspace_bytes(int64_t n_users, int64_t I, int32_t k) {
  (void)k;
  return align_up(std::max<int64_t>(n_users, 1) * topk_splits(n_users, I) * TK_KMAX * 8, 256);
}

namespace ncf {
int score_topk_impl(const ncf_tables* T, const float* dense, const float* p_hat, const float* g, const int64_t* user_ids,
                    int64_t n_users, int64_t I, int32_t k, int64_t* topk_idx, float* topk_score, void* workspace,
                    int64_t workspace_bytes, void* stream);
}
extern "C" int ncf_score_topk(const ncf_tables* T, const float* dense, const float* p_hat, const float* g,
                              const int64_t* user_ids, int64_t n_users, int64_t I, int32_t k, int64_t* topk_idx,
                              float* topk_score, void* workspace, int64_t workspace_bytes, void* stream) {
  NCF_REQUIRE(dense && user_ids, "score_topk: null argument");
  return score_topk_impl(T, dense, p_hat, g, user_ids, n_users, I, k, topk_idx, topk_score, workspace, workspace_bytes, stream);
}
// dense == NULL: the "user" rows are raw query vectors (no LayerNorm); user_ids == NULL: rows 0 .. n_users-1
int ncf::score_topk_impl(const ncf_tables* T, const float* dense, const float* p_hat, const float* g, const int64_t* user_ids,
                         int64_t n_users, int64_t I, int32_t k, int64_t* topk_idx, float* topk_score, void* workspace,
                         int64_t workspace_bytes, void* stream) {
  NCF_REQUIRE(T && p_hat && g && topk_idx && topk_score && workspace, "score_topk: null argument");
  NCF_REQUIRE(k >= 1 && k <= TK_KMAX, "score_topk: k=%d outside [1,%d]", k, TK_KMAX);
  NCF_REQUIRE(I >= 1 && I < ((int64_t)1 << 32) - 1, "score_topk: bad catalogue size");
  if (n_users == 0) return NCF_OK;
  const int nsplit = topk_splits(n_users, I);
  const int64_t need = ncf_score_topk_workspace_bytes(n_users, I, k);
  if (workspace_bytes < need) {
    set_error("score_topk: workspace %lld < %lld", (long long)workspace_bytes, (long long)need);
    return NCF_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* part = static_cast<unsigned long long*>(workspace);
  if (I < TK_SMALL_CATALOGUE) {
    const int64_t tiles_s = (n_users + TKS_UT - 1) / TKS_UT;
    NCF_REQUIRE(tiles_s < ((int64_t)1 << 31), "score_topk: too many users in one call");
    dim3 grid_s((unsigned)tiles_s, nsplit);
    score_topk_small_kernel<<<grid_s, TK_THREADS, 0, st>>>(T->w[0], dense, p_hat, g, user_ids, n_users, I, nsplit, part, T->rows_user, T->status);
    NCF_LAUNCH_CHECK();
    topk_merge_kernel<<<(unsigned)n_users, 256, 0, st>>>(part, nsplit, k, topk_idx, topk_score);
    NCF_LAUNCH_CHECK();
    return NCF_OK;
  }
  const int64_t tiles = (n_users + TK_UT - 1) / TK_UT;
  NCF_REQUIRE(tiles < ((int64_t)1 << 31), "score_topk: too many users in one call");
  dim3 grid((unsigned)tiles, nsplit);
  static bool configured = false;
  if (!configured) {
    NCF_CUDA(cudaFuncSetAttribute(score_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TKS_TOTAL));
    configured = true;
  }
  score_topk_kernel<<<grid, TK_THREADS, TKS_TOTAL, st>>>(T->w[0], dense, p_hat, g, user_ids, n_users, I, nsplit, part, T->rows_user, T->status);
  NCF_LAUNCH_CHECK();
  topk_merge_kernel<<<(unsigned)n_users, 256, 0, st>>>(part, nsplit, k, topk_idx, topk_score);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

struct BucketWs {
  uint32_t *keys_in, *keys_out;
  int32_t *vals_in, *vals_out;
  void* cub_tmp;
  size_t cub_bytes;
  int64_t total;
};
static BucketWs carve_bucket(void* ws, int64_t n) {
  BucketWs b;
  Carver c(ws);
  b.keys_in = c.take<uint32_t>(n);
  b.keys_out = c.take<uint32_t>(n);
  b.vals_in = c.take<int32_t>(n);
  b.vals_out = c.take<int32_t>(n);
  b.cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b.cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)std::max<int64_t>(n, 1), 0, 32);
  b.cub_tmp = c.take<char>((int64_t)b.cub_bytes);
  b.total = align_up(c.used, 256);
  return b;
}

extern "C" int64_t ncf_shard_bucketize_workspace_bytes(int64_t n, int32_t world) {
  (void)world;
  return carve_bucket(nullptr, std::max<int64_t>(n, 1)).total;
}

extern "C" int ncf_shard_bucketize(const int64_t* ids, int64_t n, int64_t rows, int32_t world, int64_t* counts,
                                   int64_t* order, int64_t* local_ids, void* workspace, int64_t workspace_bytes,
                                   void* stream) {
  NCF_REQUIRE(ids && counts && order && local_ids && workspace, "shard_bucketize: null argument");
  NCF_REQUIRE(world >= 1 && world <= 1024 && rows >= 1, "shard_bucketize: bad world/rows");
  NCF_REQUIRE(n >= 0 && n < ((int64_t)1 << 31), "shard_bucketize: bad n");
  cudaStream_t st = (cudaStream_t)stream;
  NCF_CUDA(cudaMemsetAsync(counts, 0, sizeof(int64_t) * world, st));
  if (n == 0) return NCF_OK;
  BucketWs b = carve_bucket(workspace, n);
  if (workspace_bytes < b.total) {
    set_error("shard_bucketize: workspace %lld < %lld", (long long)workspace_bytes, (long long)b.total);
    return NCF_ERR_WORKSPACE;
  }
  const int64_t block = (rows + world - 1) / world;
  owner_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ids, n, block, b.keys_in, b.vals_in);
  NCF_LAUNCH_CHECK();
  int bits = 1;
  while ((1 << bits) < world) ++bits;
  size_t tmp = b.cub_bytes;
  NCF_CUDA(cub::DeviceRadixSort::SortPairs(b.cub_tmp, tmp, b.keys_in, b.keys_out, b.vals_in, b.vals_out, (int)n, 0, bits, st));
  const int64_t threads = std::max<int64_t>(n, world);
  bucket_finish_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(ids, b.keys_out, b.vals_out, n, block, world,
                                                                           counts, order, local_ids);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

struct RunBucketWs {
  uint32_t *keys_in, *keys_out;
  int32_t *vals_in, *vals_out, *head_idx, *slot_of_head;
  void* cub_tmp;
  size_t cub_bytes;
  int64_t total;
};
static RunBucketWs carve_run_bucket(void* ws, int64_t n) {
  RunBucketWs b;
  Carver c(ws);
  b.keys_in = c.take<uint32_t>(n);
  b.keys_out = c.take<uint32_t>(n);
  b.vals_in = c.take<int32_t>(n);
  b.vals_out = c.take<int32_t>(n);
  b.head_idx = c.take<int32_t>(n);
  b.slot_of_head = c.take<int32_t>(n);
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)std::max<int64_t>(n, 1), 0, 32);
  cub::DeviceScan::InclusiveScan(nullptr, scan_bytes, (int32_t*)nullptr, (int32_t*)nullptr, cub::Max(), (int)std::max<int64_t>(n, 1));
  b.cub_bytes = std::max(sort_bytes, scan_bytes);
  b.cub_tmp = c.take<char>((int64_t)b.cub_bytes);
  b.total = align_up(c.used, 256);
  return b;
}

extern "C" int64_t ncf_shard_bucketize_runs_workspace_bytes(int64_t n, int32_t world) {
  (void)world;
  return carve_run_bucket(nullptr, std::max<int64_t>(n, 1)).total;
}

extern "C" int ncf_shard_bucketize_runs(const int64_t* ids, int64_t n, int64_t rows, int32_t world, int64_t* counts,
                                        int64_t* local_ids, int64_t* pos, void* workspace, int64_t workspace_bytes,
                                        void* stream) {
  NCF_REQUIRE(ids && counts && local_ids && pos && workspace, "shard_bucketize_runs: null argument");
  NCF_REQUIRE(world >= 1 && world <= 1024 && rows >= 1, "shard_bucketize_runs: bad world/rows");
  NCF_REQUIRE(n >= 0 && n < ((int64_t)1 << 31), "shard_bucketize_runs: bad n");
  cudaStream_t st = (cudaStream_t)stream;
  NCF_CUDA(cudaMemsetAsync(counts, 0, sizeof(int64_t) * world, st));
  if (n == 0) return NCF_OK;
  RunBucketWs b = carve_run_bucket(workspace, n);
  if (workspace_bytes < b.total) {
    set_error("shard_bucketize_runs: workspace %lld < %lld", (long long)workspace_bytes, (long long)b.total);
    return NCF_ERR_WORKSPACE;
  }
  const int64_t block = (rows + world - 1) / world;
  const unsigned grid = (unsigned)((std::max<int64_t>(n, world) + 255) / 256);
  run_keys_kernel<<<grid, 256, 0, st>>>(ids, n, block, (uint32_t)world, b.keys_in, b.vals_in, b.head_idx);
  NCF_LAUNCH_CHECK();
  size_t tmp = b.cub_bytes;
  NCF_CUDA(cub::DeviceScan::InclusiveScan(b.cub_tmp, tmp, b.head_idx, b.head_idx, cub::Max(), (int)n, st));
  int bits = 1;
  while ((1 << bits) < world + 1) ++bits;
  tmp = b.cub_bytes;
  NCF_CUDA(cub::DeviceRadixSort::SortPairs(b.cub_tmp, tmp, b.keys_in, b.keys_out, b.vals_in, b.vals_out, (int)n, 0, bits, st));
  run_finish_kernel<<<grid, 256, 0, st>>>(ids, b.keys_out, b.vals_out, n, block, world, counts, local_ids, b.slot_of_head);
  NCF_LAUNCH_CHECK();
  run_pos_kernel<<<grid, 256, 0, st>>>(b.head_idx, b.slot_of_head, n, pos);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}
