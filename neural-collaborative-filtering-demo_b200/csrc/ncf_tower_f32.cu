// fp32 (CUDA-core) dense towers of AdvancedNCF: q/k/v/out projections, the per-interaction
// multi-head attention, the 3-layer MLP (Linear -> ReLU -> LayerNorm -> Dropout), the output head,
// BCELoss, and the backward of all of them.  This is the exact-arithmetic path (<= 1e-5 relative vs
// the reference, BASELINE config 1); the tcgen05 bf16 path lives in ncf_tower_tc.cu.
//
// Reference: architecture.py:18-57 (MultiHeadAttention), :230-252 (mlp, heads), :311-354 (forward),
// trainer.py:78,271 (BCELoss).
#include <cstdlib>

#include "ncf_tower.cuh"

namespace ncf {

// =============================================================================================
// Y[M,J] = X[M,R] . B[R,J] (+ bias) with an optional ReLU -> LayerNorm -> Dropout epilogue.
//   TRANS == false : B[k][j] = W[j*ldw + k]   (nn.Linear forward, W is [J, >=R] row-major)
//   TRANS == true  : B[k][j] = W[k*ldw + j]   (input gradient dX = dY . W)
// 8 warps x 8 rows per CTA; a warp owns whole rows (LayerNorm = warp shuffles), a lane owns J/32
// columns in groups of VW contiguous columns.
// =============================================================================================
constexpr int LG_THREADS = 256;
constexpr int LG_RM = 8;                 // rows per warp
constexpr int LG_BM = 8 * LG_RM;         // rows per CTA
constexpr int LG_KC = 32;                // reduction chunk staged in shared memory

enum { EPI_NONE = 0, EPI_RELU_LN_DROP = 1 };

template <int J>
struct ColMap {
  static constexpr int CN = J / 32;                 // columns per lane
  static constexpr int VW = CN >= 4 ? 4 : CN;       // contiguous vector width
  static constexpr int NG = CN / VW;                // groups per lane
  __device__ static __forceinline__ int col(int lane, int g) { return g * (32 * VW) + lane * VW; }
};

struct LinearArgs {
  const float* X; int64_t ldx;     // [M, R]
  const float* W; int64_t ldw;
  const float* bias;               // [J] or null
  float* Y; int64_t ldy;           // [M, J]
  int64_t M; int R;
  // epilogue (EPI_RELU_LN_DROP)
  const float* gamma; const float* beta;
  float* Rout;                     // [M, J] relu output saved for backward (may be null)
  const int64_t* hour; const float* tail1;   // optional per-row additive table [24, J]
  DropoutRng rng;
};

template <int J, bool TRANS, int EPI>
__global__ void __launch_bounds__(LG_THREADS) linear_kernel(LinearArgs A) {
  using CM = ColMap<J>;
  __shared__ __align__(16) float Bs[LG_KC][J];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * LG_BM + warp * LG_RM;

  float acc[LG_RM][CM::CN];
#pragma unroll
  for (int r = 0; r < LG_RM; ++r)
#pragma unroll
    for (int c = 0; c < CM::CN; ++c) acc[r][c] = 0.f;

  const float* xrow[LG_RM];
#pragma unroll
  for (int r = 0; r < LG_RM; ++r) xrow[r] = A.X + min(row0 + r, A.M - 1) * A.ldx;

  for (int k0 = 0; k0 < A.R; k0 += LG_KC) {
    __syncthreads();
    if (TRANS) {
      for (int i = threadIdx.x; i < LG_KC * (J / 4); i += LG_THREADS) {
        const int k = i / (J / 4), j4 = i % (J / 4);
        *reinterpret_cast<float4*>(&Bs[k][4 * j4]) = ldg4(A.W + (int64_t)(k0 + k) * A.ldw + 4 * j4);
      }
    } else {
      for (int i = threadIdx.x; i < (LG_KC / 4) * J; i += LG_THREADS) {
        const int j = i % J, kq = i / J;
        const float4 v = ldg4(A.W + (int64_t)j * A.ldw + k0 + 4 * kq);
        Bs[4 * kq + 0][j] = v.x; Bs[4 * kq + 1][j] = v.y; Bs[4 * kq + 2][j] = v.z; Bs[4 * kq + 3][j] = v.w;
      }
    }
    __syncthreads();
#pragma unroll 2
    for (int k4 = 0; k4 < LG_KC; k4 += 4) {
      float4 xv[LG_RM];
#pragma unroll
      for (int r = 0; r < LG_RM; ++r) xv[r] = ldg4(xrow[r] + k0 + k4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float b[CM::CN];
#pragma unroll
        for (int g = 0; g < CM::NG; ++g) {
          const float* p = &Bs[k4 + kk][CM::col(lane, g)];
          if (CM::VW == 4) {
            const float4 t = *reinterpret_cast<const float4*>(p);
            b[g * 4 + 0] = t.x; b[g * 4 + 1] = t.y; b[g * 4 + 2] = t.z; b[g * 4 + 3] = t.w;
          } else {
            const float2 t = *reinterpret_cast<const float2*>(p);
            b[0] = t.x; b[1] = t.y;
          }
        }
#pragma unroll
        for (int r = 0; r < LG_RM; ++r) {
          const float x = kk == 0 ? xv[r].x : kk == 1 ? xv[r].y : kk == 2 ? xv[r].z : xv[r].w;
#pragma unroll
          for (int c = 0; c < CM::CN; ++c) acc[r][c] = fmaf(x, b[c], acc[r][c]);
        }
      }
    }
  }

  // ---- epilogue -------------------------------------------------------------------------------
  float bias[CM::CN], gam[CM::CN], bet[CM::CN];
#pragma unroll
  for (int g = 0; g < CM::NG; ++g)
#pragma unroll
    for (int v = 0; v < CM::VW; ++v) {
      const int c = CM::col(lane, g) + v;
      bias[g * CM::VW + v] = A.bias ? __ldg(A.bias + c) : 0.f;
      if (EPI == EPI_RELU_LN_DROP) {
        gam[g * CM::VW + v] = __ldg(A.gamma + c);
        bet[g * CM::VW + v] = __ldg(A.beta + c);
      }
    }
#pragma unroll
  for (int r = 0; r < LG_RM; ++r) {
    const int64_t row = row0 + r;
    const bool live = row < A.M;   // warp-uniform
    float y[CM::CN];
#pragma unroll
    for (int c = 0; c < CM::CN; ++c) y[c] = acc[r][c] + bias[c];
    if (EPI == EPI_RELU_LN_DROP) {
      if (A.hour && live) {
        const float* t = A.tail1 + clamp_id(A.hour[row], 24) * J;
#pragma unroll
        for (int g = 0; g < CM::NG; ++g)
#pragma unroll
          for (int v = 0; v < CM::VW; ++v) y[g * CM::VW + v] += __ldg(t + CM::col(lane, g) + v);
      }
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < CM::CN; ++c) {
        y[c] = fmaxf(y[c], 0.f);
        s += y[c];
      }
      if (A.Rout && live) {
#pragma unroll
        for (int g = 0; g < CM::NG; ++g) {
          float* p = A.Rout + row * J + CM::col(lane, g);
          if (CM::VW == 4) st4(p, make_float4(y[g * 4], y[g * 4 + 1], y[g * 4 + 2], y[g * 4 + 3]));
          else *reinterpret_cast<float2*>(p) = make_float2(y[0], y[1]);
        }
      }
      const float mean = warp_sum(s) * (1.0f / J);
      float q = 0.f;
#pragma unroll
      for (int c = 0; c < CM::CN; ++c) {
        const float d = y[c] - mean;
        q = fmaf(d, d, q);
      }
      const float rstd = rsqrtf(warp_sum(q) * (1.0f / J) + LN_EPS);
#pragma unroll
      for (int c = 0; c < CM::CN; ++c) y[c] = fmaf((y[c] - mean) * rstd, gam[c], bet[c]);
      if (A.rng.thresh != 0u) {
#pragma unroll
        for (int g = 0; g < CM::NG; ++g) {
          const uint64_t e0 = (uint64_t)row * J + CM::col(lane, g);
          if (CM::VW == 4) A.rng.apply4(e0, &y[g * 4]);
          else A.rng.apply2(e0, &y[0]);
        }
      }
    }
    if (live) {
#pragma unroll
      for (int g = 0; g < CM::NG; ++g) {
        float* p = A.Y + row * A.ldy + CM::col(lane, g);
        if (CM::VW == 4) st4(p, make_float4(y[g * 4], y[g * 4 + 1], y[g * 4 + 2], y[g * 4 + 3]));
        else *reinterpret_cast<float2*>(p) = make_float2(y[0], y[1]);
      }
    }
  }
}

template <int J, bool TRANS, int EPI>
static int launch_linear(const LinearArgs& A, cudaStream_t st) {
  if (A.M == 0) return NCF_OK;
  if (A.R % LG_KC != 0) { set_error("linear: reduction length %d not a multiple of %d", A.R, LG_KC); return NCF_ERR_ARG; }
  const unsigned grid = (unsigned)((A.M + LG_BM - 1) / LG_BM);
  linear_kernel<J, TRANS, EPI><<<grid, LG_THREADS, 0, st>>>(A);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

static LinearArgs lin(const float* X, int64_t ldx, const float* W, int64_t ldw, const float* bias, float* Y, int64_t ldy,
                      int64_t M, int R) {
  LinearArgs A{};
  A.X = X; A.ldx = ldx; A.W = W; A.ldw = ldw; A.bias = bias; A.Y = Y; A.ldy = ldy; A.M = M; A.R = R;
  A.rng.thresh = 0; A.rng.scale = 1.f;
  return A;
}

// =============================================================================================
// weight gradient: dW[o][k] += sum_m dZ[m][o] * X[m][k]   (and dbias[o] += sum_m dZ[m][o])
// grid (row slabs, O/64, K/64); 64x64 output tile per CTA, 4x4 per thread; float atomics at the end.
// =============================================================================================
constexpr int WG_THREADS = 256;
constexpr int WG_MC = 32;

__global__ void __launch_bounds__(WG_THREADS) wgrad_kernel(const float* __restrict__ dZ, int64_t ldz,
                                                            const float* __restrict__ X, int64_t ldx, int64_t M,
                                                            int64_t slab, float* __restrict__ dW, int64_t ldw,
                                                            float* __restrict__ dbias) {
  __shared__ __align__(16) float Zs[WG_MC][64];
  __shared__ __align__(16) float Xs[WG_MC][64];
  const int to = threadIdx.x >> 4, tk = threadIdx.x & 15;
  const int o0 = blockIdx.y * 64, k0 = blockIdx.z * 64;
  const int64_t m_begin = (int64_t)blockIdx.x * slab, m_end = min(M, m_begin + slab);
  float acc[4][4] = {};
  float bsum[4] = {};
  const bool do_bias = dbias != nullptr && blockIdx.z == 0 && tk == 0;
  for (int64_t m0 = m_begin; m0 < m_end; m0 += WG_MC) {
    __syncthreads();
    for (int i = threadIdx.x; i < WG_MC * 16; i += WG_THREADS) {
      const int r = i >> 4, c4 = i & 15;
      const int64_t m = m0 + r;
      float4 z = make_float4(0, 0, 0, 0), x = z;
      if (m < m_end) {
        z = ldg4(dZ + m * ldz + o0 + 4 * c4);
        x = ldg4(X + m * ldx + k0 + 4 * c4);
      }
      *reinterpret_cast<float4*>(&Zs[r][4 * c4]) = z;
      *reinterpret_cast<float4*>(&Xs[r][4 * c4]) = x;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < WG_MC; ++r) {
      const float4 z = *reinterpret_cast<const float4*>(&Zs[r][4 * to]);
      const float4 x = *reinterpret_cast<const float4*>(&Xs[r][4 * tk]);
      const float zz[4] = {z.x, z.y, z.z, z.w}, xx[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(zz[a], xx[b], acc[a][b]);
        bsum[a] += zz[a];
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int b = 0; b < 4; ++b) atomicAdd(dW + (int64_t)(o0 + 4 * to + a) * ldw + k0 + 4 * tk + b, acc[a][b]);
    if (do_bias) atomicAdd(dbias + o0 + 4 * to + a, bsum[a]);
  }
}

static int launch_wgrad(const float* dZ, int64_t ldz, int O, const float* X, int64_t ldx, int K, int64_t M, float* dW,
                        int64_t ldw, float* dbias, cudaStream_t st) {
  if (M == 0) return NCF_OK;
  const int tiles = (O / 64) * (K / 64);
  int64_t slabs = std::max<int64_t>(1, (int64_t)num_sms() * 4 / tiles);
  int64_t slab = align_up((M + slabs - 1) / slabs, WG_MC);
  slab = std::max<int64_t>(slab, 4 * WG_MC);
  slabs = (M + slab - 1) / slab;
  dim3 grid((unsigned)slabs, O / 64, K / 64);
  wgrad_kernel<<<grid, WG_THREADS, 0, st>>>(dZ, ldz, X, ldx, M, slab, dW, ldw, dbias);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

// =============================================================================================
// backward of ReLU -> LayerNorm -> Dropout for one MLP layer (one warp per row)
//   dY [M,J] gradient wrt the layer output; Rsv [M,J] saved relu output; -> dZ [M,J] gradient wrt
//   the pre-activation; accumulates d gamma, d beta, d bias.
// =============================================================================================
template <int J>
__global__ void __launch_bounds__(256) relu_ln_drop_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ Rsv,
                                                                const float* __restrict__ gamma, float* __restrict__ dZ,
                                                                int64_t M, DropoutRng rng, float* __restrict__ dgamma,
                                                                float* __restrict__ dbeta, float* __restrict__ dbias) {
  using CM = ColMap<J>;
  __shared__ float s_red[3][8][J];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t gw = (int64_t)blockIdx.x * 8 + warp, gstride = (int64_t)gridDim.x * 8;
  float gam[CM::CN], a_dg[CM::CN], a_db[CM::CN], a_dz[CM::CN];
#pragma unroll
  for (int g = 0; g < CM::NG; ++g)
#pragma unroll
    for (int v = 0; v < CM::VW; ++v) {
      gam[g * CM::VW + v] = __ldg(gamma + CM::col(lane, g) + v);
      a_dg[g * CM::VW + v] = a_db[g * CM::VW + v] = a_dz[g * CM::VW + v] = 0.f;
    }
  for (int64_t row = gw; row < M; row += gstride) {
    float r[CM::CN], dy[CM::CN];
#pragma unroll
    for (int g = 0; g < CM::NG; ++g) {
      const int64_t o = row * J + CM::col(lane, g);
      if (CM::VW == 4) {
        const float4 a = ldg4(Rsv + o), b = ldg4(dY + o);
        r[g * 4] = a.x; r[g * 4 + 1] = a.y; r[g * 4 + 2] = a.z; r[g * 4 + 3] = a.w;
        dy[g * 4] = b.x; dy[g * 4 + 1] = b.y; dy[g * 4 + 2] = b.z; dy[g * 4 + 3] = b.w;
      } else {
        const float2 a = __ldg(reinterpret_cast<const float2*>(Rsv + o)), b = __ldg(reinterpret_cast<const float2*>(dY + o));
        r[0] = a.x; r[1] = a.y; dy[0] = b.x; dy[1] = b.y;
      }
      if (rng.thresh != 0u) {
        if (CM::VW == 4) rng.apply4((uint64_t)o, &dy[g * 4]);
        else rng.apply2((uint64_t)o, &dy[0]);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CM::CN; ++c) s += r[c];
    const float mean = warp_sum(s) * (1.0f / J);
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CM::CN; ++c) q = fmaf(r[c] - mean, r[c] - mean, q);
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / J) + LN_EPS);
    float xh[CM::CN], dyg[CM::CN], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < CM::CN; ++c) {
      xh[c] = (r[c] - mean) * rstd;
      a_dg[c] = fmaf(dy[c], xh[c], a_dg[c]);
      a_db[c] += dy[c];
      dyg[c] = dy[c] * gam[c];
      s1 += dyg[c];
      s2 = fmaf(dyg[c], xh[c], s2);
    }
    const float m1 = warp_sum(s1) * (1.0f / J), m2 = warp_sum(s2) * (1.0f / J);
    float dz[CM::CN];
#pragma unroll
    for (int c = 0; c < CM::CN; ++c) {
      const float dr = rstd * (dyg[c] - m1 - xh[c] * m2);
      dz[c] = r[c] > 0.f ? dr : 0.f;
      a_dz[c] += dz[c];
    }
#pragma unroll
    for (int g = 0; g < CM::NG; ++g) {
      float* p = dZ + row * J + CM::col(lane, g);
      if (CM::VW == 4) st4(p, make_float4(dz[g * 4], dz[g * 4 + 1], dz[g * 4 + 2], dz[g * 4 + 3]));
      else *reinterpret_cast<float2*>(p) = make_float2(dz[0], dz[1]);
    }
  }
  // block reduce the three column sums, one atomic per column per block
#pragma unroll
  for (int g = 0; g < CM::NG; ++g)
#pragma unroll
    for (int v = 0; v < CM::VW; ++v) {
      const int c = CM::col(lane, g) + v, i = g * CM::VW + v;
      s_red[0][warp][c] = a_dg[i];
      s_red[1][warp][c] = a_db[i];
      s_red[2][warp][c] = a_dz[i];
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * J; i += 256) {
    const int which = i / J, c = i % J;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_red[which][w][c];
    float* dst = which == 0 ? dgamma : which == 1 ? dbeta : dbias;
    atomicAdd(dst + c, t);
  }
}

template <int J>
static int launch_relu_ln_drop_bwd(const float* dY, const float* Rsv, const float* gamma, float* dZ, int64_t M,
                                   const DropoutRng& rng, float* dgamma, float* dbeta, float* dbias, cudaStream_t st) {
  if (M == 0) return NCF_OK;
  const int grid = (int)std::min<int64_t>((M + 7) / 8, (int64_t)num_sms() * 8);
  relu_ln_drop_bwd_kernel<J><<<grid, 256, 0, st>>>(dY, Rsv, gamma, dZ, M, rng, dgamma, dbeta, dbias);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

// =============================================================================================
// attention core over the S rows of one interaction (architecture.py:40-55)
//   q [N,64]; kv [N,128] = [k | v]; ctx [N,64]
// 16 lanes own one (interaction, head): lane d holds element d of the head slice of every row, so
// each row slice is ONE coalesced 64-byte load per half warp (instead of every thread re-reading whole
// rows), the S x S dot products are 16-lane shuffle all-reduces, and the outputs are written the
// same way.  The keep mask of the probability dropout is drawn once per element and shared by ballot.
// =============================================================================================
__device__ __forceinline__ float hw_allreduce(float v) {      // sum over the 16 lanes of a half warp
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// bit e of the result = keep decision of element base + e (e < S*S), identical on all 16 lanes
template <int S>
__device__ __forceinline__ uint64_t attn_keep_mask(const DropoutRng& rng, uint64_t base, int d, int lane) {
  uint64_t m = ~0ull;
  if (rng.thresh != 0u) {
    m = 0;
#pragma unroll
    for (int r = 0; r < (S * S + 15) / 16; ++r) {
      const int e = r * 16 + d;
      const bool k = e < S * S ? rng.keep(base + e) : false;
      const uint32_t b = __ballot_sync(0xffffffffu, k);
      m |= (uint64_t)((lane & 16) ? (b >> 16) : (b & 0xffffu)) << (16 * r);
    }
  }
  return m;
}

// forward: thread = (row, head); each thread reads its query slice and the S key / value slices of its group
__device__ __forceinline__ void load16(const float* p, float* d) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = ldg4(p + 4 * i);
    d[4 * i] = t.x; d[4 * i + 1] = t.y; d[4 * i + 2] = t.z; d[4 * i + 3] = t.w;
  }
}
template <int S>
__global__ void __launch_bounds__(256) attn_core_fwd_kernel(const float* __restrict__ q, const float* __restrict__ kv,
                                                            float* __restrict__ ctx, int64_t N, DropoutRng rng) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * HEADS) return;
  const int64_t n = t / HEADS;
  const int h = (int)(t % HEADS);
  const int64_t g = n / S;
  const int i = (int)(n % S);
  float qv[16], p[S], o[16];
  load16(q + n * D + h * HD, qv);
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < S; ++j) {
    float kk[16];
    load16(kv + (g * S + j) * (2 * D) + h * HD, kk);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) s = fmaf(qv[c], kk[c], s);
    p[j] = s * 0.25f;   // / sqrt(head_dim = 16)
    mx = fmaxf(mx, p[j]);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < S; ++j) {
    p[j] = expf(p[j] - mx);
    sum += p[j];
  }
  const float inv = 1.0f / sum;
#pragma unroll
  for (int c = 0; c < 16; ++c) o[c] = 0.f;
#pragma unroll
  for (int j = 0; j < S; ++j) {
    float pj = p[j] * inv;
    if (rng.thresh != 0u) pj = rng.keep((((uint64_t)g * HEADS + h) * S + i) * S + j) ? pj * rng.scale : 0.f;
    float vv[16];
    load16(kv + (g * S + j) * (2 * D) + D + h * HD, vv);
#pragma unroll
    for (int c = 0; c < 16; ++c) o[c] = fmaf(pj, vv[c], o[c]);
  }
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) st4(ctx + n * D + h * HD + 4 * c4, make_float4(o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2], o[4 * c4 + 3]));
}

template <int S>
__global__ void __launch_bounds__(256) attn_core_bwd_kernel(const float* __restrict__ q, const float* __restrict__ kv,
                                                            const float* __restrict__ dctx, float* __restrict__ dq,
                                                            float* __restrict__ dkv, int64_t N, DropoutRng rng) {
  const int lane = threadIdx.x & 31, d = lane & 15;
  const int64_t hw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int64_t ngh = (N / S) * HEADS;
  const bool live = hw < ngh;
  const int64_t g = live ? hw / HEADS : 0;
  const int h = live ? (int)(hw % HEADS) : 0;
  float qv[S], kk[S], vv[S], dc[S], dk[S], dv[S];
#pragma unroll
  for (int i = 0; i < S; ++i) {
    const int64_t n = g * S + i;
    qv[i] = __ldg(q + n * D + h * HD + d);
    kk[i] = __ldg(kv + n * (2 * D) + h * HD + d);
    vv[i] = __ldg(kv + n * (2 * D) + D + h * HD + d);
    dc[i] = __ldg(dctx + n * D + h * HD + d);
    dk[i] = 0.f;
    dv[i] = 0.f;
  }
  const uint64_t keep = attn_keep_mask<S>(rng, ((uint64_t)g * HEADS + h) * (S * S), d, lane);
#pragma unroll
  for (int i = 0; i < S; ++i) {
    float p[S], dp[S], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      p[j] = hw_allreduce(qv[i] * kk[j]) * 0.25f;
      dp[j] = hw_allreduce(dc[i] * vv[j]);              // dL/dp'_ij
      mx = fmaxf(mx, p[j]);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      p[j] = expf(p[j] - mx);
      sum += p[j];
    }
    const float inv = 1.0f / sum;
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      p[j] *= inv;
      const float ks = ((keep >> (i * S + j)) & 1ull) ? rng.scale : 0.f;
      dv[j] = fmaf(p[j] * ks, dc[i], dv[j]);            // dv_j += p'_ij dctx_i
      dp[j] *= ks;                                      // dL/dp_ij
      dot = fmaf(p[j], dp[j], dot);
    }
    float dqi = 0.f;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      const float ds = p[j] * (dp[j] - dot) * 0.25f;    // dL/ds_ij (scaled by 1/sqrt(16))
      dqi = fmaf(ds, kk[j], dqi);
      dk[j] = fmaf(ds, qv[i], dk[j]);
    }
    if (live) dq[(g * S + i) * D + h * HD + d] = dqi;
  }
  if (live) {
#pragma unroll
    for (int j = 0; j < S; ++j) {
      dkv[(g * S + j) * (2 * D) + h * HD + d] = dk[j];
      dkv[(g * S + j) * (2 * D) + D + h * HD + d] = dv[j];
    }
  }
}

#define NCF_DISPATCH_S(S_, CALL)                                     \
  switch (S_) {                                                      \
    case 1: { constexpr int SS = 1; CALL; break; }                   \
    case 2: { constexpr int SS = 2; CALL; break; }                   \
    case 3: { constexpr int SS = 3; CALL; break; }                   \
    case 4: { constexpr int SS = 4; CALL; break; }                   \
    case 5: { constexpr int SS = 5; CALL; break; }                   \
    case 6: { constexpr int SS = 6; CALL; break; }                   \
    case 7: { constexpr int SS = 7; CALL; break; }                   \
    case 8: { constexpr int SS = 8; CALL; break; }                   \
    default: set_error("attention: S=%d unsupported", S_); return NCF_ERR_ARG; \
  }

static int launch_attn_core_fwd(const float* q, const float* kv, float* ctx, int64_t N, int S, const DropoutRng& rng,
                                cudaStream_t st) {
  if (N == 0) return NCF_OK;
  const int64_t threads = N * HEADS;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  NCF_DISPATCH_S(S, (attn_core_fwd_kernel<SS><<<grid, 256, 0, st>>>(q, kv, ctx, N, rng)));
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}
static int launch_attn_core_bwd(const float* q, const float* kv, const float* dctx, float* dq, float* dkv, int64_t N, int S,
                                const DropoutRng& rng, cudaStream_t st) {
  if (N == 0) return NCF_OK;
  const int64_t threads = (N / S) * HEADS * 16;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  NCF_DISPATCH_S(S, (attn_core_bwd_kernel<SS><<<grid, 256, 0, st>>>(q, kv, dctx, dq, dkv, N, rng)));
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

// =============================================================================================
// output head: mlp_pred = mlp_output(h3); out = sigmoid(final([mf_pred, mlp_pred]))
// (architecture.py:345-354); half warp per row
// =============================================================================================
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ h3, const float* __restrict__ mf_pred,
                                                       const float* __restrict__ dense, float* __restrict__ mlp_pred,
                                                       float* __restrict__ out, float* __restrict__ out2, int64_t N) {
  const int lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15;
  const int64_t hw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int64_t n = min(hw, N - 1);
  const float4 w = ldg4(dense + NCF_OFF(NCF_P_MLP_OUT_W) + 4 * l16);
  const float s = half_warp_sum(f4_dot(ldg4(h3 + n * H3 + 4 * l16), w));
  if (hw < N && l16 == 0) {
    const float mp = s + __ldg(dense + NCF_OFF(NCF_P_MLP_OUT_B));
    const float z = fmaf(__ldg(dense + NCF_OFF(NCF_P_FINAL_W)), mf_pred[n],
                         fmaf(__ldg(dense + NCF_OFF(NCF_P_FINAL_W) + 1), mp, __ldg(dense + NCF_OFF(NCF_P_FINAL_B))));
    const float p = 1.0f / (1.0f + expf(-z));
    mlp_pred[n] = mp;
    out[n] = p;
    if (out2) out2[n] = p;
  }
  (void)half;
}

// backward of the head: grad_out = dL/d out.  Produces d_mf_pred [N], dh3 [N,64]; accumulates the
// gradients of final.0, mlp_output, mf_output.bias.
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ grad_out, const float* __restrict__ p_saved,
                                                       const float* __restrict__ mf_pred, const float* __restrict__ mlp_pred,
                                                       const float* __restrict__ h3, const float* __restrict__ dense,
                                                       float* __restrict__ d_mf_pred, float* __restrict__ dh3,
                                                       float* __restrict__ d_mlp_pred, float* __restrict__ dense_grad,
                                                       int64_t N) {
  __shared__ float s_w[8][H3];
  __shared__ float s_sc[8][5];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, l16 = lane & 15, half = lane >> 4;
  const int64_t hw0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int64_t hstride = ((int64_t)gridDim.x * blockDim.x) >> 4;
  const float a = __ldg(dense + NCF_OFF(NCF_P_FINAL_W)), c = __ldg(dense + NCF_OFF(NCF_P_FINAL_W) + 1);
  const float4 w = ldg4(dense + NCF_OFF(NCF_P_MLP_OUT_W) + 4 * l16);
  float4 dw = make_float4(0, 0, 0, 0);
  float s_a = 0.f, s_c = 0.f, s_d = 0.f, s_bmf = 0.f, s_bmlp = 0.f;
  // all 16 lanes of a half warp walk the same rows, four rows per trip so that their loads are in flight together
  for (int64_t n0 = hw0; n0 < N; n0 += 4 * hstride) {
    float p[4], go[4], mfp[4], mlpp[4];
    float4 h[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t n = min(n0 + u * hstride, N - 1);
      p[u] = p_saved[n];
      go[u] = grad_out[n];
      mfp[u] = mf_pred[n];
      mlpp[u] = mlp_pred[n];
      h[u] = ldg4(h3 + n * H3 + 4 * l16);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t n = n0 + u * hstride;
      if (n < N) {
        const float dz = go[u] * p[u] * (1.0f - p[u]);
        const float dmf = dz * a, dml = dz * c;
        dw = f4_fma(dml, h[u], dw);
        if (dh3) st4(dh3 + n * H3 + 4 * l16, make_float4(dml * w.x, dml * w.y, dml * w.z, dml * w.w));
        if (l16 == 0) {
          d_mf_pred[n] = dmf;
          if (d_mlp_pred) d_mlp_pred[n] = dml;     // the tcgen05 MLP backward forms dh3 = d_mlp_pred * w itself
          s_a = fmaf(dz, mfp[u], s_a);
          s_c = fmaf(dz, mlpp[u], s_c);
          s_d += dz;
          s_bmf += dmf;
          s_bmlp += dml;
        }
      }
    }
  }
  // combine the two half warps, then the block
  dw.x += __shfl_xor_sync(0xffffffffu, dw.x, 16);
  dw.y += __shfl_xor_sync(0xffffffffu, dw.y, 16);
  dw.z += __shfl_xor_sync(0xffffffffu, dw.z, 16);
  dw.w += __shfl_xor_sync(0xffffffffu, dw.w, 16);
  s_a += __shfl_xor_sync(0xffffffffu, s_a, 16);
  s_c += __shfl_xor_sync(0xffffffffu, s_c, 16);
  s_d += __shfl_xor_sync(0xffffffffu, s_d, 16);
  s_bmf += __shfl_xor_sync(0xffffffffu, s_bmf, 16);
  s_bmlp += __shfl_xor_sync(0xffffffffu, s_bmlp, 16);
  if (half == 0) {
    s_w[warp][4 * l16 + 0] = dw.x; s_w[warp][4 * l16 + 1] = dw.y; s_w[warp][4 * l16 + 2] = dw.z; s_w[warp][4 * l16 + 3] = dw.w;
  }
  if (lane == 0) {
    s_sc[warp][0] = s_a; s_sc[warp][1] = s_c; s_sc[warp][2] = s_d; s_sc[warp][3] = s_bmf; s_sc[warp][4] = s_bmlp;
  }
  __syncthreads();
  if (threadIdx.x < H3) {
    float t = 0.f;
    for (int wv = 0; wv < 8; ++wv) t += s_w[wv][threadIdx.x];
    atomicAdd(dense_grad + NCF_OFF(NCF_P_MLP_OUT_W) + threadIdx.x, t);
  } else if (threadIdx.x < H3 + 5) {
    const int k = threadIdx.x - H3;
    float t = 0.f;
    for (int wv = 0; wv < 8; ++wv) t += s_sc[wv][k];
    const int64_t off = k == 0 ? NCF_OFF(NCF_P_FINAL_W) : k == 1 ? NCF_OFF(NCF_P_FINAL_W) + 1
                        : k == 2 ? NCF_OFF(NCF_P_FINAL_B) : k == 3 ? NCF_OFF(NCF_P_MF_OUT_B) : NCF_OFF(NCF_P_MLP_OUT_B);
    atomicAdd(dense_grad + off, t);
  }
}

// The tcgen05 MLP backward forms dh3 and d mlp_output.weight itself (from its LayerNorm column sums), so its head
// backward is per-row scalar work only: d_mf_pred, d_mlp_pred and the gradients of final.0 and the two output biases.
__global__ void __launch_bounds__(256) head_bwd_scalar_kernel(const float* __restrict__ grad_out, const float* __restrict__ p_saved,
                                                              const float* __restrict__ mf_pred, const float* __restrict__ mlp_pred,
                                                              const float* __restrict__ dense, float* __restrict__ d_mf_pred,
                                                              float* __restrict__ d_mlp_pred, float* __restrict__ dense_grad,
                                                              int64_t N) {
  __shared__ float s_sc[8][5];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float a = __ldg(dense + NCF_OFF(NCF_P_FINAL_W)), c = __ldg(dense + NCF_OFF(NCF_P_FINAL_W) + 1);
  float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    const float p = p_saved[n];
    const float dz = grad_out[n] * p * (1.0f - p);
    const float dmf = dz * a, dml = dz * c;
    d_mf_pred[n] = dmf;
    d_mlp_pred[n] = dml;
    s[0] = fmaf(dz, mf_pred[n], s[0]);
    s[1] = fmaf(dz, mlp_pred[n], s[1]);
    s[2] += dz;
    s[3] += dmf;
    s[4] += dml;
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], off);
    if (lane == 0) s_sc[warp][k] = s[k];
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    const int k = threadIdx.x;
    float t = 0.f;
    for (int wv = 0; wv < 8; ++wv) t += s_sc[wv][k];
    const int64_t off = k == 0 ? NCF_OFF(NCF_P_FINAL_W) : k == 1 ? NCF_OFF(NCF_P_FINAL_W) + 1
                        : k == 2 ? NCF_OFF(NCF_P_FINAL_B) : k == 3 ? NCF_OFF(NCF_P_MF_OUT_B) : NCF_OFF(NCF_P_MLP_OUT_B);
    atomicAdd(dense_grad + off, t);
  }
}

// nn.BCELoss (mean) + gradient (ATen binary_cross_entropy: log clamped at -100, backward
// (p - y) / max(p (1 - p), 1e-12) / N)
__global__ void __launch_bounds__(256) bce_kernel(const float* __restrict__ out, const float* __restrict__ tgt, int64_t N,
                                                  float* __restrict__ loss, float* __restrict__ grad) {
  __shared__ float s_part[8];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc = 0.f;
  const float invN = 1.0f / (float)N;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    const float p = out[n], y = tgt[n];
    const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.0f - p), -100.f);
    acc -= y * lp + (1.0f - y) * l1p;
    if (grad) grad[n] = (p - y) / fmaxf(p * (1.0f - p), 1e-12f) * invN;
  }
  acc = warp_sum(acc);
  if (lane == 0) s_part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    atomicAdd(loss, t * invN);
  }
}

// bce_kernel + head_bwd_scalar_kernel in one pass (ncf_train_step, bf16 towers): the loss gradient of a row is consumed where
// it is formed instead of travelling through memory to a second small kernel between the forward and the backward towers.
// Same expressions in the same order as the two kernels, so d_mf / d_mlp come out bit-identical.
__global__ void __launch_bounds__(256) bce_head_bwd_kernel(const float* __restrict__ out, const float* __restrict__ tgt, int64_t N,
                                                           float* __restrict__ loss, const float* __restrict__ mf_pred,
                                                           const float* __restrict__ mlp_pred, const float* __restrict__ dense,
                                                           float* __restrict__ d_mf_pred, float* __restrict__ d_mlp_pred,
                                                           float* __restrict__ dense_grad) {
  __shared__ float s_sc[8][6];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float a = __ldg(dense + NCF_OFF(NCF_P_FINAL_W)), c = __ldg(dense + NCF_OFF(NCF_P_FINAL_W) + 1);
  const float invN = 1.0f / (float)N;
  float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    const float p = out[n], y = tgt[n];
    const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.0f - p), -100.f);
    s[5] -= y * lp + (1.0f - y) * l1p;
    const float g = (p - y) / fmaxf(p * (1.0f - p), 1e-12f) * invN;
    const float dz = g * p * (1.0f - p);
    const float dmf = dz * a, dml = dz * c;
    d_mf_pred[n] = dmf;
    d_mlp_pred[n] = dml;
    s[0] = fmaf(dz, mf_pred[n], s[0]);
    s[1] = fmaf(dz, mlp_pred[n], s[1]);
    s[2] += dz;
    s[3] += dmf;
    s[4] += dml;
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], off);
    if (lane == 0) s_sc[warp][k] = s[k];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    const int k = threadIdx.x;
    float t = 0.f;
    for (int wv = 0; wv < 8; ++wv) t += s_sc[wv][k];
    if (k == 5) {
      atomicAdd(loss, t * invN);
    } else {
      const int64_t off = k == 0 ? NCF_OFF(NCF_P_FINAL_W) : k == 1 ? NCF_OFF(NCF_P_FINAL_W) + 1
                          : k == 2 ? NCF_OFF(NCF_P_FINAL_B) : k == 3 ? NCF_OFF(NCF_P_MF_OUT_B) : NCF_OFF(NCF_P_MLP_OUT_B);
      atomicAdd(dense_grad + off, t);
    }
  }
}

// =============================================================================================
// orchestration
// =============================================================================================
// NCF_ATTN_FUSED: 1 (default) = fused tcgen05 attention block for S = 5; 0 = separate projection / core kernels
static int attn_fused_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("NCF_ATTN_FUSED");
    mode = e ? atoi(e) : 1;
  }
  return mode;
}
bool tower_bf16_rows(const ncf_run_cfg& cfg) { return cfg.precision == NCF_BF16_TC && cfg.S == 5 && attn_fused_mode() == 1; }

TowerWs carve_tower_ws(void* ws, int64_t N, const ncf_run_cfg& cfg) {
  TowerWs w{};
  Carver c(ws);
  const bool train = cfg.training != 0;
  w.mf_pred = c.take<float>(N);
  w.mlp_pred = c.take<float>(N);
  w.p_saved = c.take<float>(N);
  w.xu = c.take<float>(N * D);
  w.xp = c.take<float>(N * D);
  w.q = c.take<float>(N * D);
  w.kv = c.take<float>(N * 2 * D);
  w.ctx = c.take<float>(N * D);
  w.a = c.take<float>(N * D);
  w.r1 = c.take<float>(N * H1);
  w.y1 = c.take<float>(N * H1);
  w.r2 = c.take<float>(N * H2);
  w.y2 = c.take<float>(N * H2);
  w.r3 = c.take<float>(N * H3);
  w.y3 = c.take<float>(N * H3);
  if (cfg.precision == NCF_BF16_TC) {
    const int64_t Np = align_up(N, 128);
    w.a_img = c.take<uint16_t>(Np * D);
    w.st1 = c.take<float>(Np * 2);
    w.st2 = c.take<float>(Np * 2);
    w.st3 = c.take<float>(Np * 2);
  }
  if (train) {
    w.y_pmf = c.take<float>(N * D);
    w.y_umf = c.take<float>(N * D);
    w.d_mf = c.take<float>(N);
    w.d_mlp = c.take<float>(N);
    w.g64a = c.take<float>(N * D);
    w.g64b = c.take<float>(N * D);
    w.g128 = c.take<float>(N * 2 * D);
    w.g128b = c.take<float>(N * 2 * D);
    w.g256 = c.take<float>(N * H1);
    w.g256b = c.take<float>(N * H1);
    if (cfg.precision == NCF_BF16_TC) {
      const int64_t Np = align_up(N, 128);      // tile images: whole 128-row tiles
      w.r1b = c.take<uint16_t>(Np * H1);
      w.y1b = c.take<uint16_t>(Np * H1);
      w.r2b = c.take<uint16_t>(Np * H2);
      w.y2b = c.take<uint16_t>(Np * H2);
      w.r3b = c.take<uint16_t>(Np * H3);
      w.dz1b = c.take<uint16_t>(Np * H1);
      w.dz2b = c.take<uint16_t>(Np * H2);
      w.dz3b = c.take<uint16_t>(Np * H3);
      w.wg_partial = c.take<float>((int64_t)num_sms() * MLP_WG_PART_COLS * 128);
      w.at_partial = c.take<float>(attn_tc_partial_floats());
    }
    w.emb_bytes = ncf_emb_bwd_workspace_bytes(N);
    w.emb = c.take<char>(w.emb_bytes);
  }
  w.total = align_up(c.used, 256);
  return w;
}

int tower_f32_forward(const ncf_run_cfg& cfg, const float* dense, int64_t N, const int64_t* hour, const float* tail1,
                      float* out, TowerWs& w, cudaStream_t st) {
  NCF_TRY(tower_attn_forward(cfg, dense, N, w, st));
  return tower_mlp_forward(cfg, dense, N, hour, tail1, out, w, st);
}

// attention block: w.xu, w.xp -> w.a (fp32 path) / w.a_img (tcgen05 path)
int tower_attn_forward(const ncf_run_cfg& cfg, const float* dense, int64_t N, TowerWs& w, cudaStream_t st) {
  const bool train = cfg.training != 0;
  const int S = cfg.S;
  const float* P = dense;
  const bool tc = cfg.precision == NCF_BF16_TC;
  const bool fused_attn = tower_bf16_rows(cfg);
  if (fused_attn) {
    NCF_TRY(attn_tc_forward(cfg, dense, N, w, st));       // xu, xp -> a_img in one kernel; nothing else is kept
  } else if (train || S > 1) {
    // q = q_proj(xu); [k|v] = [k_proj; v_proj](xp)        (architecture.py:40-42)
    if (tc) {
      NCF_TRY(tc_proj_forward(0, w.xu, P + NCF_OFF(NCF_P_Q_W), P + NCF_OFF(NCF_P_Q_B), w.q, N, st));
      NCF_TRY(tc_proj_forward(1, w.xp, P + NCF_OFF(NCF_P_K_W), P + NCF_OFF(NCF_P_K_B), w.kv, N, st));
    } else {
      NCF_TRY((launch_linear<64, false, EPI_NONE>(lin(w.xu, D, P + NCF_OFF(NCF_P_Q_W), D, P + NCF_OFF(NCF_P_Q_B), w.q, D, N, D), st)));
      NCF_TRY((launch_linear<128, false, EPI_NONE>(lin(w.xp, D, P + NCF_OFF(NCF_P_K_W), D, P + NCF_OFF(NCF_P_K_B), w.kv, 2 * D, N, D), st)));
    }
    NCF_TRY(launch_attn_core_fwd(w.q, w.kv, w.ctx, N, S, make_rng(cfg, 0), st));
  } else if (tc) {
    NCF_TRY(tc_proj_forward(0, w.xp, P + NCF_OFF(NCF_P_V_W), P + NCF_OFF(NCF_P_V_B), w.ctx, N, st));
  } else {
    // one key per query: softmax == 1, ctx = v_proj(xp)      (architecture.py:275-276)
    NCF_TRY((launch_linear<64, false, EPI_NONE>(lin(w.xp, D, P + NCF_OFF(NCF_P_V_W), D, P + NCF_OFF(NCF_P_V_B), w.ctx, D, N, D), st)));
  }
  if (fused_attn) {
  } else if (tc) {
    NCF_TRY(tc_proj_forward_img(w.ctx, P + NCF_OFF(NCF_P_O_W), P + NCF_OFF(NCF_P_O_B), w.a_img, N, st));
  } else
    NCF_TRY((launch_linear<64, false, EPI_NONE>(lin(w.ctx, D, P + NCF_OFF(NCF_P_O_W), D, P + NCF_OFF(NCF_P_O_B), w.a, D, N, D), st)));
  return NCF_OK;
}

// MLP tower + output head: w.a / w.a_img (+ w.mf_pred) -> out
int tower_mlp_forward(const ncf_run_cfg& cfg, const float* dense, int64_t N, const int64_t* hour, const float* tail1,
                      float* out, TowerWs& w, cudaStream_t st) {
  const bool train = cfg.training != 0;
  const float* P = dense;
  if (cfg.precision == NCF_BF16_TC) return mlp_tc_forward(cfg, dense, N, hour, tail1, out, w, st);
  // MLP: the 32 temporal input columns are zeros in forward (architecture.py:329-340), so only the
  // first 64 columns of mlp.0.weight take part; forward_simple's hour path adds tail1[hour].
  LinearArgs l1 = lin(w.a, D, P + NCF_OFF(NCF_P_MLP0_W), K0, P + NCF_OFF(NCF_P_MLP0_B), w.y1, H1, N, D);
  l1.gamma = P + NCF_OFF(NCF_P_LN0_W); l1.beta = P + NCF_OFF(NCF_P_LN0_B); l1.Rout = train ? w.r1 : nullptr;
  l1.hour = hour; l1.tail1 = tail1; l1.rng = make_rng(cfg, 1);
  NCF_TRY((launch_linear<256, false, EPI_RELU_LN_DROP>(l1, st)));
  LinearArgs l2 = lin(w.y1, H1, P + NCF_OFF(NCF_P_MLP1_W), H1, P + NCF_OFF(NCF_P_MLP1_B), w.y2, H2, N, H1);
  l2.gamma = P + NCF_OFF(NCF_P_LN1_W); l2.beta = P + NCF_OFF(NCF_P_LN1_B); l2.Rout = train ? w.r2 : nullptr;
  l2.rng = make_rng(cfg, 2);
  NCF_TRY((launch_linear<128, false, EPI_RELU_LN_DROP>(l2, st)));
  LinearArgs l3 = lin(w.y2, H2, P + NCF_OFF(NCF_P_MLP2_W), H2, P + NCF_OFF(NCF_P_MLP2_B), w.y3, H3, N, H2);
  l3.gamma = P + NCF_OFF(NCF_P_LN2_W); l3.beta = P + NCF_OFF(NCF_P_LN2_B); l3.Rout = train ? w.r3 : nullptr;
  l3.rng = make_rng(cfg, 3);
  NCF_TRY((launch_linear<64, false, EPI_RELU_LN_DROP>(l3, st)));
  head_fwd_kernel<<<(unsigned)((N * 16 + 255) / 256), 256, 0, st>>>(w.y3, w.mf_pred, P, w.mlp_pred, out, w.p_saved, N);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

// produces w.d_mf [N], dxu = w.g64b, dxp = w.g256 and accumulates every dense gradient
int tower_side_join(cudaStream_t st) {
  AuxCtx* aux = aux_ctx();
  if (aux->side_pending) {
    aux->side_pending = false;
    NCF_CUDA(cudaStreamWaitEvent(st, aux->ev[5], 0));
  }
  return NCF_OK;
}

int tower_f32_backward(const ncf_run_cfg& cfg, const float* dense, float* dg, int64_t N, const float* grad_out,
                       TowerWs& w, cudaStream_t st, bool defer_join, bool head_done) {
  // bf16 towers with an auxiliary stream set: the MLP weight-gradient kernel (HBM-bound) runs on a few SMs NEXT TO the
  // attention backward (issue-bound, 1.3 TB/s of DRAM traffic), which leaves them those SMs
  AuxCtx* aux = aux_ctx();
  const int side_sms = wgrad_side_sms();
  if (side_sms > 0 && side_sms < tower_sms() && aux->stream && tower_bf16_rows(cfg) && N >= (int64_t)128 * num_sms()) {
    NCF_TRY(aux_events(aux));
    NCF_TRY(tower_mlp_backward(cfg, dense, dg, N, grad_out, w, st, aux->side, side_sms, aux->ev[4], nullptr, head_done));
    NCF_TRY(attn_tc_backward(cfg, dense, dg, N, w, st, side_sms, aux->side, aux->ev[6]));     // its partial-sum reduction: side stream too
    NCF_CUDA(cudaEventRecord(aux->ev[5], aux->side));
    w.dxu = w.g64b;
    w.dxp = w.g256;
    // nothing up to the dense Adam reads the MLP weight gradients: a caller that has more work for `st` (the embedding
    // backward) joins the side stream behind it (tower_side_join), everyone else here
    if (defer_join) aux->side_pending = true;
    else NCF_CUDA(cudaStreamWaitEvent(st, aux->ev[5], 0));
    return NCF_OK;
  }
  NCF_TRY(tower_mlp_backward(cfg, dense, dg, N, grad_out, w, st, nullptr, 0, nullptr, nullptr, head_done));
  return tower_attn_backward(cfg, dense, dg, N, w, st);
}

// head + MLP tower: grad_out -> w.d_mf, da (w.g64a) + their parameter gradients
int tower_mlp_backward(const ncf_run_cfg& cfg, const float* dense, float* dg, int64_t N, const float* grad_out,
                       TowerWs& w, cudaStream_t st, cudaStream_t side, int side_sms, cudaEvent_t fork, cudaEvent_t join,
                       bool head_done) {
  const float* P = dense;
  if (!cfg.training) { set_error("backward needs a training-mode forward"); return NCF_ERR_ARG; }
  const int hgrid = (int)std::min<int64_t>((N * 16 + 255) / 256, (int64_t)num_sms() * 8);
  // head: d_mf, dh3 (-> g64a)
  const bool tcm = cfg.precision == NCF_BF16_TC;
  if (tcm && head_done) {
    // launch_bce_head_bwd has produced w.d_mf, w.d_mlp and the head's parameter gradients already
  } else if (tcm) {
    const int sgrid = (int)std::min<int64_t>((N + 255) / 256, (int64_t)num_sms() * 4);
    NCF_CUDA(launch_pdl(PDL_HEAD, head_bwd_scalar_kernel, dim3(sgrid), dim3(256), 0, st, grad_out, (const float*)w.p_saved, (const float*)w.mf_pred,
                        (const float*)w.mlp_pred, P, w.d_mf, w.d_mlp, dg, N));
  } else {
    head_bwd_kernel<<<hgrid, 256, 0, st>>>(grad_out, w.p_saved, w.mf_pred, w.mlp_pred, w.y3, P, w.d_mf, w.g64a, nullptr, dg, N);
  }
  NCF_LAUNCH_CHECK();
  if (cfg.precision == NCF_BF16_TC) {
    NCF_TRY(mlp_tc_backward(cfg, dense, dg, N, w, st, side, side_sms, fork, join));   // dy3 (g64a) -> da (g64a), all MLP parameter gradients
  } else {
  // layer 3: dz3 (g64b); dW8 += dz3^T y2 ; dy2 (g128) = dz3 . W8
  NCF_TRY(launch_relu_ln_drop_bwd<64>(w.g64a, w.r3, P + NCF_OFF(NCF_P_LN2_W), w.g64b, N, make_rng(cfg, 3),
                                      dg + NCF_OFF(NCF_P_LN2_W), dg + NCF_OFF(NCF_P_LN2_B), dg + NCF_OFF(NCF_P_MLP2_B), st));
  NCF_TRY(launch_wgrad(w.g64b, H3, H3, w.y2, H2, H2, N, dg + NCF_OFF(NCF_P_MLP2_W), H2, nullptr, st));
  NCF_TRY((launch_linear<128, true, EPI_NONE>(lin(w.g64b, H3, P + NCF_OFF(NCF_P_MLP2_W), H2, nullptr, w.g128, H2, N, H3), st)));
  // layer 2: dz2 (g128b); dW4 += dz2^T y1 ; dy1 (g256) = dz2 . W4
  NCF_TRY(launch_relu_ln_drop_bwd<128>(w.g128, w.r2, P + NCF_OFF(NCF_P_LN1_W), w.g128b, N, make_rng(cfg, 2),
                                       dg + NCF_OFF(NCF_P_LN1_W), dg + NCF_OFF(NCF_P_LN1_B), dg + NCF_OFF(NCF_P_MLP1_B), st));
  NCF_TRY(launch_wgrad(w.g128b, H2, H2, w.y1, H1, H1, N, dg + NCF_OFF(NCF_P_MLP1_W), H1, nullptr, st));
  NCF_TRY((launch_linear<256, true, EPI_NONE>(lin(w.g128b, H2, P + NCF_OFF(NCF_P_MLP1_W), H1, nullptr, w.g256, H1, N, H2), st)));
  // layer 1: dz1 (g256b); dW0[:, :64] += dz1^T a ; da (g64a) = dz1 . W0[:, :64]
  NCF_TRY(launch_relu_ln_drop_bwd<256>(w.g256, w.r1, P + NCF_OFF(NCF_P_LN0_W), w.g256b, N, make_rng(cfg, 1),
                                       dg + NCF_OFF(NCF_P_LN0_W), dg + NCF_OFF(NCF_P_LN0_B), dg + NCF_OFF(NCF_P_MLP0_B), st));
  NCF_TRY(launch_wgrad(w.g256b, H1, H1, w.a, D, D, N, dg + NCF_OFF(NCF_P_MLP0_W), K0, nullptr, st));
  NCF_TRY((launch_linear<64, true, EPI_NONE>(lin(w.g256b, H1, P + NCF_OFF(NCF_P_MLP0_W), K0, nullptr, w.g64a, D, N, H1), st)));
  }
  return NCF_OK;
}

// attention block: da (w.g64a) -> dxu (w.g64b), dxp (w.g256) + the projection gradients
int tower_attn_backward(const ncf_run_cfg& cfg, const float* dense, float* dg, int64_t N, TowerWs& w, cudaStream_t st) {
  const int S = cfg.S;
  const float* P = dense;
  if (!cfg.training) { set_error("backward needs a training-mode forward"); return NCF_ERR_ARG; }
  const bool tc = cfg.precision == NCF_BF16_TC;
  if (tower_bf16_rows(cfg)) {
    NCF_TRY(attn_tc_backward(cfg, dense, dg, N, w, st));
    w.dxu = w.g64b;
    w.dxp = w.g256;
    return NCF_OK;
  }
  // out_proj: dWo += da^T ctx, dbo ; dctx (g64b) = da . Wo
  if (tc) {
    NCF_TRY(tc_proj_backward(0, w.g64a, w.ctx, P + NCF_OFF(NCF_P_O_W), w.g64b, dg + NCF_OFF(NCF_P_O_W), dg + NCF_OFF(NCF_P_O_B), N, st));
  } else {
    NCF_TRY(launch_wgrad(w.g64a, D, D, w.ctx, D, D, N, dg + NCF_OFF(NCF_P_O_W), D, dg + NCF_OFF(NCF_P_O_B), st));
    NCF_TRY((launch_linear<64, true, EPI_NONE>(lin(w.g64a, D, P + NCF_OFF(NCF_P_O_W), D, nullptr, w.g64b, D, N, D), st)));
  }
  // attention core: dq (g64a), dkv (g128)
  NCF_TRY(launch_attn_core_bwd(w.q, w.kv, w.g64b, w.g64a, w.g128, N, S, make_rng(cfg, 0), st));
  // projections: weights/biases, then dxu (g64b) = dq . Wq and dxp (g256 reused as [N,64]) = dkv . [Wk;Wv]
  if (tc) {
    NCF_TRY(tc_proj_backward(0, w.g64a, w.xu, P + NCF_OFF(NCF_P_Q_W), w.g64b, dg + NCF_OFF(NCF_P_Q_W), dg + NCF_OFF(NCF_P_Q_B), N, st));
    NCF_TRY(tc_proj_backward(1, w.g128, w.xp, P + NCF_OFF(NCF_P_K_W), w.g256, dg + NCF_OFF(NCF_P_K_W), dg + NCF_OFF(NCF_P_K_B), N, st));
  } else {
    NCF_TRY(launch_wgrad(w.g64a, D, D, w.xu, D, D, N, dg + NCF_OFF(NCF_P_Q_W), D, dg + NCF_OFF(NCF_P_Q_B), st));
    NCF_TRY(launch_wgrad(w.g128, 2 * D, 2 * D, w.xp, D, D, N, dg + NCF_OFF(NCF_P_K_W), D, dg + NCF_OFF(NCF_P_K_B), st));
    NCF_TRY((launch_linear<64, true, EPI_NONE>(lin(w.g64a, D, P + NCF_OFF(NCF_P_Q_W), D, nullptr, w.g64b, D, N, D), st)));
    NCF_TRY((launch_linear<64, true, EPI_NONE>(lin(w.g128, 2 * D, P + NCF_OFF(NCF_P_K_W), D, nullptr, w.g256, D, N, 2 * D), st)));
  }
  w.dxu = w.g64b;
  w.dxp = w.g256;
  return NCF_OK;
}

// loss_zeroed: the caller has already zeroed *loss_out on the stream (ncf_train_step does it before its first kernel, so
// that no memset node interrupts the chain of programmatically dependent tower kernels)
int launch_bce(const float* out, const float* targets, int64_t N, float* loss_out, float* grad_out, cudaStream_t st, bool loss_zeroed) {
  if (!loss_zeroed) NCF_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
  if (N == 0) return NCF_OK;
  const int grid = (int)std::min<int64_t>((N + 255) / 256, (int64_t)num_sms() * 4);
  NCF_CUDA(launch_pdl(PDL_HEAD, bce_kernel, dim3(grid), dim3(256), 0, st, out, targets, N, loss_out, grad_out));
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}


// BCELoss + its gradient + the scalar half of the head's backward in one kernel (see bce_head_bwd_kernel); *loss_out is zeroed
// by the caller on the stream
int launch_bce_head_bwd(const float* out, const float* targets, int64_t N, float* loss_out, const float* dense, float* dense_grad,
                        TowerWs& w, cudaStream_t st) {
  if (N == 0) return NCF_OK;
  const int grid = (int)std::min<int64_t>((N + 255) / 256, (int64_t)num_sms() * 4);
  NCF_CUDA(launch_pdl(PDL_HEAD, bce_head_bwd_kernel, dim3(grid), dim3(256), 0, st, out, targets, N, loss_out, (const float*)w.mf_pred,
                      (const float*)w.mlp_pred, dense, w.d_mf, w.d_mlp, dense_grad));
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

}  // namespace ncf
