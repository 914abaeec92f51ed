// tcgen05 (5th-gen tensor core) kernels: bf16 operands, fp32 accumulation in TMEM.
//   ncf_tc_selftest   one 128-row GEMM tile in the three operand arrangements the towers use
//                     (forward, input-gradient, weight-gradient); unit-tested against fp32.
//   mlp_tc_*          the 3-layer MLP tower (architecture.py:230-242) forward / dgrad / wgrad.
#include "ncf_tower.cuh"
#include "ncf_umma.cuh"

namespace ncf {
using namespace umma;

// Fill a [R x C] bf16 operand tile (canonical layout, see ncf_umma.cuh) from an fp32 row-major source.
// Thread mapping: each quarter warp writes one 128-byte core matrix (8 rows x 16 B, conflict-free) and
// the four quarters of a warp read four adjacent 32-byte chunks of the same 8 rows (full sectors).
template <int C>
__device__ __forceinline__ void fill_tile_f32(uint8_t* tile, const float* __restrict__ src, int64_t ld, int64_t row0,
                                              int64_t rows_avail, int R, int tid, int nthreads) {
  constexpr int CH = C / 8;                       // 16-byte chunks per row
  const int total = R * CH;
  for (int q = tid; q < total; q += nthreads) {
    const int blk = q >> 5, l = q & 31;           // 32 chunks per (8 rows x 4 chunks) block
    const int blocks_per_rowgroup = CH / 4;
    const int rg = blk / blocks_per_rowgroup, cb = blk % blocks_per_rowgroup;
    const int r = rg * 8 + (l & 7), j = cb * 4 + (l >> 3);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < rows_avail) {
      const float* p = src + (row0 + r) * ld + 8 * j;
      const float4 a = ldg4(p), b = ldg4(p + 4);
      v = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
    }
    *reinterpret_cast<uint4*>(tile + tile_off(r, 8 * j, C)) = v;
  }
}
// same, bf16 row-major source
template <int C>
__device__ __forceinline__ void fill_tile_bf16(uint8_t* tile, const __nv_bfloat16* __restrict__ src, int64_t ld,
                                               int64_t row0, int64_t rows_avail, int R, int tid, int nthreads) {
  constexpr int CH = C / 8;
  const int total = R * CH;
  for (int q = tid; q < total; q += nthreads) {
    const int blk = q >> 5, l = q & 31;
    const int blocks_per_rowgroup = CH / 4;
    const int rg = blk / blocks_per_rowgroup, cb = blk % blocks_per_rowgroup;
    const int r = rg * 8 + (l & 7), j = cb * 4 + (l >> 3);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < rows_avail) v = __ldg(reinterpret_cast<const uint4*>(src + (row0 + r) * ld + 8 * j));
    *reinterpret_cast<uint4*>(tile + tile_off(r, 8 * j, C)) = v;
  }
}

// One GEMM = `ksteps` tcgen05.mma of K=16 each.  a_step / b_step: byte advance of the operand start
// address per K step (2 core matrices along K).
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_addr, uint32_t a_lbo, uint32_t a_sbo,
                                           uint32_t a_step, uint32_t b_addr, uint32_t b_lbo, uint32_t b_sbo,
                                           uint32_t b_step, uint32_t idesc, int ksteps, bool accumulate_first) {
  for (int k = 0; k < ksteps; ++k) {
    mma_bf16(tmem_d, make_desc(a_addr + k * a_step, a_lbo, a_sbo), make_desc(b_addr + k * b_step, b_lbo, b_sbo), idesc,
             accumulate_first || k > 0);
  }
}

// ---------------------------------------------------------------------------------------------
// self test: D[128 x N] for one tile
//   mode 0: D = A[128,K] . B[N,K]^T            (A K-major, B K-major)         forward
//   mode 1: D = A[128,K] . B[K,N]              (A K-major, B = [K rows, N cols] read MN-major)  dgrad
//   mode 2: D = A[128r,128]^T . B[128r,N]      (both read MN-major, K = 128 rows)               wgrad
// ---------------------------------------------------------------------------------------------
template <int MODE, int K, int N>
__global__ void __launch_bounds__(128) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                          float* __restrict__ Dout) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int AR = 128, AC = MODE == 2 ? 128 : K;            // A tile rows x cols as stored
  constexpr int BR = MODE == 0 ? N : (MODE == 1 ? K : 128), BC = MODE == 0 ? K : N;
  uint8_t* sA = smem;
  uint8_t* sB = smem + AR * AC * 2;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  fill_tile_f32<AC>(sA, A, AC, 0, AR, AR, tid, 128);
  fill_tile_f32<BC>(sB, B, BC, 0, BR, BR, tid, 128);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t a = smem_addr(sA), b = smem_addr(sB);
    if (MODE == 0) {
      issue_gemm(tmem, a, 128, AC * 16, 256, b, 128, BC * 16, 256, make_idesc(128, N, false, false), K / 16, false);
    } else if (MODE == 1) {
      issue_gemm(tmem, a, 128, AC * 16, 256, b, BC * 16, 128, 2 * BC * 16, make_idesc(128, N, false, true), K / 16, false);
    } else {
      issue_gemm(tmem, a, AC * 16, 128, 2 * AC * 16, b, BC * 16, 128, 2 * BC * 16, make_idesc(128, N, true, true), 128 / 16,
                 false);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) Dout[row * N + c0 + i] = v[i];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}


// =============================================================================================
// MLP tower forward (architecture.py:230-246, 344-354) fused per 128-row tile:
//   a[128,64] -> (W0) -> ReLU/LN/dropout -> (W1) -> ... -> (W2) -> ... -> mlp_output -> final -> sigmoid
// One persistent CTA per SM keeps the three bf16 weight images (112 KB) in shared memory; the
// accumulators live in TMEM (256 + 128 + 64 columns); activations go TMEM -> registers (thread = row,
// so LayerNorm needs no shuffles) -> bf16 -> shared memory as the next layer's A operand.
// 8 warps: warp w reads TMEM lane quarter (w & 3) and column half (w >> 2) of every accumulator.
// =============================================================================================
constexpr int TCM_THREADS = 256;
constexpr int TCM_ROWS = 128;
constexpr uint32_t SM_W0 = 0;                       // [256][64]  bf16  32 KB
constexpr uint32_t SM_W1 = SM_W0 + 256 * 64 * 2;    // [128][256]       64 KB
constexpr uint32_t SM_W2 = SM_W1 + 128 * 256 * 2;   // [64][128]        16 KB
constexpr uint32_t SM_A0 = SM_W2 + 64 * 128 * 2;    // [128][64]        16 KB  input tile
constexpr uint32_t SM_Y = SM_A0 + 128 * 64 * 2;     // [128][256]       64 KB  Y1, later Y2 / dz tiles
constexpr uint32_t SM_PAR = SM_Y + 128 * 256 * 2;   // fp32 parameters
constexpr int PAR_B0 = 0, PAR_G0 = 256, PAR_E0 = 512, PAR_B1 = 768, PAR_G1 = 896, PAR_E1 = 1024, PAR_B2 = 1152,
              PAR_G2 = 1216, PAR_E2 = 1280, PAR_WOUT = 1344, PAR_SCAL = 1408, PAR_COUNT = 1416;
constexpr uint32_t SM_STAT = SM_PAR + PAR_COUNT * 4;           // [128][2][2] floats
constexpr uint32_t SM_HEAD = SM_STAT + 128 * 2 * 2 * 4;        // [128][2] floats
constexpr uint32_t SM_MLP_TOTAL = SM_HEAD + 128 * 2 * 4;

struct MlpFwdArgs {
  const float* a;            // [N,64] fp32 attention output
  const float* dense;
  const int64_t* hour;       // optional (forward_simple hour path)
  const float* tail1;        // [24,256]
  const float* mf_pred;      // [N]
  float *out, *out2, *mlp_pred, *y3;   // [N], [N], [N], [N,64] fp32
  __nv_bfloat16 *r1, *y1, *r2, *y2, *r3;   // saved activations (training) or null
  int64_t N;
  DropoutRng rng[3];
};

// weights fp32 [ROWS, ld] (first COLS columns) -> bf16 canonical image
template <int ROWS, int COLS>
__device__ __forceinline__ void load_weight_image(uint8_t* img, const float* __restrict__ w, int ld, int tid, int nthreads) {
  fill_tile_f32<COLS>(img, w, ld, 0, ROWS, ROWS, tid, nthreads);
}

// Epilogue of one layer for this thread's row and column half.  Pass 1: bias (+tail) + ReLU, bf16
// rounding, row statistics; pass 2: LayerNorm, dropout, bf16 -> next A operand / saved tensors.
template <int C, bool LAST>
__device__ __forceinline__ void mlp_epilogue(uint32_t tmem_acc, int q, int h, int lane, int64_t grow, bool live,
                                             const float* __restrict__ par_b, const float* __restrict__ par_g,
                                             const float* __restrict__ par_e, const float* __restrict__ tail_row,
                                             const DropoutRng& rng, float* s_stat, uint8_t* ytile,
                                             __nv_bfloat16* __restrict__ r_out, __nv_bfloat16* __restrict__ y_out,
                                             float* __restrict__ y3_out, const float* __restrict__ par_wout,
                                             float& head_partial) {
  constexpr int HALF = C / 2, NCH = HALF / 32;
  const int rt = q * 32 + lane;                 // row inside the tile
  const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + h * HALF;
  float sum = 0.f, sq = 0.f;
#pragma unroll 1
  for (int ch = 0; ch < NCH; ++ch) {
    float v[32];
    tmem_ld32(taddr + ch * 32, v);
    const int c0 = h * HALF + ch * 32;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float x = v[i] + par_b[c0 + i];
      if (tail_row) x += __ldg(tail_row + c0 + i);
      x = bf16_round(fmaxf(x, 0.f));
      v[i] = x;
      sum += x;
      sq = fmaf(x, x, sq);
    }
    if (r_out && live) {
      uint4* dst = reinterpret_cast<uint4*>(r_out + grow * C + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        dst[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                            pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
    }
  }
  s_stat[(rt * 2 + h) * 2 + 0] = sum;
  s_stat[(rt * 2 + h) * 2 + 1] = sq;
  __syncthreads();
  sum += s_stat[(rt * 2 + (h ^ 1)) * 2 + 0];
  sq += s_stat[(rt * 2 + (h ^ 1)) * 2 + 1];
  const float mean = sum * (1.0f / C);
  const float rstd = rsqrtf(fmaxf(sq * (1.0f / C) - mean * mean, 0.f) + LN_EPS);
  float hp = 0.f;
#pragma unroll 1
  for (int ch = 0; ch < NCH; ++ch) {
    float v[32];
    tmem_ld32(taddr + ch * 32, v);
    const int c0 = h * HALF + ch * 32;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float x = v[i] + par_b[c0 + i];
      if (tail_row) x += __ldg(tail_row + c0 + i);
      x = bf16_round(fmaxf(x, 0.f));
      v[i] = fmaf((x - mean) * rstd, par_g[c0 + i], par_e[c0 + i]);
    }
    if (rng.thresh != 0u) {
#pragma unroll
      for (int g4 = 0; g4 < 8; ++g4) {
        const uint4 rnd = rng.draw4(((uint64_t)grow * C + c0 + 4 * g4) >> 2);
        v[4 * g4 + 0] = rnd.x >= rng.thresh ? v[4 * g4 + 0] * rng.scale : 0.f;
        v[4 * g4 + 1] = rnd.y >= rng.thresh ? v[4 * g4 + 1] * rng.scale : 0.f;
        v[4 * g4 + 2] = rnd.z >= rng.thresh ? v[4 * g4 + 2] * rng.scale : 0.f;
        v[4 * g4 + 3] = rnd.w >= rng.thresh ? v[4 * g4 + 3] * rng.scale : 0.f;
      }
    }
    if (LAST) {
#pragma unroll
      for (int i = 0; i < 32; ++i) hp = fmaf(v[i], par_wout[c0 + i], hp);
      if (y3_out && live) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st4(y3_out + grow * C + c0 + 4 * j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 pk = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                    pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
        *reinterpret_cast<uint4*>(ytile + tile_off(rt, c0 + 8 * j, C)) = pk;
        if (y_out && live) *reinterpret_cast<uint4*>(y_out + grow * C + c0 + 8 * j) = pk;
      }
    }
  }
  head_partial = hp;
}

__global__ void __launch_bounds__(TCM_THREADS, 1) mlp_tc_fwd_kernel(MlpFwdArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, h = warp >> 2;
  float* par = reinterpret_cast<float*>(smem + SM_PAR);
  float* s_stat = reinterpret_cast<float*>(smem + SM_STAT);
  float* s_head = reinterpret_cast<float*>(smem + SM_HEAD);
  const float* P = A.dense;

  load_weight_image<256, 64>(smem + SM_W0, P + NCF_OFF(NCF_P_MLP0_W), K0, tid, TCM_THREADS);
  load_weight_image<128, 256>(smem + SM_W1, P + NCF_OFF(NCF_P_MLP1_W), H1, tid, TCM_THREADS);
  load_weight_image<64, 128>(smem + SM_W2, P + NCF_OFF(NCF_P_MLP2_W), H2, tid, TCM_THREADS);
  for (int i = tid; i < 256; i += TCM_THREADS) {
    par[PAR_B0 + i] = P[NCF_OFF(NCF_P_MLP0_B) + i];
    par[PAR_G0 + i] = P[NCF_OFF(NCF_P_LN0_W) + i];
    par[PAR_E0 + i] = P[NCF_OFF(NCF_P_LN0_B) + i];
    if (i < 128) {
      par[PAR_B1 + i] = P[NCF_OFF(NCF_P_MLP1_B) + i];
      par[PAR_G1 + i] = P[NCF_OFF(NCF_P_LN1_W) + i];
      par[PAR_E1 + i] = P[NCF_OFF(NCF_P_LN1_B) + i];
    }
    if (i < 64) {
      par[PAR_B2 + i] = P[NCF_OFF(NCF_P_MLP2_B) + i];
      par[PAR_G2 + i] = P[NCF_OFF(NCF_P_LN2_W) + i];
      par[PAR_E2 + i] = P[NCF_OFF(NCF_P_LN2_B) + i];
      par[PAR_WOUT + i] = P[NCF_OFF(NCF_P_MLP_OUT_W) + i];
    }
  }
  if (tid == 0) {
    par[PAR_SCAL + 0] = P[NCF_OFF(NCF_P_MLP_OUT_B)];
    par[PAR_SCAL + 1] = P[NCF_OFF(NCF_P_FINAL_W)];
    par[PAR_SCAL + 2] = P[NCF_OFF(NCF_P_FINAL_W) + 1];
    par[PAR_SCAL + 3] = P[NCF_OFF(NCF_P_FINAL_B)];
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t sW0 = smem_addr(smem + SM_W0), sW1 = smem_addr(smem + SM_W1), sW2 = smem_addr(smem + SM_W2);
  const uint32_t sA0 = smem_addr(smem + SM_A0), sY = smem_addr(smem + SM_Y);
  uint32_t phase = 0;

  const int64_t ntiles = (A.N + TCM_ROWS - 1) / TCM_ROWS;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * TCM_ROWS;
    const int64_t avail = min((int64_t)TCM_ROWS, A.N - row0);
    const int rt = q * 32 + lane;
    const int64_t grow = row0 + rt;
    const bool live = rt < avail;
    fill_tile_f32<64>(smem + SM_A0, A.a, D, row0, avail, TCM_ROWS, tid, TCM_THREADS);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_gemm(tmem + 0, sA0, 128, 64 * 16, 256, sW0, 128, 64 * 16, 256, make_idesc(128, 256, false, false), 4, false);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();
    const float* tail_row = (A.hour && live) ? A.tail1 + A.hour[grow] * H1 : nullptr;
    float hp;
    mlp_epilogue<256, false>(tmem + 0, q, h, lane, grow, live, par + PAR_B0, par + PAR_G0, par + PAR_E0, tail_row, A.rng[0],
                             s_stat, smem + SM_Y, A.r1, A.y1, nullptr, nullptr, hp);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_gemm(tmem + 256, sY, 128, 256 * 16, 256, sW1, 128, 256 * 16, 256, make_idesc(128, 128, false, false), 16, false);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();
    mlp_epilogue<128, false>(tmem + 256, q, h, lane, grow, live, par + PAR_B1, par + PAR_G1, par + PAR_E1, nullptr, A.rng[1],
                             s_stat, smem + SM_Y, A.r2, A.y2, nullptr, nullptr, hp);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_gemm(tmem + 384, sY, 128, 128 * 16, 256, sW2, 128, 128 * 16, 256, make_idesc(128, 64, false, false), 8, false);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();
    mlp_epilogue<64, true>(tmem + 384, q, h, lane, grow, live, par + PAR_B2, par + PAR_G2, par + PAR_E2, nullptr, A.rng[2],
                           s_stat, nullptr, A.r3, nullptr, A.y3, par + PAR_WOUT, hp);
    s_head[rt * 2 + h] = hp;
    fence_before_sync();
    __syncthreads();
    if (h == 0 && live) {
      const float mp = s_head[rt * 2] + s_head[rt * 2 + 1] + par[PAR_SCAL + 0];
      const float z = fmaf(par[PAR_SCAL + 1], A.mf_pred[grow], fmaf(par[PAR_SCAL + 2], mp, par[PAR_SCAL + 3]));
      const float pr = 1.0f / (1.0f + expf(-z));
      A.mlp_pred[grow] = mp;
      A.out[grow] = pr;
      if (A.out2) A.out2[grow] = pr;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int launch_mlp_tc_fwd(const MlpFwdArgs& A, cudaStream_t st) {
  if (A.N == 0) return NCF_OK;
  static bool configured = false;
  if (!configured) {
    NCF_CUDA(cudaFuncSetAttribute(mlp_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_MLP_TOTAL));
    configured = true;
  }
  const int64_t ntiles = (A.N + TCM_ROWS - 1) / TCM_ROWS;
  const int grid = (int)std::min<int64_t>(ntiles, num_sms());
  mlp_tc_fwd_kernel<<<grid, TCM_THREADS, SM_MLP_TOTAL, st>>>(A);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

int mlp_tc_forward(const ncf_run_cfg& cfg, const float* dense, int64_t N, const int64_t* hour, const float* tail1,
                   float* out, TowerWs& w, cudaStream_t st) {
  MlpFwdArgs A{};
  const bool train = cfg.training != 0;
  A.a = w.a;
  A.dense = dense;
  A.hour = hour;
  A.tail1 = tail1;
  A.mf_pred = w.mf_pred;
  A.out = out;
  A.out2 = w.p_saved;
  A.mlp_pred = w.mlp_pred;
  A.y3 = w.y3;
  A.r1 = train ? (__nv_bfloat16*)w.r1b : nullptr;
  A.y1 = train ? (__nv_bfloat16*)w.y1b : nullptr;
  A.r2 = train ? (__nv_bfloat16*)w.r2b : nullptr;
  A.y2 = train ? (__nv_bfloat16*)w.y2b : nullptr;
  A.r3 = train ? (__nv_bfloat16*)w.r3b : nullptr;
  A.N = N;
  for (int l = 0; l < 3; ++l) A.rng[l] = make_rng(cfg, 1 + l);
  return launch_mlp_tc_fwd(A, st);
}

int mlp_tc_backward(const ncf_run_cfg&, const float*, float*, int64_t, TowerWs&, cudaStream_t) {
  set_error("tcgen05 MLP backward is not built yet");
  return NCF_ERR_UNSUPPORTED;
}

}  // namespace ncf

using namespace ncf;

template <int MODE, int K, int N>
static int run_selftest(const float* A, const float* B, float* D, cudaStream_t st) {
  constexpr int AC = MODE == 2 ? 128 : K;
  constexpr int BR = MODE == 0 ? N : (MODE == 1 ? K : 128), BC = MODE == 0 ? K : N;
  const int smem = 128 * AC * 2 + BR * BC * 2;
  NCF_CUDA(cudaFuncSetAttribute(tc_selftest_kernel<MODE, K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc_selftest_kernel<MODE, K, N><<<1, 128, smem, st>>>(A, B, D);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

// A, B, D: fp32 device buffers; shapes by mode (see kernel).  (mode, K, N) in {(0,64,256),(0,256,128),
// (1,64,128),(1,128,256),(2,128,64),(2,128,256)}.
extern "C" int ncf_tc_selftest(int32_t mode, int32_t K, int32_t N, const float* A, const float* B, float* D, void* stream) {
  NCF_REQUIRE(A && B && D, "tc_selftest: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0 && K == 64 && N == 256) return run_selftest<0, 64, 256>(A, B, D, st);
  if (mode == 0 && K == 256 && N == 128) return run_selftest<0, 256, 128>(A, B, D, st);
  if (mode == 1 && K == 64 && N == 128) return run_selftest<1, 64, 128>(A, B, D, st);
  if (mode == 1 && K == 128 && N == 256) return run_selftest<1, 128, 256>(A, B, D, st);
  if (mode == 2 && K == 128 && N == 64) return run_selftest<2, 128, 64>(A, B, D, st);
  if (mode == 2 && K == 128 && N == 256) return run_selftest<2, 128, 256>(A, B, D, st);
  set_error("tc_selftest: unsupported (mode,K,N) = (%d,%d,%d)", mode, K, N);
  return NCF_ERR_UNSUPPORTED;
}
