// tcgen05 (5th-gen tensor core) kernels: bf16 operands, fp32 accumulation in TMEM.
//   ncf_tc_selftest   one 128-row GEMM tile in the three operand arrangements the towers use
//                     (forward, input-gradient, weight-gradient); unit-tested against fp32.
//   mlp_tc_*          the 3-layer MLP tower (architecture.py:230-242) forward / dgrad / wgrad.
#include <cstdlib>

#include "ncf_tower.cuh"
#include "ncf_umma.cuh"

namespace ncf {
using namespace umma;

// ---------------------------------------------------------------------------------------------
// self test: D[128 x N] for one tile
//   mode 0: D = A[128,K] . B[N,K]^T            (A K-major, B K-major)         forward
//   mode 1: D = A[128,K] . B[K,N]              (A K-major, B = [K rows, N cols] read MN-major)  dgrad
//   mode 2: D = A[128r,128]^T . B[128r,N]      (both read MN-major, K = 128 rows)               wgrad
//   mode 3: as mode 0 with A staged in TENSOR MEMORY (packed bf16 pairs written by tcgen05.st)   MLP forward
// ---------------------------------------------------------------------------------------------
template <int MODE, int K, int N>
__global__ void __launch_bounds__(128) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                          float* __restrict__ Dout) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int AR = 128, AC = MODE == 2 ? 128 : K;            // A tile rows x cols as stored
  constexpr int BR = (MODE == 0 || MODE == 3) ? N : (MODE == 1 ? K : 128), BC = (MODE == 0 || MODE == 3) ? K : N;
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
  if (MODE == 3) {      // A operand through tensor memory: thread = row writes its K bf16 values as packed pairs
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < K; c0 += 32) {
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = pack_bf16(A[row * K + c0 + 2 * i], A[row * K + c0 + 2 * i + 1]);
      tmem_st16u(tmem + 128 + ((uint32_t)(warp * 32) << 16) + c0 / 2, w);
    }
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }
  if (tid == 0) {
    const uint32_t a = smem_addr(sA), b = smem_addr(sB);
    if (MODE == 3) {
      issue_gemm_ts(tmem, tmem + 128, b, 128, BC * 16, 256, make_idesc(128, N, false, false), K / 16, false);
    } else if (MODE == 0) {
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
constexpr int MLP_THREADS = 512;                               // 16 warps: 4 TMEM lane quarters x 4 column parts
constexpr int MLP_NH = 4;
constexpr uint32_t SM_STAT = SM_PAR + PAR_COUNT * 4;           // [128][4][2] floats
constexpr uint32_t SM_HEAD = SM_STAT + 128 * MLP_NH * 2 * 4;   // [128][4] floats
constexpr uint32_t SM_A1 = (SM_HEAD + 128 * MLP_NH * 4 + 1023) / 1024 * 1024;   // second input-tile buffer (TMA double buffering)
constexpr uint32_t SM_MLP_TOTAL = SM_A1 + 128 * 64 * 2;

struct MlpFwdArgs {
  const __nv_bfloat16* a_img; // attention output, bf16 tile image
  float *st1, *st2, *st3;    // LayerNorm (mean, rstd) per row, saved for the backward (or null)
  const float* dense;
  const int64_t* hour;       // optional (forward_simple hour path)
  const float* tail1;        // [24,256]
  const float* mf_pred;      // [N]
  float *out, *out2, *mlp_pred, *y3;   // [N], [N], [N], [N,64] fp32
  __nv_bfloat16 *r1, *y1, *r2, *y2, *r3;   // saved activations (training) or null
  int64_t N;
  DropoutRng rng[3];
};

// named barrier of the 4 warps (128 threads) that own TMEM lane quarter q in the 16-warp MLP kernels
__device__ __forceinline__ void quarter_sync(int q) { asm volatile("bar.sync %0, 128;" ::"r"(q + 1) : "memory"); }

// weights fp32 [ROWS, ld] (first COLS columns) -> bf16 canonical image
template <int ROWS, int COLS>
__device__ __forceinline__ void load_weight_image(uint8_t* img, const float* __restrict__ w, int ld, int tid, int nthreads) {
  fill_tile_f32<COLS>(img, w, ld, 0, ROWS, ROWS, tid, nthreads);
}

// Saved activations live in global memory as TILE IMAGES: tile t of a [N,C] tensor is the 128 x C
// canonical operand image at byte offset t*128*C*2 (rows beyond N are padding).  A warp store of one
// 16-byte chunk per lane then covers 4 x 128 contiguous bytes (full sectors), and the consumers
// (backward, weight gradient) copy tiles linearly.
// Epilogue of one layer for this thread's row and column part (C/4 columns).
// Pass 1: bias (+tail) + ReLU, bf16 rounding (packed pairs), row statistics, dropout decisions.  The packed
// ReLU outputs stay in registers; a DROPPED element carries its decision in the (otherwise unused) sign bit,
// and that is also the form the saved r tensor takes, so the backward needs no random numbers.
// Pass 2: LayerNorm (gamma/beta pre-multiplied by 1/(1-p)), dropout select, bf16 -> next A operand / saved y.
template <int C, bool LAST, bool HAS_TAIL, bool Y_TMEM = false>
__device__ __forceinline__ void mlp_epilogue(uint32_t tmem_acc, int q, int h, int lane, int64_t grow, bool live,
                                             const float* __restrict__ par_b, const float* __restrict__ par_g,
                                             const float* __restrict__ par_e, const float* __restrict__ tail_row,
                                             const DropoutRng& rng, float* s_stat, uint8_t* ytile,
                                             uint8_t* __restrict__ r_img, uint8_t* __restrict__ y_img,
                                             float* __restrict__ y3_out, const float* __restrict__ par_wout,
                                             float& head_partial, float* __restrict__ st_tile, uint32_t tmem_y = 0) {
  constexpr int PART = C / MLP_NH, CW = PART >= 32 ? 32 : 16, NCH = PART / CW;
  const int rt = q * 32 + lane;                 // row inside the tile
  const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + h * PART;
  // Y_TMEM: the layer output becomes the next GEMM's A operand IN TENSOR MEMORY (packed bf16 pairs, lane = row)
  const uint32_t yaddr = tmem_y + ((uint32_t)(q * 32) << 16) + (h * PART) / 2;
  uint32_t yw[16];
  uint32_t pk[PART / 2];
  float sum4[2] = {0.f, 0.f}, sq4[2] = {0.f, 0.f};     // independent chains: the adds do not serialise
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    float v[CW];
    tmem_ldw<CW>(taddr + ch * CW, v);
    const int c0 = h * PART + ch * CW;
#pragma unroll
    for (int i4 = 0; i4 < CW / 4; ++i4) {
      const float4 b4 = *reinterpret_cast<const float4*>(par_b + c0 + 4 * i4);
      float4 t4 = make_float4(0, 0, 0, 0);
      if (HAS_TAIL) t4 = tail_row ? ldg4(tail_row + c0 + 4 * i4) : t4;
      const float bb[4] = {b4.x + t4.x, b4.y + t4.y, b4.z + t4.z, b4.w + t4.w};
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {
        const uint32_t w = pack_bf16(fmaxf(v[4 * i4 + 2 * k2] + bb[2 * k2], 0.f), fmaxf(v[4 * i4 + 2 * k2 + 1] + bb[2 * k2 + 1], 0.f));
        const float x0 = __uint_as_float(w << 16), x1 = __uint_as_float(w & 0xffff0000u);
        sum4[k2] += x0 + x1;
        sq4[k2] = fmaf(x0, x0, fmaf(x1, x1, sq4[k2]));
        pk[ch * (CW / 2) + i4 * 2 + k2] = w;
      }
    }
#pragma unroll
    for (int g8 = 0; g8 < CW / 8; ++g8) {
      uint32_t* w4 = &pk[ch * (CW / 2) + 4 * g8];
      if (rng.thresh != 0u) {
        const uint4 r = rng.draw8(((uint64_t)grow * C + c0 + 8 * g8) >> 3);
        w4[0] |= rng.drop_signs(r.x);
        w4[1] |= rng.drop_signs(r.y);
        w4[2] |= rng.drop_signs(r.z);
        w4[3] |= rng.drop_signs(r.w);
      }
      if (r_img) *reinterpret_cast<uint4*>(r_img + tile_off(rt, c0 + 8 * g8, C)) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
  }
  float sum = sum4[0] + sum4[1], sq = sq4[0] + sq4[1];
  s_stat[(rt * MLP_NH + h) * 2 + 0] = sum;
  s_stat[(rt * MLP_NH + h) * 2 + 1] = sq;
  quarter_sync(q);          // only the four warps that share these 32 rows exchange statistics
  sum = 0.f;
  sq = 0.f;
#pragma unroll
  for (int k = 0; k < MLP_NH; ++k) {
    sum += s_stat[(rt * MLP_NH + k) * 2 + 0];
    sq += s_stat[(rt * MLP_NH + k) * 2 + 1];
  }
  const float mean = sum * (1.0f / C);
  const float rstd = rsqrtf(fmaxf(sq * (1.0f / C) - mean * mean, 0.f) + LN_EPS);
  const float nmr = -mean * rstd;
  if (st_tile && h == 0) *reinterpret_cast<float2*>(st_tile + rt * 2) = make_float2(mean, rstd);
  float hp = 0.f;
#pragma unroll
  for (int g8 = 0; g8 < PART / 8; ++g8) {
    const int c0 = h * PART + 8 * g8;
    float y[8];
#pragma unroll
    for (int i4 = 0; i4 < 2; ++i4) {
      const float4 g4 = *reinterpret_cast<const float4*>(par_g + c0 + 4 * i4);
      const float4 e4 = *reinterpret_cast<const float4*>(par_e + c0 + 4 * i4);
      const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, ee[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {
        const uint32_t w = pk[4 * g8 + 2 * i4 + k2];
        const uint32_t lo = w << 16;
        const float x0 = fabsf(__uint_as_float(lo)), x1 = fabsf(__uint_as_float(w & 0xffff0000u));
        const float y0 = fmaf(fmaf(x0, rstd, nmr), gg[2 * k2], ee[2 * k2]);
        const float y1 = fmaf(fmaf(x1, rstd, nmr), gg[2 * k2 + 1], ee[2 * k2 + 1]);
        y[4 * i4 + 2 * k2] = (int32_t)lo < 0 ? 0.f : y0;
        y[4 * i4 + 2 * k2 + 1] = (int32_t)w < 0 ? 0.f : y1;
      }
    }
    if (LAST) {
#pragma unroll
      for (int i4 = 0; i4 < 2; ++i4) {
        const float4 w4 = *reinterpret_cast<const float4*>(par_wout + c0 + 4 * i4);
        hp = fmaf(y[4 * i4], w4.x, fmaf(y[4 * i4 + 1], w4.y, fmaf(y[4 * i4 + 2], w4.z, fmaf(y[4 * i4 + 3], w4.w, hp))));
      }
      if (y3_out && live) {
        st4(y3_out + grow * C + c0, make_float4(y[0], y[1], y[2], y[3]));
        st4(y3_out + grow * C + c0 + 4, make_float4(y[4], y[5], y[6], y[7]));
      }
    } else {
      const uint4 o = make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
      const uint32_t off = tile_off(rt, c0, C);
      if (Y_TMEM) {
        yw[4 * (g8 & 3)] = o.x; yw[4 * (g8 & 3) + 1] = o.y; yw[4 * (g8 & 3) + 2] = o.z; yw[4 * (g8 & 3) + 3] = o.w;
        if ((g8 & 3) == 3) tmem_st16u(yaddr + (g8 >> 2) * 16, yw);
      } else {
        *reinterpret_cast<uint4*>(ytile + off) = o;
      }
      if (y_img) *reinterpret_cast<uint4*>(y_img + off) = o;
    }
  }
  if (Y_TMEM && !LAST) tmem_st_wait();
  head_partial = hp;
}

__global__ void __launch_bounds__(MLP_THREADS, 1) mlp_tc_fwd_kernel(MlpFwdArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar, bar1;
  __shared__ __align__(8) uint64_t full[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, h = warp >> 2;
  float* par = reinterpret_cast<float*>(smem + SM_PAR);
  float* s_stat = reinterpret_cast<float*>(smem + SM_STAT);
  float* s_head = reinterpret_cast<float*>(smem + SM_HEAD);
  const float* P = A.dense;

  load_weight_image<256, 64>(smem + SM_W0, P + NCF_OFF(NCF_P_MLP0_W), K0, tid, MLP_THREADS);
  load_weight_image<128, 256>(smem + SM_W1, P + NCF_OFF(NCF_P_MLP1_W), H1, tid, MLP_THREADS);
  load_weight_image<64, 128>(smem + SM_W2, P + NCF_OFF(NCF_P_MLP2_W), H2, tid, MLP_THREADS);
  // LayerNorm gamma / beta carry the dropout scale 1/(1-p): y = keep ? LN(x) * scale : 0
  const float sc0 = A.rng[0].thresh ? A.rng[0].scale : 1.f, sc1 = A.rng[1].thresh ? A.rng[1].scale : 1.f,
              sc2 = A.rng[2].thresh ? A.rng[2].scale : 1.f;
  for (int i = tid; i < 256; i += MLP_THREADS) {
    par[PAR_B0 + i] = P[NCF_OFF(NCF_P_MLP0_B) + i];
    par[PAR_G0 + i] = P[NCF_OFF(NCF_P_LN0_W) + i] * sc0;
    par[PAR_E0 + i] = P[NCF_OFF(NCF_P_LN0_B) + i] * sc0;
    if (i < 128) {
      par[PAR_B1 + i] = P[NCF_OFF(NCF_P_MLP1_B) + i];
      par[PAR_G1 + i] = P[NCF_OFF(NCF_P_LN1_W) + i] * sc1;
      par[PAR_E1 + i] = P[NCF_OFF(NCF_P_LN1_B) + i] * sc1;
    }
    if (i < 64) {
      par[PAR_B2 + i] = P[NCF_OFF(NCF_P_MLP2_B) + i];
      par[PAR_G2 + i] = P[NCF_OFF(NCF_P_LN2_W) + i] * sc2;
      par[PAR_E2 + i] = P[NCF_OFF(NCF_P_LN2_B) + i] * sc2;
      par[PAR_WOUT + i] = P[NCF_OFF(NCF_P_MLP_OUT_W) + i];
    }
  }
  if (tid == 0) {
    par[PAR_SCAL + 0] = P[NCF_OFF(NCF_P_MLP_OUT_B)];
    par[PAR_SCAL + 1] = P[NCF_OFF(NCF_P_FINAL_W)];
    par[PAR_SCAL + 2] = P[NCF_OFF(NCF_P_FINAL_W) + 1];
    par[PAR_SCAL + 3] = P[NCF_OFF(NCF_P_FINAL_B)];
    mbar_init(&bar, 1);
    mbar_init(&bar1, 1);
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t sW0 = smem_addr(smem + SM_W0), sW1 = smem_addr(smem + SM_W1), sW2 = smem_addr(smem + SM_W2);
  const uint32_t sY = smem_addr(smem + SM_Y);
  uint32_t phase = 0;

  const int64_t ntiles = (A.N + TCM_ROWS - 1) / TCM_ROWS;
  constexpr uint32_t A_BYTES = 128 * 64 * 2;
  const uint8_t* a_img = reinterpret_cast<const uint8_t*>(A.a_img);
  // Layer-1 GEMMs run ONE TILE AHEAD: the accumulator columns [0,256) are free as soon as epilogue 1 of the
  // current tile is done, so the next tile's GEMM is issued together with the current tile's layer-2 GEMM
  // and its latency (and the TMA load feeding it) hides behind epilogues 2 and 3.
  auto issue_l1 = [&](int b) {
    issue_gemm(tmem + 0, smem_addr(smem + (b ? SM_A1 : SM_A0)), 128, 64 * 16, 256, sW0, 128, 64 * 16, 256,
               make_idesc(128, 256, false, false), 4, false);
    mma_commit(&bar1);
  };
  if (tid == 0 && (int64_t)blockIdx.x < ntiles) {          // TMA bulk copies of the first two input tiles
    mbar_arrive_expect_tx(&full[0], A_BYTES);
    bulk_g2s(smem + SM_A0, a_img + (int64_t)blockIdx.x * A_BYTES, A_BYTES, &full[0]);
    const int64_t second = (int64_t)blockIdx.x + gridDim.x;
    if (second < ntiles) {
      mbar_arrive_expect_tx(&full[1], A_BYTES);
      bulk_g2s(smem + SM_A1, a_img + second * A_BYTES, A_BYTES, &full[1]);
    }
    mbar_wait(&full[0], 0);
    fence_after_sync();
    issue_l1(0);
  }
  int it = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    const int64_t row0 = tile * TCM_ROWS;
    const int64_t avail = min((int64_t)TCM_ROWS, A.N - row0);
    const int rt = q * 32 + lane;
    const int64_t grow = row0 + rt;
    const bool live = rt < avail;
    if (warp == 0) mbar_wait(&bar1, it & 1);    // one warp polls the mbarrier, the rest park on the CTA barrier
    __syncthreads();
    fence_after_sync();
    float hp;
    uint8_t* r1i = A.r1 ? reinterpret_cast<uint8_t*>(A.r1) + tile * (128 * 256 * 2) : nullptr;
    uint8_t* y1i = A.y1 ? reinterpret_cast<uint8_t*>(A.y1) + tile * (128 * 256 * 2) : nullptr;
    uint8_t* r2i = A.r2 ? reinterpret_cast<uint8_t*>(A.r2) + tile * (128 * 128 * 2) : nullptr;
    uint8_t* y2i = A.y2 ? reinterpret_cast<uint8_t*>(A.y2) + tile * (128 * 128 * 2) : nullptr;
    uint8_t* r3i = A.r3 ? reinterpret_cast<uint8_t*>(A.r3) + tile * (128 * 64 * 2) : nullptr;
    float* st1 = A.st1 ? A.st1 + tile * 256 : nullptr;
    float* st2 = A.st2 ? A.st2 + tile * 256 : nullptr;
    float* st3 = A.st3 ? A.st3 + tile * 256 : nullptr;
    if (A.hour) {
      const float* tail_row = live ? A.tail1 + clamp_id(A.hour[grow], 24) * H1 : nullptr;
      mlp_epilogue<256, false, true>(tmem + 0, q, h, lane, grow, live, par + PAR_B0, par + PAR_G0, par + PAR_E0, tail_row,
                                     A.rng[0], s_stat, smem + SM_Y, r1i, y1i, nullptr, nullptr, hp, st1);
    } else {
      mlp_epilogue<256, false, false>(tmem + 0, q, h, lane, grow, live, par + PAR_B0, par + PAR_G0, par + PAR_E0, nullptr,
                                      A.rng[0], s_stat, smem + SM_Y, r1i, y1i, nullptr, nullptr, hp, st1);
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_gemm(tmem + 256, sY, 128, 256 * 16, 256, sW1, 128, 256 * 16, 256, make_idesc(128, 128, false, false), 16, false);
      mma_commit(&bar);
      const int64_t next = tile + gridDim.x;
      if (next < ntiles) {
        mbar_wait(&full[buf ^ 1], ((it + 1) >> 1) & 1);
        fence_after_sync();
        issue_l1(buf ^ 1);
        const int64_t next2 = next + gridDim.x;
        if (next2 < ntiles) {                               // this tile's input buffer is free: layer-1 GEMM done
          mbar_arrive_expect_tx(&full[buf], A_BYTES);
          bulk_g2s(smem + (buf ? SM_A1 : SM_A0), a_img + next2 * A_BYTES, A_BYTES, &full[buf]);
        }
      }
    }
    if (warp == 0) mbar_wait(&bar, phase);      // one warp polls the mbarrier, the rest park on the CTA barrier
    phase ^= 1;
    __syncthreads();
    fence_after_sync();
    mlp_epilogue<128, false, false>(tmem + 256, q, h, lane, grow, live, par + PAR_B1, par + PAR_G1, par + PAR_E1, nullptr,
                                    A.rng[1], s_stat, smem + SM_Y, r2i, y2i, nullptr, nullptr, hp, st2);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_gemm(tmem + 384, sY, 128, 128 * 16, 256, sW2, 128, 128 * 16, 256, make_idesc(128, 64, false, false), 8, false);
      mma_commit(&bar);
    }
    if (warp == 0) mbar_wait(&bar, phase);      // one warp polls the mbarrier, the rest park on the CTA barrier
    phase ^= 1;
    __syncthreads();
    fence_after_sync();
    mlp_epilogue<64, true, false>(tmem + 384, q, h, lane, grow, live, par + PAR_B2, par + PAR_G2, par + PAR_E2, nullptr,
                                  A.rng[2], s_stat, nullptr, r3i, nullptr, A.y3, par + PAR_WOUT, hp, st3);
    s_head[rt * MLP_NH + h] = hp;
    fence_before_sync();
    __syncthreads();
    if (h == 0 && live) {
      const float mp = (s_head[rt * MLP_NH] + s_head[rt * MLP_NH + 1]) + (s_head[rt * MLP_NH + 2] + s_head[rt * MLP_NH + 3]) +
                       par[PAR_SCAL + 0];
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

// ---------------------------------------------------------------------------------------------
// Two tiles in flight per CTA.  Each tile owns a 256-column TMEM slot; the layer outputs y1 / y2 are written
// back to the slot as packed bf16 and consumed by the next GEMM as its A operand straight from tensor memory,
// so no activation tile lives in shared memory and the two slots never compete for it:
//   [0,256) layer-1 accumulator -> [0,128) y1 bf16 | [128,256) layer-2 accumulator -> [0,64) y2 bf16 |
//   [64,128) layer-3 accumulator.
// All 16 warps work on ONE epilogue at a time, alternating between the slots, and every GEMM is issued one
// epilogue ahead of its consumer: the tensor-core round trip of a tile hides behind the other tile's epilogue.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t SM2_A0 = SM_W2 + 64 * 128 * 2;               // input tile of slot 0 (16 KB)
constexpr uint32_t SM2_A1 = SM2_A0 + 128 * 64 * 2;              // input tile of slot 1
constexpr uint32_t SM2_PAR = SM2_A1 + 128 * 64 * 2;
constexpr uint32_t SM2_STAT = SM2_PAR + PAR_COUNT * 4;
constexpr uint32_t SM2_HEAD = SM2_STAT + 128 * MLP_NH * 2 * 4;
constexpr uint32_t SM2_TOTAL = SM2_HEAD + 128 * MLP_NH * 4;

__global__ void __launch_bounds__(MLP_THREADS, 1) mlp_tc_fwd2_kernel(MlpFwdArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[2], bl1[2], bl2[2], bl3[2];
  __shared__ uint32_t tmem_slot;
  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, h = warp >> 2;
  float* par = reinterpret_cast<float*>(smem + SM2_PAR);
  float* s_stat = reinterpret_cast<float*>(smem + SM2_STAT);
  float* s_head = reinterpret_cast<float*>(smem + SM2_HEAD);
  const float* P = A.dense;

  load_weight_image<256, 64>(smem + SM_W0, P + NCF_OFF(NCF_P_MLP0_W), K0, tid, MLP_THREADS);
  load_weight_image<128, 256>(smem + SM_W1, P + NCF_OFF(NCF_P_MLP1_W), H1, tid, MLP_THREADS);
  load_weight_image<64, 128>(smem + SM_W2, P + NCF_OFF(NCF_P_MLP2_W), H2, tid, MLP_THREADS);
  // LayerNorm gamma / beta carry the dropout scale 1/(1-p): y = keep ? LN(x) * scale : 0
  const float sc0 = A.rng[0].thresh ? A.rng[0].scale : 1.f, sc1 = A.rng[1].thresh ? A.rng[1].scale : 1.f,
              sc2 = A.rng[2].thresh ? A.rng[2].scale : 1.f;
  for (int i = tid; i < 256; i += MLP_THREADS) {
    par[PAR_B0 + i] = P[NCF_OFF(NCF_P_MLP0_B) + i];
    par[PAR_G0 + i] = P[NCF_OFF(NCF_P_LN0_W) + i] * sc0;
    par[PAR_E0 + i] = P[NCF_OFF(NCF_P_LN0_B) + i] * sc0;
    if (i < 128) {
      par[PAR_B1 + i] = P[NCF_OFF(NCF_P_MLP1_B) + i];
      par[PAR_G1 + i] = P[NCF_OFF(NCF_P_LN1_W) + i] * sc1;
      par[PAR_E1 + i] = P[NCF_OFF(NCF_P_LN1_B) + i] * sc1;
    }
    if (i < 64) {
      par[PAR_B2 + i] = P[NCF_OFF(NCF_P_MLP2_B) + i];
      par[PAR_G2 + i] = P[NCF_OFF(NCF_P_LN2_W) + i] * sc2;
      par[PAR_E2 + i] = P[NCF_OFF(NCF_P_LN2_B) + i] * sc2;
      par[PAR_WOUT + i] = P[NCF_OFF(NCF_P_MLP_OUT_W) + i];
    }
  }
  if (tid == 0) {
    par[PAR_SCAL + 0] = P[NCF_OFF(NCF_P_MLP_OUT_B)];
    par[PAR_SCAL + 1] = P[NCF_OFF(NCF_P_FINAL_W)];
    par[PAR_SCAL + 2] = P[NCF_OFF(NCF_P_FINAL_W) + 1];
    par[PAR_SCAL + 3] = P[NCF_OFF(NCF_P_FINAL_B)];
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&bl1[s], 1);
      mbar_init(&bl2[s], 1);
      mbar_init(&bl3[s], 1);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  pdl_wait();      // prologue done: from here on the kernel reads what the previous kernels of the step wrote
  const uint32_t tmem = tmem_slot;
  const uint32_t sW0 = smem_addr(smem + SM_W0), sW1 = smem_addr(smem + SM_W1), sW2 = smem_addr(smem + SM_W2);
  const int64_t ntiles = (A.N + TCM_ROWS - 1) / TCM_ROWS;
  const int64_t nk = (int64_t)blockIdx.x < ntiles ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;   // my tiles
  constexpr uint32_t A_BYTES = 128 * 64 * 2;
  const uint8_t* a_img = reinterpret_cast<const uint8_t*>(A.a_img);
  const int rt = q * 32 + lane;

  auto tile_of = [&](int64_t k) { return (int64_t)blockIdx.x + k * gridDim.x; };
  auto load_input = [&](int s, int64_t k) {          // tid 0: TMA bulk copy of tile k's input into slot s
    mbar_arrive_expect_tx(&full[s], A_BYTES);
    bulk_g2s(smem + (s ? SM2_A1 : SM2_A0), a_img + tile_of(k) * A_BYTES, A_BYTES, &full[s]);
  };
  auto issue_l1 = [&](int s) {                        // tid 0
    issue_gemm(tmem + 256 * s, smem_addr(smem + (s ? SM2_A1 : SM2_A0)), 128, 64 * 16, 256, sW0, 128, 64 * 16, 256,
               make_idesc(128, 256, false, false), 4, false);
    mma_commit(&bl1[s]);
  };
  auto wait_sync = [&](uint64_t* b, uint32_t parity) {
    mbar_wait(b, parity);                              // every warp polls for itself: no CTA barrier just to wait
    fence_after_sync();
  };

  if (tid == 0 && nk > 0) {
    load_input(0, 0);
    if (nk > 1) load_input(1, 1);
    mbar_wait(&full[0], 0);
    fence_after_sync();
    issue_l1(0);
    if (nk > 1) {
      mbar_wait(&full[1], 0);
      fence_after_sync();
      issue_l1(1);
    }
  }

  // ---- the three epilogues of one tile; each ends by publishing its output and issuing the next GEMM ----
  auto phase1 = [&](int s, int64_t k, uint32_t par_bit) {
    const int64_t tile = tile_of(k), grow = tile * TCM_ROWS + rt;
    const bool live = grow < A.N;
    wait_sync(&bl1[s], par_bit);
    if (tid == 0 && k + 2 < nk) load_input(s, k + 2);   // the layer-1 GEMM is done with this slot's input tile
    uint8_t* r1i = A.r1 ? reinterpret_cast<uint8_t*>(A.r1) + tile * (128 * 256 * 2) : nullptr;
    uint8_t* y1i = A.y1 ? reinterpret_cast<uint8_t*>(A.y1) + tile * (128 * 256 * 2) : nullptr;
    float* st1 = A.st1 ? A.st1 + tile * 256 : nullptr;
    float hp;
    const uint32_t slot = tmem + 256 * s;
    if (A.hour) {
      const float* tail_row = live ? A.tail1 + clamp_id(A.hour[grow], 24) * H1 : nullptr;
      mlp_epilogue<256, false, true, true>(slot, q, h, lane, grow, live, par + PAR_B0, par + PAR_G0, par + PAR_E0, tail_row,
                                           A.rng[0], s_stat, nullptr, r1i, y1i, nullptr, nullptr, hp, st1, slot);
    } else {
      mlp_epilogue<256, false, false, true>(slot, q, h, lane, grow, live, par + PAR_B0, par + PAR_G0, par + PAR_E0, nullptr,
                                            A.rng[0], s_stat, nullptr, r1i, y1i, nullptr, nullptr, hp, st1, slot);
    }
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_gemm_ts(slot + 128, slot, sW1, 128, 256 * 16, 256, make_idesc(128, 128, false, false), 16, false);
      mma_commit(&bl2[s]);
    }
  };
  auto phase2 = [&](int s, int64_t k, uint32_t par_bit) {
    const int64_t tile = tile_of(k), grow = tile * TCM_ROWS + rt;
    const bool live = grow < A.N;
    wait_sync(&bl2[s], par_bit);
    uint8_t* r2i = A.r2 ? reinterpret_cast<uint8_t*>(A.r2) + tile * (128 * 128 * 2) : nullptr;
    uint8_t* y2i = A.y2 ? reinterpret_cast<uint8_t*>(A.y2) + tile * (128 * 128 * 2) : nullptr;
    float* st2 = A.st2 ? A.st2 + tile * 256 : nullptr;
    float hp;
    const uint32_t slot = tmem + 256 * s;
    mlp_epilogue<128, false, false, true>(slot + 128, q, h, lane, grow, live, par + PAR_B1, par + PAR_G1, par + PAR_E1, nullptr,
                                          A.rng[1], s_stat, nullptr, r2i, y2i, nullptr, nullptr, hp, st2, slot);
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_gemm_ts(slot + 64, slot, sW2, 128, 128 * 16, 256, make_idesc(128, 64, false, false), 8, false);
      mma_commit(&bl3[s]);
    }
  };
  auto phase3 = [&](int s, int64_t k, uint32_t par_bit) {
    const int64_t tile = tile_of(k), grow = tile * TCM_ROWS + rt;
    const bool live = grow < A.N;
    wait_sync(&bl3[s], par_bit);
    uint8_t* r3i = A.r3 ? reinterpret_cast<uint8_t*>(A.r3) + tile * (128 * 64 * 2) : nullptr;
    float* st3 = A.st3 ? A.st3 + tile * 256 : nullptr;
    float hp;
    const uint32_t slot = tmem + 256 * s;
    mlp_epilogue<64, true, false, true>(slot + 64, q, h, lane, grow, live, par + PAR_B2, par + PAR_G2, par + PAR_E2, nullptr,
                                        A.rng[2], s_stat, nullptr, r3i, nullptr, A.y3, par + PAR_WOUT, hp, st3, 0);
    s_head[rt * MLP_NH + h] = hp;
    fence_before_sync();
    __syncthreads();
    if (tid == 0 && k + 2 < nk) {      // the slot is free: layer-1 GEMM of its next tile (input prefetched in phase 1)
      mbar_wait(&full[s], par_bit ^ 1u);
      fence_after_sync();
      issue_l1(s);
    }
    if (h == 0 && live) {
      const float mp = (s_head[rt * MLP_NH] + s_head[rt * MLP_NH + 1]) + (s_head[rt * MLP_NH + 2] + s_head[rt * MLP_NH + 3]) +
                       par[PAR_SCAL + 0];
      const float z = fmaf(par[PAR_SCAL + 1], A.mf_pred[grow], fmaf(par[PAR_SCAL + 2], mp, par[PAR_SCAL + 3]));
      const float pr = 1.0f / (1.0f + expf(-z));
      A.mlp_pred[grow] = mp;
      A.out[grow] = pr;
      if (A.out2) A.out2[grow] = pr;
    }
    // s_head is next written two CTA barriers later at the earliest (stats barrier + publish barrier of the
    // following epilogue), after every h == 0 thread has passed them: no extra barrier needed here
  };

  uint32_t it = 0;
  for (int64_t k0 = 0; k0 < nk; k0 += 2, ++it) {
    const int ns = k0 + 1 < nk ? 2 : 1;
    const uint32_t pb = it & 1u;
    // one copy of each epilogue in the instruction stream (the slot is a run-time value): the kernel stays
    // inside the instruction cache
#pragma unroll 1
    for (int s = 0; s < ns; ++s) phase1(s, k0 + s, pb);
#pragma unroll 1
    for (int s = 0; s < ns; ++s) phase2(s, k0 + s, pb);
#pragma unroll 1
    for (int s = 0; s < ns; ++s) phase3(s, k0 + s, pb);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int launch_mlp_tc_fwd(const MlpFwdArgs& A, cudaStream_t st) {
  if (A.N == 0) return NCF_OK;
  static int variant = -1;      // NCF_MLP_FWD=1 selects the single-tile kernel (activation tiles in shared memory)
  if (variant < 0) {
    const char* e = getenv("NCF_MLP_FWD");
    variant = e ? atoi(e) : 2;
    NCF_CUDA(cudaFuncSetAttribute(mlp_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_MLP_TOTAL));
    NCF_CUDA(cudaFuncSetAttribute(mlp_tc_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM2_TOTAL));
  }
  const int64_t ntiles = (A.N + TCM_ROWS - 1) / TCM_ROWS;
  const int grid = even_grid(ntiles, tower_sms());
  if (variant == 1) mlp_tc_fwd_kernel<<<grid, MLP_THREADS, SM_MLP_TOTAL, st>>>(A);
  else NCF_CUDA(launch_pdl(PDL_MLP_FWD, mlp_tc_fwd2_kernel, dim3(grid), dim3(MLP_THREADS), SM2_TOTAL, st, A));
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}


// Register-staged operand tiles: the global loads of the NEXT tile are issued before the current tile's MMA
// is awaited, so their latency overlaps the tensor work and the epilogue; the bf16 conversion and the
// shared-memory stores happen when the buffer is free again.
template <int C, int NT>      // C source columns (fp32, row-major, ld = C); NT threads
struct TileRegs {
  static constexpr int CHUNKS = 128 * (C / 8) / NT;
  float4 a[CHUNKS], b[CHUNKS];
  __device__ __forceinline__ static void rc(int c, int& r, int& j) {
    const int blk = c >> 5, l = c & 31;
    constexpr int bpr = (C / 8) / 4;
    r = (blk / bpr) * 8 + (l & 7);
    j = (blk % bpr) * 4 + (l >> 3);
  }
  __device__ __forceinline__ void load(const float* __restrict__ src, int64_t row0, int64_t avail, int tid) {
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k) {
      int r, j;
      rc(tid + k * NT, r, j);
      a[k] = make_float4(0, 0, 0, 0);
      b[k] = a[k];
      if (r < avail) {
        // volatile asm: the compiler must not sink these loads towards their first use (the point of the prefetch)
        const float* p = src + (row0 + r) * C + 8 * j;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(a[k].x), "=f"(a[k].y), "=f"(a[k].z), "=f"(a[k].w)
                     : "l"(p));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(b[k].x), "=f"(b[k].y), "=f"(b[k].z), "=f"(b[k].w)
                     : "l"(p + 4));
      }
    }
  }
  // LAYOUT = column count of the shared-memory tile layout (>= C)
  template <int LAYOUT>
  __device__ __forceinline__ void store(uint8_t* tile, int tid) const {
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k) {
      int r, j;
      rc(tid + k * NT, r, j);
      *reinterpret_cast<uint4*>(tile + tile_off(r, 8 * j, LAYOUT)) =
          make_uint4(pack_bf16(a[k].x, a[k].y), pack_bf16(a[k].z, a[k].w), pack_bf16(b[k].x, b[k].y), pack_bf16(b[k].z, b[k].w));
    }
  }
};

// =============================================================================================
// Generic tcgen05 linear layers for the attention projections (q, k|v, out):
//   tc_linear_kernel<K, NOUT, false>:  Y[N,NOUT] = X[N,K] . W[NOUT,K]^T + b      (nn.Linear forward)
//   tc_linear_kernel<K, NOUT, true> :  Y[N,NOUT] = X[N,K] . W[K,NOUT]            (input gradient dX = dY . W)
//   tc_wgrad64_kernel<CZ>           :  dW[CZ,64] += Z[N,CZ]^T . X[N,64],  db[CZ] += colsum(Z)
// fp32 tensors in global memory, bf16 operands in shared memory, fp32 accumulation in TMEM.  Several
// CTAs per SM hide the load -> MMA -> store latency of the short per-tile pipeline.
// =============================================================================================
template <int K, int NOUT, bool DGRAD, bool OUT_IMG = false>
__global__ void __launch_bounds__(TCM_THREADS, 2) tc_linear_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                                   const float* __restrict__ bias, float* __restrict__ Y,
                                                                   int64_t N) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  constexpr int WR = DGRAD ? K : NOUT, WC = DGRAD ? NOUT : K;     // stored weight image [WR][WC]
  constexpr uint32_t OFF_X = WR * WC * 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, h = warp >> 2;
  fill_tile_f32<WC>(smem, W, WC, 0, WR, WR, tid, TCM_THREADS);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, NOUT < 32 ? 32 : NOUT);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t sW = smem_addr(smem), sX = smem_addr(smem + OFF_X);
  uint32_t phase = 0;
  const int64_t ntiles = (N + TCM_ROWS - 1) / TCM_ROWS;
  TileRegs<K, TCM_THREADS> xr;
  if ((int64_t)blockIdx.x < ntiles)
    xr.load(X, (int64_t)blockIdx.x * TCM_ROWS, min((int64_t)TCM_ROWS, N - (int64_t)blockIdx.x * TCM_ROWS), tid);
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * TCM_ROWS;
    const int64_t avail = min((int64_t)TCM_ROWS, N - row0);
    xr.template store<K>(smem + OFF_X, tid);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      if (DGRAD)
        issue_gemm(tmem, sX, 128, K * 16, 256, sW, WC * 16, 128, 2 * WC * 16, make_idesc(128, NOUT, false, true), K / 16, false);
      else
        issue_gemm(tmem, sX, 128, K * 16, 256, sW, 128, WC * 16, 256, make_idesc(128, NOUT, false, false), K / 16, false);
      mma_commit(&bar);
    }
    {
      const int64_t next = tile + gridDim.x;
      if (next < ntiles) xr.load(X, next * TCM_ROWS, min((int64_t)TCM_ROWS, N - next * TCM_ROWS), tid);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();
    const int rt = q * 32 + lane;
    const int64_t grow = row0 + rt;
    constexpr int HALF = NOUT / 2;
#pragma unroll 1
    for (int ch = 0; ch < HALF / 32; ++ch) {
      float v[32];
      const int c0 = h * HALF + ch * 32;
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
      if (OUT_IMG) {     // bf16 tile image (rows beyond N hold the bias: finite padding)
        uint8_t* img = reinterpret_cast<uint8_t*>(Y) + tile * (128 * NOUT * 2);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = v[8 * j + k] + (bias ? __ldg(bias + c0 + 8 * j + k) : 0.f);
          *reinterpret_cast<uint4*>(img + tile_off(rt, c0 + 8 * j, NOUT)) =
              make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
        }
      } else if (rt < avail) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          if (bias) {
            const float4 b = ldg4(bias + c0 + 4 * j);
            o = f4_add(o, b);
          }
          st4(Y + grow * NOUT + c0 + 4 * j, o);
        }
      }
    }
    fence_before_sync();
    __syncthreads();     // TMEM and the X tile are free again
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, NOUT < 32 ? 32 : NOUT);
}

template <int K, int NOUT, bool DGRAD, bool OUT_IMG = false>
static int launch_tc_linear(const float* X, const float* W, const float* bias, float* Y, int64_t N, cudaStream_t st) {
  if (N == 0) return NCF_OK;
  constexpr int smem = K * NOUT * 2 + 128 * K * 2;
  static bool configured = false;
  if (!configured) {
    NCF_CUDA(cudaFuncSetAttribute(tc_linear_kernel<K, NOUT, DGRAD, OUT_IMG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int64_t ntiles = (N + TCM_ROWS - 1) / TCM_ROWS;
  const int grid = (int)std::min<int64_t>(ntiles, (int64_t)num_sms() * 2);
  tc_linear_kernel<K, NOUT, DGRAD, OUT_IMG><<<grid, TCM_THREADS, smem, st>>>(X, W, bias, Y, N);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

// Z tile is stored with the 128-column layout (upper half stays zero when CZ == 64) so that M = 128;
// the X tile carries 16 extra columns whose first one is 1.0: accumulator column 64 = bias gradient.
template <int CZ>
__global__ void __launch_bounds__(TCM_THREADS, 2) tc_wgrad64_kernel(const float* __restrict__ Z, const float* __restrict__ X,
                                                                    float* __restrict__ dW, float* __restrict__ db,
                                                                    int64_t N) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  constexpr int CXL = 80;                               // X tile layout width: 64 data + 16 (ones | zeros)
  constexpr uint32_t OFF_X = 128 * 128 * 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, h = warp >> 2;
  for (int i = tid; i < (128 * 128 * 2 + 128 * CXL * 2) / 16; i += TCM_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t sZ = smem_addr(smem), sX = smem_addr(smem + OFF_X);
  uint32_t phase = 0;
  bool first = true;
  const int64_t ntiles = (N + TCM_ROWS - 1) / TCM_ROWS;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * TCM_ROWS;
    const int64_t avail = min((int64_t)TCM_ROWS, N - row0);
    if (!first) {
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    // Z: CZ columns into the 128-wide layout; X: 64 columns into the 80-wide layout + the ones column
    for (int c = tid; c < 128 * (CZ / 8); c += TCM_THREADS) {
      const int blk = c >> 5, l = c & 31;
      const int bpr = (CZ / 8) / 4;
      const int r = (blk / bpr) * 8 + (l & 7), j = (blk % bpr) * 4 + (l >> 3);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (r < avail) {
        const float* p = Z + (row0 + r) * CZ + 8 * j;
        const float4 a = ldg4(p), b = ldg4(p + 4);
        v = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
      }
      *reinterpret_cast<uint4*>(smem + tile_off(r, 8 * j, 128)) = v;
    }
    for (int c = tid; c < 128 * 9; c += TCM_THREADS) {
      int r, j;
      if (c < 128 * 8) {
        const int blk = c >> 5, l = c & 31;
        r = (blk / 2) * 8 + (l & 7);
        j = (blk % 2) * 4 + (l >> 3);
      } else {
        r = c - 128 * 8;
        j = 8;
      }
      uint4 v = make_uint4(0, 0, 0, 0);
      if (r < avail) {
        if (j < 8) {
          const float* p = X + (row0 + r) * 64 + 8 * j;
          const float4 a = ldg4(p), b = ldg4(p + 4);
          v = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
        } else {
          v.x = 0x00003f80u;     // bf16(1.0) in the low half: column 64 = 1, columns 65..71 = 0
        }
      }
      *reinterpret_cast<uint4*>(smem + OFF_X + tile_off(r, 8 * j, CXL)) = v;
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_gemm(tmem, sZ, 128 * 16, 128, 2 * 128 * 16, sX, CXL * 16, 128, 2 * CXL * 16, make_idesc(128, CXL, true, true), 8, !first);
      mma_commit(&bar);
    }
    first = false;
  }
  if (!first) {
    mbar_wait(&bar, phase);
    fence_after_sync();
    const int m = q * 32 + lane;
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + h * 32, v);
    if (m < CZ) {
#pragma unroll
      for (int i = 0; i < 32; ++i) atomicAdd(dW + (int64_t)m * 64 + h * 32 + i, v[i]);
    }
    if (h == 0 && db) {
      float b[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 64, b);
      if (m < CZ) atomicAdd(db + m, b[0]);
    } else {
      float b[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 64, b);   // keep the warp-collective load uniform
      (void)b;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

template <int CZ>
static int launch_tc_wgrad64(const float* Z, const float* X, float* dW, float* db, int64_t N, cudaStream_t st) {
  if (N == 0) return NCF_OK;
  constexpr int smem = 128 * 128 * 2 + 128 * 80 * 2;
  static bool configured = false;
  if (!configured) {
    NCF_CUDA(cudaFuncSetAttribute(tc_wgrad64_kernel<CZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int64_t ntiles = (N + TCM_ROWS - 1) / TCM_ROWS;
  const int grid = (int)std::min<int64_t>(ntiles, (int64_t)num_sms() * 2);
  tc_wgrad64_kernel<CZ><<<grid, TCM_THREADS, smem, st>>>(Z, X, dW, db, N);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

// Backward of one projection Y = X.W^T + b in ONE pass over the rows: the dY tile is loaded once and used
// as K-major A operand of the input gradient (dX = dY.W, fresh TMEM columns per tile) and as MN-major A
// operand of the weight gradient (dW += dY^T.X, bias gradient through the ones column; TMEM columns that
// accumulate over all tiles of the CTA).
template <int CZ>
__global__ void __launch_bounds__(TCM_THREADS, 2) tc_proj_bwd_kernel(const float* __restrict__ Z, const float* __restrict__ X,
                                                                     const float* __restrict__ W, float* __restrict__ dX,
                                                                     float* __restrict__ dW, float* __restrict__ db,
                                                                     int64_t N) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  constexpr int CXL = 80;
  constexpr uint32_t OFF_X = 128 * 128 * 2, OFF_W = OFF_X + 128 * CXL * 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, h = warp >> 2;
  for (int i = tid; i < (int)OFF_W / 16; i += TCM_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fill_tile_f32<64>(smem + OFF_W, W, 64, 0, CZ, CZ, tid, TCM_THREADS);          // W image [CZ rows][64 cols]
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
  const uint32_t sZ = smem_addr(smem), sX = smem_addr(smem + OFF_X), sW = smem_addr(smem + OFF_W);
  uint32_t phase = 0;
  bool first = true;
  const int64_t ntiles = (N + TCM_ROWS - 1) / TCM_ROWS;
  TileRegs<CZ, TCM_THREADS> zr;
  TileRegs<64, TCM_THREADS> xr;
  if ((int64_t)blockIdx.x < ntiles) {
    const int64_t r0 = (int64_t)blockIdx.x * TCM_ROWS;
    zr.load(Z, r0, min((int64_t)TCM_ROWS, N - r0), tid);
    xr.load(X, r0, min((int64_t)TCM_ROWS, N - r0), tid);
  }
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * TCM_ROWS;
    const int64_t avail = min((int64_t)TCM_ROWS, N - row0);
    zr.template store<128>(smem, tid);
    xr.template store<CXL>(smem + OFF_X, tid);
    if (tid < 128) {      // the ones column (bias gradient): 1.0 for live rows
      uint4 v = make_uint4(0, 0, 0, 0);
      if (tid < avail) v.x = 0x00003f80u;
      *reinterpret_cast<uint4*>(smem + OFF_X + tile_off(tid, 64, CXL)) = v;
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      // dX[128 x 64] = dY[128 x CZ] . W[CZ x 64]      (A K-major on the 128-wide layout, B MN-major)
      issue_gemm(tmem + 0, sZ, 128, 128 * 16, 256, sW, 64 * 16, 128, 2 * 64 * 16, make_idesc(128, 64, false, true), CZ / 16, false);
      // dW[CZ x 64 (+bias col)] += dY^T . [X | 1]
      issue_gemm(tmem + 64, sZ, 128 * 16, 128, 2 * 128 * 16, sX, CXL * 16, 128, 2 * CXL * 16, make_idesc(128, CXL, true, true), 8, !first);
      mma_commit(&bar);
    }
    first = false;
    {
      const int64_t next = tile + gridDim.x;
      if (next < ntiles) {
        zr.load(Z, next * TCM_ROWS, min((int64_t)TCM_ROWS, N - next * TCM_ROWS), tid);
        xr.load(X, next * TCM_ROWS, min((int64_t)TCM_ROWS, N - next * TCM_ROWS), tid);
      }
    }
    if (warp == 0) mbar_wait(&bar, phase);
    phase ^= 1;
    __syncthreads();
    fence_after_sync();
    {
      const int rt = q * 32 + lane;
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + h * 32, v);
      if (rt < avail) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st4(dX + (row0 + rt) * 64 + h * 32 + 4 * j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
      }
    }
    fence_before_sync();
    __syncthreads();
  }
  if (!first) {
    const int m = q * 32 + lane;
    float v[32];
    tmem_ld32(tmem + 64 + ((uint32_t)(q * 32) << 16) + h * 32, v);
    if (m < CZ) {
#pragma unroll
      for (int i = 0; i < 32; ++i) atomicAdd(dW + (int64_t)m * 64 + h * 32 + i, v[i]);
    }
    float b[32];
    tmem_ld32(tmem + 64 + ((uint32_t)(q * 32) << 16) + 64, b);
    if (h == 0 && db && m < CZ) atomicAdd(db + m, b[0]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

template <int CZ>
static int launch_tc_proj_bwd(const float* Z, const float* X, const float* W, float* dX, float* dW, float* db, int64_t N,
                              cudaStream_t st) {
  if (N == 0) return NCF_OK;
  constexpr int smem = 128 * 128 * 2 + 128 * 80 * 2 + CZ * 64 * 2;
  static bool configured = false;
  if (!configured) {
    NCF_CUDA(cudaFuncSetAttribute(tc_proj_bwd_kernel<CZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int64_t ntiles = (N + TCM_ROWS - 1) / TCM_ROWS;
  const int grid = (int)std::min<int64_t>(ntiles, (int64_t)num_sms() * 2);
  tc_proj_bwd_kernel<CZ><<<grid, TCM_THREADS, smem, st>>>(Z, X, W, dX, dW, db, N);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}
int tc_proj_backward(int which, const float* dY, const float* X, const float* W, float* dX, float* dW, float* db, int64_t N,
                     cudaStream_t st) {
  return which == 0 ? launch_tc_proj_bwd<64>(dY, X, W, dX, dW, db, N, st) : launch_tc_proj_bwd<128>(dY, X, W, dX, dW, db, N, st);
}

// entry points used by the tower orchestration (ncf_tower_f32.cu) when precision == NCF_BF16_TC
int tc_proj_forward(int which, const float* X, const float* W, const float* bias, float* Y, int64_t N, cudaStream_t st) {
  // which: 0 = 64 -> 64 (q, v, out), 1 = 64 -> 128 (k|v)
  return which == 0 ? launch_tc_linear<64, 64, false>(X, W, bias, Y, N, st) : launch_tc_linear<64, 128, false>(X, W, bias, Y, N, st);
}
int tc_proj_forward_img(const float* X, const float* W, const float* bias, void* img, int64_t N, cudaStream_t st) {
  return launch_tc_linear<64, 64, false, true>(X, W, bias, reinterpret_cast<float*>(img), N, st);
}
int tc_proj_dgrad(int which, const float* dY, const float* W, float* dX, int64_t N, cudaStream_t st) {
  // which: 0 = dY[N,64] . W[64,64], 1 = dY[N,128] . W[128,64]
  return which == 0 ? launch_tc_linear<64, 64, true>(dY, W, nullptr, dX, N, st) : launch_tc_linear<128, 64, true>(dY, W, nullptr, dX, N, st);
}
int tc_proj_wgrad(int which, const float* Z, const float* X, float* dW, float* db, int64_t N, cudaStream_t st) {
  return which == 0 ? launch_tc_wgrad64<64>(Z, X, dW, db, N, st) : launch_tc_wgrad64<128>(Z, X, dW, db, N, st);
}

int mlp_tc_forward(const ncf_run_cfg& cfg, const float* dense, int64_t N, const int64_t* hour, const float* tail1,
                   float* out, TowerWs& w, cudaStream_t st) {
  MlpFwdArgs A{};
  const bool train = cfg.training != 0;
  A.a_img = (const __nv_bfloat16*)w.a_img;
  A.st1 = train ? w.st1 : nullptr;
  A.st2 = train ? w.st2 : nullptr;
  A.st3 = train ? w.st3 : nullptr;
  A.dense = dense;
  A.hour = hour;
  A.tail1 = tail1;
  A.mf_pred = w.mf_pred;
  A.out = out;
  A.out2 = w.p_saved;
  A.mlp_pred = w.mlp_pred;
  A.y3 = nullptr;      // the head runs inside the kernel and the backward rebuilds y3's column sums from r3 (mlp_bwd_layer)
  A.r1 = train ? (__nv_bfloat16*)w.r1b : nullptr;
  A.y1 = train ? (__nv_bfloat16*)w.y1b : nullptr;
  A.r2 = train ? (__nv_bfloat16*)w.r2b : nullptr;
  A.y2 = train ? (__nv_bfloat16*)w.y2b : nullptr;
  A.r3 = train ? (__nv_bfloat16*)w.r3b : nullptr;
  A.N = N;
  for (int l = 0; l < 3; ++l) A.rng[l] = make_rng(cfg, 1 + l);
  return launch_mlp_tc_fwd(A, st);
}

// =============================================================================================
// MLP tower backward, part 1: the input-gradient chain, fused per 128-row tile.
//   dy3 [N,64] fp32 -> (LN/ReLU/dropout backward) dz3 -> dy2 = dz3.W2 -> dz2 -> dy1 = dz2.W1 -> dz1
//   -> da = dz1.W0[:, :64]
// The forward's weight images serve as MN-major B operands (no transposed copies).  dz tiles are
// written to global memory (bf16) for the weight-gradient kernel; the LayerNorm-affine and bias
// gradients are column sums reduced with a shuffle transpose and shared-memory atomics.
// =============================================================================================
constexpr uint32_t SMB_STAT = SM_PAR + PAR_COUNT * 4;                 // 2 x [128][4][2] floats
constexpr uint32_t SMB_ACC = SMB_STAT + 2 * 128 * MLP_NH * 2 * 4;     // column-sum accumulators
constexpr int ACC_L0 = 0, ACC_L1 = 768, ACC_L2 = 1152, ACC_COUNT = 1344;   // per layer: [dgamma | dbeta | dbias]
constexpr uint32_t SMB_TOTAL = SMB_ACC + ACC_COUNT * 4;

struct MlpBwdArgs {
  const float* dense;
  float* dense_grad;
  float* da;                                   // out: da [N,64] fp32, or bf16 rows when da_bf16 (fused attention backward)
  bool da_bf16;
  const float* d_mlp_pred;                     // [N] dL/d mlp_pred from head_bwd_kernel
  const __nv_bfloat16 *r1, *r2, *r3;
  const float *st1, *st2, *st3;                // LayerNorm (mean, rstd) per row from the forward
  __nv_bfloat16 *dz1, *dz2, *dz3;
  int64_t N;
  DropoutRng rng[3];
};

// column sums of a [32 rows (lanes) x W cols (registers)] block: lane l (< W) ends up with column l
template <int W>
__device__ __forceinline__ float warp_transpose_sum(float (&v)[W], int lane) {
  if constexpr (W == 16) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
  }
#pragma unroll
  for (int off = (W == 32 ? 16 : 8); off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// W consecutive bf16 of one row of a tile image: 16-byte chunks sit 128 bytes apart
template <int W>
__device__ __forceinline__ void load_imgw(const uint8_t* __restrict__ p, float (&v)[W]) {
#pragma unroll
  for (int j = 0; j < W / 8; ++j) {
    const uint4 q4 = __ldg(reinterpret_cast<const uint4*>(p + j * 128));
    float t[8];
    unpack_bf16x8(q4, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[8 * j + i] = t[i];
  }
}

// One layer of the backward chain for this thread's row and column part.
// The saved r tile carries relu(z) as non-negative bf16 with the dropout decision in the sign bit (set =
// dropped), so no random numbers are drawn here.  `gam` = gamma * 1/(1-p) (dropout scale folded in); the
// column sums for d gamma / d beta are accumulated WITHOUT that scale and multiplied once at the flush.
// Pass B: row sums of dy*gamma and dy*gamma*xhat, column sums for d gamma / d beta.
// Pass C: dz = relu'(r) * LayerNorm backward -> bf16 tile (next A operand) + global + bias column sums.
// chunk `ch` (CW columns) of this thread's part of a saved r tile: CW/8 16-byte pieces, 128 bytes apart
template <int C>
struct BwdGeom {
  static constexpr int PART = C / MLP_NH, CW = PART >= 32 ? 32 : 16, NCH = PART / CW;
};
template <int C>
__device__ __forceinline__ void load_r_chunk(const uint8_t* __restrict__ r_img, int rt, int h, int ch,
                                             uint32_t (&rw)[BwdGeom<C>::CW / 2]) {
  constexpr int PART = BwdGeom<C>::PART, CW = BwdGeom<C>::CW;
  const uint8_t* p = r_img + tile_off(rt, h * PART + ch * CW, C);
#pragma unroll
  for (int j = 0; j < CW / 8; ++j) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p + j * 128));
    rw[4 * j] = t.x; rw[4 * j + 1] = t.y; rw[4 * j + 2] = t.z; rw[4 * j + 3] = t.w;
  }
}

template <int C, bool FROM_TMEM>
__device__ __forceinline__ void mlp_bwd_layer(uint32_t tmem_dy, float dml, const float* __restrict__ wout, int q, int h, int lane,
                                              int64_t grow, bool live, const uint8_t* __restrict__ r_img,
                                              uint32_t (&rw)[BwdGeom<C>::CW / 2], const float* __restrict__ gam,
                                              float* s_statB, float* s_acc, uint8_t* dztile, uint8_t* __restrict__ dz_img,
                                              const float* __restrict__ st_tile) {
  constexpr int PART = BwdGeom<C>::PART, CW = BwdGeom<C>::CW, NCH = BwdGeom<C>::NCH;
  const int rt = q * 32 + lane;
  const uint32_t taddr = tmem_dy + ((uint32_t)(q * 32) << 16) + h * PART;
  const bool acc_lane = lane < CW;
  // LayerNorm statistics of the saved relu output come from the forward
  const float2 ms = *reinterpret_cast<const float2*>(st_tile + rt * 2);
  const float rstd = ms.y, nmr = -ms.x * ms.y;
  auto load_dy = [&](int ch, float (&dy)[CW]) {
    if (FROM_TMEM) {
      tmem_ldw<CW>(taddr + ch * CW, dy);
    } else {
      // layer 3: dL/dy3 = dL/d mlp_pred * mlp_output.weight (backward of the output head, architecture.py:345).
      // The weight factor lives in `gam` (gamma * dropout scale * w_out), so the upstream value of every column of
      // the row is dL/d mlp_pred itself and the column sums of pass B are those of keep * dml * xhat and keep * dml:
      // d gamma, d beta AND d mlp_output.weight all follow from them at the flush.  dml is 0 for rows beyond N
#pragma unroll
      for (int i = 0; i < CW; ++i) dy[i] = dml;
    }
  };

  // `rw` arrives holding chunk 0 (loaded by the caller BEFORE it waited for the upstream MMA); with more than
  // one chunk the next one is requested before the current one is consumed, wrapping around to chunk 0 for pass C.
  float s1 = 0.f, s2 = 0.f;
  float s1p[4] = {0.f, 0.f, 0.f, 0.f}, s2p[4] = {0.f, 0.f, 0.f, 0.f};      // independent chains
#pragma unroll 1
  for (int ch = 0; ch < NCH; ++ch) {
    const int c0 = h * PART + ch * CW;
    float dy[CW], t[CW];
    uint32_t rn[CW / 2];
    if (NCH > 1) load_r_chunk<C>(r_img, rt, h, ch + 1 < NCH ? ch + 1 : 0, rn);
    load_dy(ch, dy);
#pragma unroll
    for (int i2 = 0; i2 < CW / 2; ++i2) {
      const uint32_t w = rw[i2];
      const uint32_t lo = w << 16;
      const float2 g2 = *reinterpret_cast<const float2*>(gam + c0 + 2 * i2);
      const float d0 = (int32_t)lo < 0 ? 0.f : dy[2 * i2], d1 = (int32_t)w < 0 ? 0.f : dy[2 * i2 + 1];
      const float xh0 = fmaf(fabsf(__uint_as_float(lo)), rstd, nmr);
      const float xh1 = fmaf(fabsf(__uint_as_float(w & 0xffff0000u)), rstd, nmr);
      const float dg0 = d0 * g2.x, dg1 = d1 * g2.y;
      s1p[i2 & 3] += dg0 + dg1;
      s2p[i2 & 3] = fmaf(dg0, xh0, fmaf(dg1, xh1, s2p[i2 & 3]));
      t[2 * i2] = d0 * xh0;
      t[2 * i2 + 1] = d1 * xh1;
      dy[2 * i2] = d0;
      dy[2 * i2 + 1] = d1;
    }
    if (NCH > 1) {
#pragma unroll
      for (int i = 0; i < CW / 2; ++i) rw[i] = rn[i];
    }
    const float cg = warp_transpose_sum<CW>(t, lane);
    const float cb = warp_transpose_sum<CW>(dy, lane);
    if (acc_lane) {
      atomicAdd(s_acc + c0 + lane, cg);
      atomicAdd(s_acc + C + c0 + lane, cb);
    }
  }
  s1 = (s1p[0] + s1p[1]) + (s1p[2] + s1p[3]);
  s2 = (s2p[0] + s2p[1]) + (s2p[2] + s2p[3]);
  s_statB[(rt * MLP_NH + h) * 2 + 0] = s1;
  s_statB[(rt * MLP_NH + h) * 2 + 1] = s2;
  quarter_sync(q);          // only the four warps that share these 32 rows exchange their sums
  s1 = 0.f;
  s2 = 0.f;
#pragma unroll
  for (int k = 0; k < MLP_NH; ++k) {
    s1 += s_statB[(rt * MLP_NH + k) * 2 + 0];
    s2 += s_statB[(rt * MLP_NH + k) * 2 + 1];
  }
  const float nm2 = -s2 * (1.0f / C), nm1r = -s1 * (1.0f / C) * rstd;

#pragma unroll 1
  for (int ch = 0; ch < NCH; ++ch) {
    const int c0 = h * PART + ch * CW;
    float dy[CW];
    uint32_t rn[CW / 2];
    if (NCH > 1 && ch + 1 < NCH) load_r_chunk<C>(r_img, rt, h, ch + 1, rn);
    load_dy(ch, dy);
#pragma unroll
    for (int i2 = 0; i2 < CW / 2; ++i2) {
      const uint32_t w = rw[i2];
      const uint32_t lo = w << 16;
      const float2 g2 = *reinterpret_cast<const float2*>(gam + c0 + 2 * i2);
      const float r0 = fabsf(__uint_as_float(lo)), r1 = fabsf(__uint_as_float(w & 0xffff0000u));
      const float dg0 = ((int32_t)lo < 0 ? 0.f : dy[2 * i2]) * g2.x, dg1 = ((int32_t)w < 0 ? 0.f : dy[2 * i2 + 1]) * g2.y;
      // dr = rstd * (dyg - mean(dyg) - xhat * mean(dyg * xhat))
      const float dr0 = fmaf(fmaf(fmaf(r0, rstd, nmr), nm2, dg0), rstd, nm1r);
      const float dr1 = fmaf(fmaf(fmaf(r1, rstd, nmr), nm2, dg1), rstd, nm1r);
      dy[2 * i2] = r0 > 0.f ? dr0 : 0.f;
      dy[2 * i2 + 1] = r1 > 0.f ? dr1 : 0.f;
    }
    if (NCH > 1 && ch + 1 < NCH) {
#pragma unroll
      for (int i = 0; i < CW / 2; ++i) rw[i] = rn[i];
    }
#pragma unroll
    for (int j = 0; j < CW / 8; ++j) {
      const uint4 pk = make_uint4(pack_bf16(dy[8 * j], dy[8 * j + 1]), pack_bf16(dy[8 * j + 2], dy[8 * j + 3]),
                                  pack_bf16(dy[8 * j + 4], dy[8 * j + 5]), pack_bf16(dy[8 * j + 6], dy[8 * j + 7]));
      const uint32_t off = tile_off(rt, c0 + 8 * j, C);
      *reinterpret_cast<uint4*>(dztile + off) = pk;
      *reinterpret_cast<uint4*>(dz_img + off) = pk;
    }
    if constexpr (C == 64) {      // the 256- and 128-wide layers get their bias gradient from the wgrad kernel's ones-GEMM
      const float cz = warp_transpose_sum<CW>(dy, lane);
      if (acc_lane) atomicAdd(s_acc + 2 * C + c0 + lane, cz);
    }
  }
}

__global__ void __launch_bounds__(MLP_THREADS, 1) mlp_tc_bwd_kernel(MlpBwdArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, h = warp >> 2;
  float* par = reinterpret_cast<float*>(smem + SM_PAR);
  float* s_statA = reinterpret_cast<float*>(smem + SMB_STAT);
  float* s_statB = s_statA + 128 * MLP_NH * 2;
  float* s_acc = reinterpret_cast<float*>(smem + SMB_ACC);
  const float* P = A.dense;

  load_weight_image<256, 64>(smem + SM_W0, P + NCF_OFF(NCF_P_MLP0_W), K0, tid, MLP_THREADS);
  load_weight_image<128, 256>(smem + SM_W1, P + NCF_OFF(NCF_P_MLP1_W), H1, tid, MLP_THREADS);
  load_weight_image<64, 128>(smem + SM_W2, P + NCF_OFF(NCF_P_MLP2_W), H2, tid, MLP_THREADS);
  // gamma carries the dropout scale 1/(1-p) (see mlp_bwd_layer)
  const float sc0 = A.rng[0].thresh ? A.rng[0].scale : 1.f, sc1 = A.rng[1].thresh ? A.rng[1].scale : 1.f,
              sc2 = A.rng[2].thresh ? A.rng[2].scale : 1.f;
  for (int i = tid; i < 256; i += MLP_THREADS) {
    par[PAR_G0 + i] = P[NCF_OFF(NCF_P_LN0_W) + i] * sc0;
    if (i < 128) par[PAR_G1 + i] = P[NCF_OFF(NCF_P_LN1_W) + i] * sc1;
    if (i < 64) {
      par[PAR_WOUT + i] = P[NCF_OFF(NCF_P_MLP_OUT_W) + i];
      par[PAR_G2 + i] = P[NCF_OFF(NCF_P_LN2_W) + i] * sc2 * P[NCF_OFF(NCF_P_MLP_OUT_W) + i];     // see mlp_bwd_layer<64, false>
    }
  }
  for (int i = tid; i < ACC_COUNT; i += MLP_THREADS) s_acc[i] = 0.f;
  if (tid == 0) {
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
  const uint32_t sZ = smem_addr(smem + SM_Y);
  uint8_t* ztile = smem + SM_Y;
  uint32_t phase = 0;

  const int64_t ntiles = (A.N + TCM_ROWS - 1) / TCM_ROWS;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * TCM_ROWS;
    const int64_t avail = min((int64_t)TCM_ROWS, A.N - row0);
    const int rt = q * 32 + lane;
    const int64_t grow = row0 + rt;
    const bool live = rt < avail;
    // layer 3 (64 wide): dy3 from global
    const uint8_t* r1i = reinterpret_cast<const uint8_t*>(A.r1) + tile * (128 * 256 * 2);
    const uint8_t* r2i = reinterpret_cast<const uint8_t*>(A.r2) + tile * (128 * 128 * 2);
    const uint8_t* r3i = reinterpret_cast<const uint8_t*>(A.r3) + tile * (128 * 64 * 2);
    uint8_t* z1i = reinterpret_cast<uint8_t*>(A.dz1) + tile * (128 * 256 * 2);
    uint8_t* z2i = reinterpret_cast<uint8_t*>(A.dz2) + tile * (128 * 128 * 2);
    uint8_t* z3i = reinterpret_cast<uint8_t*>(A.dz3) + tile * (128 * 64 * 2);
    uint32_t rw3[BwdGeom<64>::CW / 2], rw2[BwdGeom<128>::CW / 2], rw1[BwdGeom<256>::CW / 2];
    if (tid == 0) {       // pull the NEXT tile's saved tensors into L2 while this one is processed
      const int64_t nt = tile + gridDim.x;
      if (nt < ntiles) {
        bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(A.r1) + nt * (128 * 256 * 2), 128 * 256 * 2);
        bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(A.r2) + nt * (128 * 128 * 2), 128 * 128 * 2);
        bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(A.r3) + nt * (128 * 64 * 2), 128 * 64 * 2);
      }
    }
    load_r_chunk<64>(r3i, rt, h, 0, rw3);
    load_r_chunk<128>(r2i, rt, h, 0, rw2);       // consumed after the first MMA: its latency hides behind layer 3
    mlp_bwd_layer<64, false>(0, live ? A.d_mlp_pred[grow] : 0.f, par + PAR_WOUT, q, h, lane, grow, live, r3i, rw3, par + PAR_G2, s_statB, s_acc + ACC_L2, ztile,
                             z3i, A.st3 + tile * 256);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();   // dy2[128x128] = dz3[128x64] . W2[64x128]
      issue_gemm(tmem + 0, sZ, 128, 64 * 16, 256, sW2, 128 * 16, 128, 2 * 128 * 16, make_idesc(128, 128, false, true), 4, false);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);                     // every warp polls for itself
    phase ^= 1;
    fence_after_sync();
    load_r_chunk<256>(r1i, rt, h, 0, rw1);       // consumed after the second MMA
    mlp_bwd_layer<128, true>(tmem + 0, 0.f, nullptr, q, h, lane, grow, live, r2i, rw2, par + PAR_G1, s_statB, s_acc + ACC_L1,
                             ztile, z2i, A.st2 + tile * 256);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();   // dy1[128x256] = dz2[128x128] . W1[128x256]
      issue_gemm(tmem + 128, sZ, 128, 128 * 16, 256, sW1, 256 * 16, 128, 2 * 256 * 16, make_idesc(128, 256, false, true), 8, false);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);                     // every warp polls for itself
    phase ^= 1;
    fence_after_sync();
    mlp_bwd_layer<256, true>(tmem + 128, 0.f, nullptr, q, h, lane, grow, live, r1i, rw1, par + PAR_G0, s_statB, s_acc + ACC_L0,
                             ztile, z1i, A.st1 + tile * 256);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();   // da[128x64] = dz1[128x256] . W0[256x64]
      issue_gemm(tmem + 384, sZ, 128, 256 * 16, 256, sW0, 64 * 16, 128, 2 * 64 * 16, make_idesc(128, 64, false, true), 16, false);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);                     // every warp polls for itself
    phase ^= 1;
    fence_after_sync();
    {
      float v[16];
      tmem_ld16(tmem + 384 + ((uint32_t)(q * 32) << 16) + h * 16, v);
      if (live) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_row4(A.da, grow, h * 16 + 4 * j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]), A.da_bf16);
      }
    }
    fence_before_sync();
    __syncthreads();
  }
  __syncthreads();
  float* dg = A.dense_grad;
  for (int i = tid; i < ACC_COUNT; i += MLP_THREADS) {
    int64_t off;
    int k = i;
    float sc;          // d gamma / d beta sums were taken without the dropout scale; the bias sums carry it already
    if (k < ACC_L1) {
      off = k < 256 ? NCF_OFF(NCF_P_LN0_W) + k : k < 512 ? NCF_OFF(NCF_P_LN0_B) + (k - 256) : NCF_OFF(NCF_P_MLP0_B) + (k - 512);
      sc = k < 512 ? sc0 : 1.f;
    } else if (k < ACC_L2) {
      k -= ACC_L1;
      off = k < 128 ? NCF_OFF(NCF_P_LN1_W) + k : k < 256 ? NCF_OFF(NCF_P_LN1_B) + (k - 128) : NCF_OFF(NCF_P_MLP1_B) + (k - 256);
      sc = k < 256 ? sc1 : 1.f;
    } else {
      k -= ACC_L2;
      off = k < 64 ? NCF_OFF(NCF_P_LN2_W) + k : k < 128 ? NCF_OFF(NCF_P_LN2_B) + (k - 64) : NCF_OFF(NCF_P_MLP2_B) + (k - 128);
      // layer 3 summed keep * dml * xhat and keep * dml (no w_out): d gamma / d beta get the column's w_out here and
      // d mlp_output.weight[c] = sum_r dml * y3[r,c] = scale * (gamma_c * sum keep dml xhat + beta_c * sum keep dml)
      sc = k < 128 ? sc2 * par[PAR_WOUT + (k & 63)] : 1.f;
      if (k < 64)
        atomicAdd(dg + NCF_OFF(NCF_P_MLP_OUT_W) + k,
                  sc2 * fmaf(P[NCF_OFF(NCF_P_LN2_W) + k], s_acc[i], P[NCF_OFF(NCF_P_LN2_B) + k] * s_acc[i + 64]));
    }
    atomicAdd(dg + off, s_acc[i] * sc);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// =============================================================================================
// MLP tower backward, part 1, second version (default; NCF_MLP_BWD=1 selects the kernel above).  Same GEMM chain,
// leaner CUDA-core epilogue - the first version spends 38 instructions per element, almost half of them moving bits:
//   * pass B leaves dg = keep * dy * gamma (in place, over dy) and xhat in TENSOR MEMORY, so pass C is two FMAs, a
//     compare-select and the bf16 pack per element: the saved r tile is read once, not twice, and neither unpacked nor
//     sign-tested again.  (relu' is taken from xhat > -mean * rstd, which is r > 0 up to fp32 rounding of r * rstd.)
//   * the column sums for d gamma / d beta (and the 64-wide layer's bias) go through a per-warp 2 KB shared-memory
//     transposition scratch instead of a 5-stage shuffle butterfly (2 FSEL + SHFL + FADD per element and stage): the
//     lane = row values are stored with 128-bit stores (XOR-swizzled 16-byte chunks, conflict-free), lane = (column,
//     row half) reads 16 rows back, and the partial sums stay in REGISTERS across all tiles of the CTA - no shuffles,
//     no shared-memory atomics in the loop.
// Tensor-memory map (512 columns): dy2 -> dg2 [0,128) | xhat2 [128,256) | dy1 -> dg1 [256,512) | xhat1 [0,256) |
// layer 3: dg3 [0,64), xhat3 [64,128) | da [0,64).  Every region is dead before its next writer is issued (one tile in flight).
// =============================================================================================
constexpr uint32_t SM2B_Z = SM_W2 + 64 * 128 * 2;                       // dz tile (next GEMM's A operand)   64 KB
constexpr uint32_t SM2B_SCR = SM2B_Z + 128 * 256 * 2;                   // 16 warps x 2 KB transposition scratch
constexpr uint32_t SM2B_PAR = SM2B_SCR + 16 * 2048;
constexpr uint32_t SM2B_STAT = SM2B_PAR + PAR_COUNT * 4;                // [128][4][2] floats
constexpr uint32_t SM2B_ACC = SM2B_STAT + 128 * MLP_NH * 2 * 4;
constexpr uint32_t SM2B_TOTAL = SM2B_ACC + ACC_COUNT * 4;
static_assert(SM2B_TOTAL <= 232448, "mlp_tc_bwd2: shared memory");

// per-warp scratch [32 rows][16 cols] fp32; the 16-byte chunk j of row r sits at position j ^ ((r >> 1) & 3)
struct ColScratch {
  float* base;          // the CTA's dynamic shared memory as floats (the compiler keeps shared-window addressing)
  uint32_t st[4];       // word index of this lane's (= row's) four chunks
  uint32_t ld[2][4];    // read bases (word index): [parity of the step][(step >> 1) & 3]
  __device__ __forceinline__ void init(float* smem_f, uint32_t word0, int lane) {
    base = smem_f;
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) st[j] = word0 + lane * 16 + ((j ^ sw) << 2);
    const int c = lane & 15, hf = lane >> 4;
    // step i: the lower half warp reads row i, the upper one row 16 + (i ^ 1) (opposite bank half: no conflicts)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t off = (((c >> 2) ^ k) << 2) + (c & 3);
      ld[0][k] = word0 + (hf ? 17 * 16 : 0) + off;      // even steps
      ld[1][k] = word0 + (hf ? 15 * 16 : 0) + off;      // odd steps
    }
  }
  __device__ __forceinline__ void put(const float (&v)[16]) const {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(base + st[j]) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  // sum over this lane's 16 rows of column (lane & 15)
  __device__ __forceinline__ float get() const {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float x = *reinterpret_cast<volatile const float*>(base + ld[i & 1][(i >> 1) & 3] + i * 16);
      if (i & 1) s1 += x;
      else s0 += x;
    }
    return s0 + s1;
  }
};

// one layer of the chain for this thread's row and column part, 16 columns at a time
template <int C, bool FROM_TMEM>
__device__ __forceinline__ void mlp_bwd2_layer(uint32_t tmem_dg, uint32_t tmem_xh, float dml, int q, int h, int lane, int rt,
                                               const uint8_t* __restrict__ r_img, uint4 (&rw)[2], const float* __restrict__ gam,
                                               float* s_stat, const ColScratch& cs, float* accd, float* acct, float* accz,
                                               uint8_t* dztile, uint8_t* __restrict__ dz_img, float2 ms) {
  constexpr int PART = C / MLP_NH, NSUB = PART / 16;
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
  const uint32_t a_dg = tmem_dg + lane_addr + h * PART, a_xh = tmem_xh + lane_addr + h * PART;
  const float rstd = ms.y, nmr = -ms.x * ms.y;      // LayerNorm (mean, rstd) of the row, saved by the forward
  float s1p[2] = {0.f, 0.f}, s2p[2] = {0.f, 0.f};
  // ---- pass B: dg, xhat -> tensor memory; row sums; column sums of d and d * xhat ----
#pragma unroll
  for (int sub = 0; sub < NSUB; ++sub) {
    const int c0 = h * PART + sub * 16;
    uint4 rn[2];
    if (sub + 1 < NSUB) {      // next 16 saved values of the row: two 16-byte pieces, 128 bytes apart in the tile image
      const uint8_t* p = r_img + tile_off(rt, c0 + 16, C);
      rn[0] = __ldg(reinterpret_cast<const uint4*>(p));
      rn[1] = __ldg(reinterpret_cast<const uint4*>(p + 128));
    }
    float d[16], xh[16], t[16];
    if (FROM_TMEM) {
      tmem_ld16(a_dg + sub * 16, d);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) d[i] = dml;      // layer 3: see mlp_bwd_layer
    }
    const uint32_t w8[8] = {rw[0].x, rw[0].y, rw[0].z, rw[0].w, rw[1].x, rw[1].y, rw[1].z, rw[1].w};
#pragma unroll
    for (int i2 = 0; i2 < 8; ++i2) {
      const uint32_t w = w8[i2], lo = w << 16;
      xh[2 * i2] = fmaf(fabsf(__uint_as_float(lo)), rstd, nmr);
      xh[2 * i2 + 1] = fmaf(fabsf(__uint_as_float(w & 0xffff0000u)), rstd, nmr);
      d[2 * i2] = (int32_t)lo < 0 ? 0.f : d[2 * i2];
      d[2 * i2 + 1] = (int32_t)w < 0 ? 0.f : d[2 * i2 + 1];
    }
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
      const float4 g4 = *reinterpret_cast<const float4*>(gam + c0 + 4 * i4);
      const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = 4 * i4 + k;
        const float dg = d[i] * gg[k];
        s1p[k & 1] += dg;
        s2p[k & 1] = fmaf(dg, xh[i], s2p[k & 1]);
        t[i] = dg;
      }
    }
    tmem_st16(a_dg + sub * 16, t);
    tmem_st16(a_xh + sub * 16, xh);
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = d[i] * xh[i];
    cs.put(d);
    __syncwarp();
    accd[sub] += cs.get();
    __syncwarp();
    cs.put(t);
    __syncwarp();
    acct[sub] += cs.get();
    __syncwarp();
    if (sub + 1 < NSUB) {
      rw[0] = rn[0];
      rw[1] = rn[1];
    }
  }
  tmem_st_wait();
  s_stat[(rt * MLP_NH + h) * 2 + 0] = s1p[0] + s1p[1];
  s_stat[(rt * MLP_NH + h) * 2 + 1] = s2p[0] + s2p[1];
  quarter_sync(q);          // only the four warps that share these 32 rows exchange their sums
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < MLP_NH; ++k) {
    s1 += s_stat[(rt * MLP_NH + k) * 2 + 0];
    s2 += s_stat[(rt * MLP_NH + k) * 2 + 1];
  }
  const float nm2 = -s2 * (1.0f / C), nm1r = -s1 * (1.0f / C) * rstd;
  // ---- pass C: dz = relu' * rstd * (dg - mean(dg) - xhat * mean(dg * xhat)) ----
#pragma unroll
  for (int sub = 0; sub < NSUB; ++sub) {
    const int c0 = h * PART + sub * 16;
    float dg[16], xh[16];
    tmem_ld16(a_dg + sub * 16, dg);
    tmem_ld16(a_xh + sub * 16, xh);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float dr = fmaf(fmaf(xh[i], nm2, dg[i]), rstd, nm1r);
      dg[i] = xh[i] > nmr ? dr : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint4 pk = make_uint4(pack_bf16(dg[8 * j], dg[8 * j + 1]), pack_bf16(dg[8 * j + 2], dg[8 * j + 3]),
                                  pack_bf16(dg[8 * j + 4], dg[8 * j + 5]), pack_bf16(dg[8 * j + 6], dg[8 * j + 7]));
      const uint32_t off = tile_off(rt, c0 + 8 * j, C);
      *reinterpret_cast<uint4*>(dztile + off) = pk;
      *reinterpret_cast<uint4*>(dz_img + off) = pk;
    }
    if constexpr (C == 64) {      // the 256- and 128-wide layers get their bias gradient from the wgrad kernel's ones-GEMM
      cs.put(dg);
      __syncwarp();
      accz[sub] += cs.get();
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(MLP_THREADS, 1) mlp_tc_bwd2_kernel(MlpBwdArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, h = warp >> 2;
  float* par = reinterpret_cast<float*>(smem + SM2B_PAR);
  float* s_stat = reinterpret_cast<float*>(smem + SM2B_STAT);
  float* s_acc = reinterpret_cast<float*>(smem + SM2B_ACC);
  const float* P = A.dense;

  load_weight_image<256, 64>(smem + SM_W0, P + NCF_OFF(NCF_P_MLP0_W), K0, tid, MLP_THREADS);
  load_weight_image<128, 256>(smem + SM_W1, P + NCF_OFF(NCF_P_MLP1_W), H1, tid, MLP_THREADS);
  load_weight_image<64, 128>(smem + SM_W2, P + NCF_OFF(NCF_P_MLP2_W), H2, tid, MLP_THREADS);
  const float sc0 = A.rng[0].thresh ? A.rng[0].scale : 1.f, sc1 = A.rng[1].thresh ? A.rng[1].scale : 1.f,
              sc2 = A.rng[2].thresh ? A.rng[2].scale : 1.f;
  for (int i = tid; i < 256; i += MLP_THREADS) {
    par[PAR_G0 + i] = P[NCF_OFF(NCF_P_LN0_W) + i] * sc0;
    if (i < 128) par[PAR_G1 + i] = P[NCF_OFF(NCF_P_LN1_W) + i] * sc1;
    if (i < 64) {
      par[PAR_WOUT + i] = P[NCF_OFF(NCF_P_MLP_OUT_W) + i];
      par[PAR_G2 + i] = P[NCF_OFF(NCF_P_LN2_W) + i] * sc2 * P[NCF_OFF(NCF_P_MLP_OUT_W) + i];     // see mlp_bwd_layer<64, false>
    }
  }
  for (int i = tid; i < ACC_COUNT; i += MLP_THREADS) s_acc[i] = 0.f;
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  pdl_wait();      // prologue done: from here on the kernel reads what the previous kernels of the step wrote
  const uint32_t tmem = tmem_slot;
  const uint32_t sW0 = smem_addr(smem + SM_W0), sW1 = smem_addr(smem + SM_W1), sW2 = smem_addr(smem + SM_W2);
  const uint32_t sZ = smem_addr(smem + SM2B_Z);
  uint8_t* ztile = smem + SM2B_Z;
  ColScratch cs;
  cs.init(reinterpret_cast<float*>(smem), (SM2B_SCR + warp * 2048) / 4, lane);
  // column-sum partials of this lane = (column lane & 15 of each 16-column group, row half lane >> 4), all tiles
  float ad1[4] = {0.f, 0.f, 0.f, 0.f}, at1[4] = {0.f, 0.f, 0.f, 0.f}, ad2[2] = {0.f, 0.f}, at2[2] = {0.f, 0.f}, ad3[1] = {0.f},
        at3[1] = {0.f}, az3[1] = {0.f};
  uint32_t phase = 0;
  const int rt = q * 32 + lane;

  const int64_t ntiles = (A.N + TCM_ROWS - 1) / TCM_ROWS;
  auto load16 = [&](const uint8_t* img, int C, int c0, uint4 (&w)[2]) {
    const uint8_t* p = img + tile_off(rt, c0, C);
    w[0] = __ldg(reinterpret_cast<const uint4*>(p));
    w[1] = __ldg(reinterpret_cast<const uint4*>(p + 128));
  };
  // everything layer 3 needs from global memory is requested ONE TILE AHEAD (while the previous tile waits for its last
  // GEMM), and the LayerNorm statistics of layers 2 / 1 at the top of the tile: no phase starts with a cold load
  uint4 rw3[2];
  float2 ms3 = make_float2(0.f, 1.f);
  float dml = 0.f;
  auto prefetch_l3 = [&](int64_t t) {
    load16(reinterpret_cast<const uint8_t*>(A.r3) + t * (128 * 64 * 2), 64, h * 16, rw3);
    ms3 = __ldg(reinterpret_cast<const float2*>(A.st3 + t * 256 + rt * 2));
    const int64_t g = t * TCM_ROWS + rt;
    dml = g < A.N ? __ldg(A.d_mlp_pred + g) : 0.f;
  };
  if ((int64_t)blockIdx.x < ntiles) prefetch_l3(blockIdx.x);
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * TCM_ROWS;
    const int64_t avail = min((int64_t)TCM_ROWS, A.N - row0);
    const int64_t grow = row0 + rt;
    const bool live = rt < avail;
    const uint8_t* r1i = reinterpret_cast<const uint8_t*>(A.r1) + tile * (128 * 256 * 2);
    const uint8_t* r2i = reinterpret_cast<const uint8_t*>(A.r2) + tile * (128 * 128 * 2);
    const uint8_t* r3i = reinterpret_cast<const uint8_t*>(A.r3) + tile * (128 * 64 * 2);
    uint8_t* z1i = reinterpret_cast<uint8_t*>(A.dz1) + tile * (128 * 256 * 2);
    uint8_t* z2i = reinterpret_cast<uint8_t*>(A.dz2) + tile * (128 * 128 * 2);
    uint8_t* z3i = reinterpret_cast<uint8_t*>(A.dz3) + tile * (128 * 64 * 2);
    const int64_t nt = tile + gridDim.x;
    if (tid == 0 && nt < ntiles) {       // pull the NEXT tile's saved tensors into L2 while this one is processed
      bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(A.r1) + nt * (128 * 256 * 2), 128 * 256 * 2);
      bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(A.r2) + nt * (128 * 128 * 2), 128 * 128 * 2);
      bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(A.r3) + nt * (128 * 64 * 2), 128 * 64 * 2);
    }
    uint4 rw2[2], rw1[2];
    load16(r2i, 128, h * 32, rw2);      // consumed after the first MMA: its latency hides behind layer 3
    const float2 ms2 = __ldg(reinterpret_cast<const float2*>(A.st2 + tile * 256 + rt * 2));
    mlp_bwd2_layer<64, false>(tmem + 0, tmem + 64, dml, q, h, lane, rt, r3i, rw3, par + PAR_G2, s_stat, cs, ad3, at3, az3, ztile, z3i,
                              ms3);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();   // dy2[128x128] = dz3[128x64] . W2[64x128]
      issue_gemm(tmem + 0, sZ, 128, 64 * 16, 256, sW2, 128 * 16, 128, 2 * 128 * 16, make_idesc(128, 128, false, true), 4, false);
      mma_commit(&bar);
    }
    load16(r1i, 256, h * 64, rw1);      // consumed after the second MMA
    const float2 ms1 = __ldg(reinterpret_cast<const float2*>(A.st1 + tile * 256 + rt * 2));
    mbar_wait(&bar, phase);             // every warp polls for itself
    phase ^= 1;
    fence_after_sync();
    mlp_bwd2_layer<128, true>(tmem + 0, tmem + 128, 0.f, q, h, lane, rt, r2i, rw2, par + PAR_G1, s_stat, cs, ad2, at2, nullptr, ztile,
                              z2i, ms2);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();   // dy1[128x256] = dz2[128x128] . W1[128x256]
      issue_gemm(tmem + 256, sZ, 128, 128 * 16, 256, sW1, 256 * 16, 128, 2 * 256 * 16, make_idesc(128, 256, false, true), 8, false);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();
    mlp_bwd2_layer<256, true>(tmem + 256, tmem + 0, 0.f, q, h, lane, rt, r1i, rw1, par + PAR_G0, s_stat, cs, ad1, at1, nullptr, ztile,
                              z1i, ms1);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();   // da[128x64] = dz1[128x256] . W0[256x64]
      issue_gemm(tmem + 0, sZ, 128, 256 * 16, 256, sW0, 64 * 16, 128, 2 * 64 * 16, make_idesc(128, 64, false, true), 16, false);
      mma_commit(&bar);
    }
    if (nt < ntiles) prefetch_l3(nt);   // in flight while the last GEMM of this tile runs
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();
    {
      float v[16];
      tmem_ld16(tmem + 0 + ((uint32_t)(q * 32) << 16) + h * 16, v);
      if (live) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_row4(A.da, grow, h * 16 + 4 * j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]), A.da_bf16);
      }
    }
    // No CTA barrier here: the next tile's layer 3 touches, before its own barrier, only this thread's tensor-memory lanes
    // and columns (the da read above is the same thread's), this warp's scratch, and s_stat / the dz tile, whose last
    // readers passed the barrier in front of the da GEMM (and its completion wait) above.
    fence_before_sync();
  }
  __syncthreads();
  // the register partials of all warps -> s_acc ([dgamma | dbeta | dbias] per layer, as the first version keeps them)
  {
    const int c = lane & 15;
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) {
      atomicAdd(s_acc + ACC_L0 + h * 64 + sub * 16 + c, at1[sub]);
      atomicAdd(s_acc + ACC_L0 + 256 + h * 64 + sub * 16 + c, ad1[sub]);
    }
#pragma unroll
    for (int sub = 0; sub < 2; ++sub) {
      atomicAdd(s_acc + ACC_L1 + h * 32 + sub * 16 + c, at2[sub]);
      atomicAdd(s_acc + ACC_L1 + 128 + h * 32 + sub * 16 + c, ad2[sub]);
    }
    atomicAdd(s_acc + ACC_L2 + h * 16 + c, at3[0]);
    atomicAdd(s_acc + ACC_L2 + 64 + h * 16 + c, ad3[0]);
    atomicAdd(s_acc + ACC_L2 + 128 + h * 16 + c, az3[0]);
  }
  __syncthreads();
  float* dg = A.dense_grad;
  for (int i = tid; i < ACC_COUNT; i += MLP_THREADS) {
    int64_t off;
    int k = i;
    float sc;          // d gamma / d beta sums were taken without the dropout scale; the bias sums carry it already
    if (k < ACC_L1) {
      if (k >= 512) continue;        // bias of the 256-wide layer: wgrad kernel
      off = k < 256 ? NCF_OFF(NCF_P_LN0_W) + k : NCF_OFF(NCF_P_LN0_B) + (k - 256);
      sc = sc0;
    } else if (k < ACC_L2) {
      k -= ACC_L1;
      if (k >= 256) continue;        // bias of the 128-wide layer: wgrad kernel
      off = k < 128 ? NCF_OFF(NCF_P_LN1_W) + k : NCF_OFF(NCF_P_LN1_B) + (k - 128);
      sc = sc1;
    } else {
      k -= ACC_L2;
      off = k < 64 ? NCF_OFF(NCF_P_LN2_W) + k : k < 128 ? NCF_OFF(NCF_P_LN2_B) + (k - 64) : NCF_OFF(NCF_P_MLP2_B) + (k - 128);
      sc = k < 128 ? sc2 * par[PAR_WOUT + (k & 63)] : 1.f;
      if (k < 64)
        atomicAdd(dg + NCF_OFF(NCF_P_MLP_OUT_W) + k,
                  sc2 * fmaf(P[NCF_OFF(NCF_P_LN2_W) + k], s_acc[i], P[NCF_OFF(NCF_P_LN2_B) + k] * s_acc[i + 64]));
    }
    atomicAdd(dg + off, s_acc[i] * sc);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// =============================================================================================
// MLP tower backward, part 2: weight gradients.  Per 128-row tile the saved tiles are read as MN-major
// operands (K = the 128 rows) and accumulated over ALL tiles of the CTA in TMEM:
//   dW2^T [128 x 64] += y2^T . dz3      dW1 [128 x 256] += dz2^T . y1      dW0 [256 x 64] += dz1^T . a
// =============================================================================================
constexpr uint32_t SMW_Y2 = 0;                         // [128][128] bf16 32 KB
constexpr uint32_t SMW_Z3 = SMW_Y2 + 128 * 128 * 2;    // [128][64]       16 KB
constexpr uint32_t SMW_Z2 = SMW_Z3 + 128 * 64 * 2;     // [128][128]      32 KB
constexpr uint32_t SMW_Y1 = SMW_Z2 + 128 * 128 * 2;    // [128][256]      64 KB
constexpr uint32_t SMW_Z1 = SMW_Y1 + 128 * 256 * 2;    // [128][256]      64 KB
constexpr uint32_t SMW_A = SMW_Z1 + 128 * 256 * 2;     // [128][64]       16 KB
constexpr uint32_t SMW_ONES = SMW_A + 128 * 64 * 2;    // [16][16] bf16 ones, 512 B: every K step of the bias GEMMs reads it
constexpr uint32_t SMW_TOTAL = SMW_ONES + 512;         // 224.5 KB

__device__ __forceinline__ void copy_tile_image(uint8_t* dst, const __nv_bfloat16* __restrict__ img, int64_t tile, int bytes,
                                                int tid, int nthreads) {
  const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(img) + tile * bytes);
  for (int i = tid; i < bytes / 16; i += nthreads) reinterpret_cast<uint4*>(dst)[i] = __ldg(src + i);
}

// accumulator words per CTA: 448 TMEM columns x 128 lanes of weight gradients + column 0 of the three 16-column bias
// accumulators (NCF_P_MLP0_B rows 0-127, rows 128-255, NCF_P_MLP1_B)
constexpr int WG_PART = MLP_WG_PART_COLS * 128;

struct MlpWgradArgs {
  const __nv_bfloat16* a_img;                            // MLP input, bf16 tile image
  const __nv_bfloat16 *y1, *y2, *dz1, *dz2, *dz3;
  float* dense_grad;
  float* partial;                                         // [grid][WG_PART]
  int64_t N;
};

// Warp-specialised: one producer thread streams the tile images with TMA bulk copies into two stages
// (A: y2, dz3, dz2, y1 = 144 KB; B: dz1, a = 80 KB), one MMA thread consumes them; a stage is refilled
// as soon as the MMAs that read it have committed, so loads and tensor work overlap.
__global__ void __launch_bounds__(TCM_THREADS, 1) mlp_tc_wgrad_kernel(MlpWgradArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t fullA, fullB, emptyA, emptyB;
  __shared__ uint32_t tmem_slot;
  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, h = warp >> 2;
  if (tid == 0) {
    mbar_init(&fullA, 1);
    mbar_init(&fullB, 1);
    mbar_init(&emptyA, 1);
    mbar_init(&emptyB, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  // Bias gradients = column sums of the dz tiles = dz^T . ones: one more N = 16 GEMM per dz operand that is in shared
  // memory anyway (B = a 16 x 16 block of ones, the same block for every K step), instead of shuffle-transposing
  // every tile in the backward kernel's epilogue.
  if (tid < 128) reinterpret_cast<uint32_t*>(smem + SMW_ONES)[tid] = 0x3F803F80u;
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  pdl_wait();
  const uint32_t tmem = tmem_slot;
  const int64_t ntiles = (A.N + TCM_ROWS - 1) / TCM_ROWS;
  const int64_t my_tiles = (int64_t)blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  constexpr uint32_t B_Y2 = 128 * 128 * 2, B_Z3 = 128 * 64 * 2, B_Z2 = 128 * 128 * 2, B_Y1 = 128 * 256 * 2,
                     B_Z1 = 128 * 256 * 2, B_A = 128 * 64 * 2;
  if (tid == 0) {
    // ---- producer --------------------------------------------------------------------------
    const uint8_t *y2 = (const uint8_t*)A.y2, *z3 = (const uint8_t*)A.dz3, *z2 = (const uint8_t*)A.dz2,
                  *y1 = (const uint8_t*)A.y1, *z1 = (const uint8_t*)A.dz1, *ai = (const uint8_t*)A.a_img;
    for (int64_t k = 0; k < my_tiles; ++k) {
      // LAST tiles first: the dz tiles the input-gradient kernel wrote last are the ones still in the 126 MB L2
      const int64_t tile = ntiles - 1 - (blockIdx.x + k * gridDim.x);
      if (k > 0) mbar_wait(&emptyA, (k - 1) & 1);
      mbar_arrive_expect_tx(&fullA, B_Y2 + B_Z3 + B_Z2 + B_Y1);
      bulk_g2s(smem + SMW_Y2, y2 + tile * B_Y2, B_Y2, &fullA);
      bulk_g2s(smem + SMW_Z3, z3 + tile * B_Z3, B_Z3, &fullA);
      bulk_g2s(smem + SMW_Z2, z2 + tile * B_Z2, B_Z2, &fullA);
      bulk_g2s(smem + SMW_Y1, y1 + tile * B_Y1, B_Y1, &fullA);
      if (k > 0) mbar_wait(&emptyB, (k - 1) & 1);
      mbar_arrive_expect_tx(&fullB, B_Z1 + B_A);
      bulk_g2s(smem + SMW_Z1, z1 + tile * B_Z1, B_Z1, &fullB);
      bulk_g2s(smem + SMW_A, ai + tile * B_A, B_A, &fullB);
    }
  } else if (tid == 32) {
    // ---- MMA issuer ---------------------------------------------------------------------------
    const uint32_t sY2 = smem_addr(smem + SMW_Y2), sZ3 = smem_addr(smem + SMW_Z3), sZ2 = smem_addr(smem + SMW_Z2);
    const uint32_t sY1 = smem_addr(smem + SMW_Y1), sZ1 = smem_addr(smem + SMW_Z1), sA = smem_addr(smem + SMW_A);
    const uint32_t sOnes = smem_addr(smem + SMW_ONES);
    constexpr uint32_t idesc_b = make_idesc(128, 16, true, true);
    for (int64_t k = 0; k < my_tiles; ++k) {
      const bool acc = k > 0;
      mbar_wait(&fullA, k & 1);
      fence_after_sync();
      // dW2^T[k_in 128][n_out 64] += y2^T . dz3
      issue_gemm(tmem + 0, sY2, 128 * 16, 128, 2 * 128 * 16, sZ3, 64 * 16, 128, 2 * 64 * 16, make_idesc(128, 64, true, true), 8, acc);
      // dW1[n_out 128][k_in 256] += dz2^T . y1
      issue_gemm(tmem + 64, sZ2, 128 * 16, 128, 2 * 128 * 16, sY1, 256 * 16, 128, 2 * 256 * 16, make_idesc(128, 256, true, true), 8, acc);
      // bias gradient of the 256 -> 128 Linear (NCF_P_MLP1_B)[n_out 128] += dz2^T . 1
      issue_gemm(tmem + 480, sZ2, 128 * 16, 128, 2 * 128 * 16, sOnes, 16 * 16, 128, 0, idesc_b, 8, acc);
      mma_commit(&emptyA);
      mbar_wait(&fullB, k & 1);
      fence_after_sync();
      // dW0[n_out 256][k_in 64] += dz1^T . a   (two M = 128 halves: +16 MN groups = 2048 B)
      issue_gemm(tmem + 320, sZ1, 256 * 16, 128, 2 * 256 * 16, sA, 64 * 16, 128, 2 * 64 * 16, make_idesc(128, 64, true, true), 8, acc);
      issue_gemm(tmem + 384, sZ1 + 2048, 256 * 16, 128, 2 * 256 * 16, sA, 64 * 16, 128, 2 * 64 * 16, make_idesc(128, 64, true, true), 8, acc);
      // bias gradient of the 96 -> 256 Linear (NCF_P_MLP0_B)[n_out 256] += dz1^T . 1   (two halves)
      issue_gemm(tmem + 448, sZ1, 256 * 16, 128, 2 * 256 * 16, sOnes, 16 * 16, 128, 0, idesc_b, 8, acc);
      issue_gemm(tmem + 464, sZ1 + 2048, 256 * 16, 128, 2 * 256 * 16, sOnes, 16 * 16, 128, 0, idesc_b, 8, acc);
      mma_commit(&emptyB);
    }
    if (my_tiles > 0) {          // all accumulation done before the flush below
      mbar_wait(&emptyA, (my_tiles - 1) & 1);
      mbar_wait(&emptyB, (my_tiles - 1) & 1);
    }
  }
  __syncthreads();
  {
    // flush: this CTA's accumulators go to its own slice of a partial buffer (coalesced stores, no atomics);
    // mlp_wgrad_reduce_kernel adds the slices into the gradient.  CTAs without tiles store zeros.
    fence_after_sync();
    float* part = A.partial + (int64_t)blockIdx.x * WG_PART;
    const int lane_row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    for (int ch = 0; ch < 7; ++ch) {          // 448 columns: this thread's half = 7 chunks of 32
      float v[32];
      const int c0 = h * 224 + ch * 32;
      tmem_ld32(tmem + lane_addr + c0, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) part[(int64_t)(c0 + i) * 128 + lane_row] = my_tiles > 0 ? v[i] : 0.f;
    }
    if (h == 0) {                             // bias accumulators: all 16 columns hold the same sum
      for (int j = 0; j < 3; ++j) {
        float v[16];
        tmem_ld16(tmem + lane_addr + 448 + 16 * j, v);
        part[(int64_t)(448 + j) * 128 + lane_row] = my_tiles > 0 ? v[0] : 0.f;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// dense_grad += sum over CTAs of the partial accumulators; e = column * 128 + lane
__global__ void __launch_bounds__(256) mlp_wgrad_reduce_kernel(const float* __restrict__ partial, int nparts,
                                                               float* __restrict__ dg) {
  pdl_launch_dependents();
  pdl_wait();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= WG_PART) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += partial[(int64_t)p * WG_PART + e];
  const int col = e >> 7, lane = e & 127;
  int64_t off;
  if (col >= 448) off = col == 450 ? NCF_OFF(NCF_P_MLP1_B) + lane : NCF_OFF(NCF_P_MLP0_B) + (col - 448) * 128 + lane;
  else if (col < 64) off = NCF_OFF(NCF_P_MLP2_W) + (int64_t)col * H2 + lane;                  // dW2^T: lane = k_in
  else if (col < 320) off = NCF_OFF(NCF_P_MLP1_W) + (int64_t)lane * H1 + (col - 64);           // dW1: lane = n_out
  else off = NCF_OFF(NCF_P_MLP0_W) + (int64_t)(((col - 320) >> 6) * 128 + lane) * K0 + ((col - 320) & 63);
  dg[off] += s;
}

int mlp_tc_backward(const ncf_run_cfg& cfg, const float* dense, float* dense_grad, int64_t N, TowerWs& w, cudaStream_t st,
                    cudaStream_t side, int side_sms, cudaEvent_t fork, cudaEvent_t join) {
  if (N == 0) return NCF_OK;
  static bool configured = false;
  if (!configured) {
    NCF_CUDA(cudaFuncSetAttribute(mlp_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMB_TOTAL));
    NCF_CUDA(cudaFuncSetAttribute(mlp_tc_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM2B_TOTAL));
    NCF_CUDA(cudaFuncSetAttribute(mlp_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMW_TOTAL));
    configured = true;
  }
  const int64_t ntiles = (N + TCM_ROWS - 1) / TCM_ROWS;
  const int grid = even_grid(ntiles, tower_sms());
  MlpBwdArgs B{};
  B.dense = dense;
  B.dense_grad = dense_grad;
  B.da = w.g64a;
  B.da_bf16 = tower_bf16_rows(cfg);
  B.d_mlp_pred = w.d_mlp;
  B.r1 = (const __nv_bfloat16*)w.r1b;
  B.r2 = (const __nv_bfloat16*)w.r2b;
  B.r3 = (const __nv_bfloat16*)w.r3b;
  B.st1 = w.st1;
  B.st2 = w.st2;
  B.st3 = w.st3;
  B.dz1 = (__nv_bfloat16*)w.dz1b;
  B.dz2 = (__nv_bfloat16*)w.dz2b;
  B.dz3 = (__nv_bfloat16*)w.dz3b;
  B.N = N;
  for (int l = 0; l < 3; ++l) B.rng[l] = make_rng(cfg, 1 + l);
  static const int bwd_variant = getenv("NCF_MLP_BWD") ? atoi(getenv("NCF_MLP_BWD")) : 2;     // A/B switch
  if (bwd_variant == 1) mlp_tc_bwd_kernel<<<grid, MLP_THREADS, SMB_TOTAL, st>>>(B);
  else NCF_CUDA(launch_pdl(PDL_MLP_BWD, mlp_tc_bwd2_kernel, dim3(grid), dim3(MLP_THREADS), SM2B_TOTAL, st, B));
  NCF_LAUNCH_CHECK();
  MlpWgradArgs W{};
  W.a_img = (const __nv_bfloat16*)w.a_img;
  W.y1 = (const __nv_bfloat16*)w.y1b;
  W.y2 = (const __nv_bfloat16*)w.y2b;
  W.dz1 = B.dz1;
  W.dz2 = B.dz2;
  W.dz3 = B.dz3;
  W.dense_grad = dense_grad;
  W.partial = w.wg_partial;
  W.N = N;
  cudaStream_t wst = st;
  int wgrid = grid;
  if (side && side_sms > 0 && fork) {
    NCF_CUDA(cudaEventRecord(fork, st));
    NCF_CUDA(cudaStreamWaitEvent(side, fork, 0));
    wst = side;
    wgrid = (int)std::min<int64_t>(ntiles, side_sms);
  }
  NCF_CUDA(launch_pdl(PDL_MLP_WGRAD, mlp_tc_wgrad_kernel, dim3(wgrid), dim3(TCM_THREADS), SMW_TOTAL, wst, W));
  NCF_LAUNCH_CHECK();
  NCF_CUDA(launch_pdl(PDL_MLP_WGRAD, mlp_wgrad_reduce_kernel, dim3((WG_PART + 255) / 256), dim3(256), 0, wst, (const float*)w.wg_partial, wgrid, dense_grad));
  NCF_LAUNCH_CHECK();
  if (wst != st && join) NCF_CUDA(cudaEventRecord(join, wst));
  return NCF_OK;
}

}  // namespace ncf

using namespace ncf;

template <int MODE, int K, int N>
static int run_selftest(const float* A, const float* B, float* D, cudaStream_t st) {
  constexpr int AC = MODE == 2 ? 128 : K;
  constexpr int BR = (MODE == 0 || MODE == 3) ? N : (MODE == 1 ? K : 128), BC = (MODE == 0 || MODE == 3) ? K : N;
  const int smem = 128 * AC * 2 + BR * BC * 2;
  NCF_CUDA(cudaFuncSetAttribute(tc_selftest_kernel<MODE, K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc_selftest_kernel<MODE, K, N><<<1, 128, smem, st>>>(A, B, D);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

// A, B, D: fp32 device buffers; shapes by mode (see kernel).  (mode, K, N) in {(0,64,256),(0,256,128),
// (1,64,128),(1,128,256),(2,128,64),(2,128,256),(3,256,128),(3,128,64)}.
extern "C" int ncf_tc_selftest(int32_t mode, int32_t K, int32_t N, const float* A, const float* B, float* D, void* stream) {
  NCF_REQUIRE(A && B && D, "tc_selftest: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0 && K == 64 && N == 256) return run_selftest<0, 64, 256>(A, B, D, st);
  if (mode == 0 && K == 256 && N == 128) return run_selftest<0, 256, 128>(A, B, D, st);
  if (mode == 1 && K == 64 && N == 128) return run_selftest<1, 64, 128>(A, B, D, st);
  if (mode == 1 && K == 128 && N == 256) return run_selftest<1, 128, 256>(A, B, D, st);
  if (mode == 2 && K == 128 && N == 64) return run_selftest<2, 128, 64>(A, B, D, st);
  if (mode == 2 && K == 128 && N == 256) return run_selftest<2, 128, 256>(A, B, D, st);
  if (mode == 3 && K == 256 && N == 128) return run_selftest<3, 256, 128>(A, B, D, st);
  if (mode == 3 && K == 128 && N == 64) return run_selftest<3, 128, 64>(A, B, D, st);
  set_error("tc_selftest: unsupported (mode,K,N) = (%d,%d,%d)", mode, K, N);
  return NCF_ERR_UNSUPPORTED;
}
