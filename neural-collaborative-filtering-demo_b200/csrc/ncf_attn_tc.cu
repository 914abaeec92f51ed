// Fused user-product attention block on tcgen05 (architecture.py:18-57 as used at :204-208, 315-326),
// training shape S = 5 rows per interaction (1 positive + 4 negatives, data_prep.py:286-303).
//
//   forward   a = out_proj(softmax(q k^T / 4) v)  with q = q_proj(xu), [k|v] = [k_proj; v_proj](xp)
//   backward  (dxu, dxp, dWq, dbq, dWk|v, dbk|v, dWo, dbo) from da, RECOMPUTING q, k, v, the probabilities
//             and ctx from xu / xp (four small GEMMs) instead of saving them: the block keeps nothing
//             between forward and backward except its output tile image.
//
// One tile = 125 rows = 25 whole interactions (M = 128 with 3 idle rows), so the rows that attend to each
// other always share a CTA.  The projections run on the tensor core (bf16 operands, fp32 accumulators in
// TMEM); the 5 x 5 attention itself is CUDA-core work with thread = (row, head): k, v (and in the backward
// q, d ctx, p, ds) of the four other rows of the interaction are exchanged through shared memory.
#include "ncf_tower.cuh"
#include "ncf_umma.cuh"

namespace ncf {
using namespace umma;

constexpr int AT_S = 5;
constexpr int AT_RT = 125;          // live rows per tile
constexpr int AT_THREADS = 512;     // thread = (row, head): warp = 4 * head + TMEM lane quarter
constexpr int AT_XS = 72;           // bf16 elements per row of the exchange arrays (144 B rows: conflict-free 16 B reads)
constexpr int AT_PS = 21;           // floats per row of the probability exchange arrays (4 heads x 5 keys, +1 pad)

int64_t attn_tc_tiles(int64_t N) { return (align_up(N, 128) + AT_RT - 1) / AT_RT; }

// keep decisions of the 5 probabilities of one (interaction, head, query): element index e0 + j,
// the same stream as attn_core_fwd_kernel (ncf_tower_f32.cu).  Bit j = keep.
__device__ __forceinline__ uint32_t attn_keep5(const DropoutRng& rng, uint64_t e0) {
  if (rng.thresh == 0u) return 31u;
  const uint64_t g0 = e0 >> 3;
  const uint32_t o = (uint32_t)e0 & 7u;
  const uint4 a = rng.draw8(g0);
  uint4 b = a;
  if (o > 3u) b = rng.draw8(g0 + 1);
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < AT_S; ++j) {
    const uint32_t l = o + j;
    const uint4 r = l < 8u ? a : b;
    const uint32_t c = (l >> 1) & 3u;
    const uint32_t w = c == 0 ? r.x : c == 1 ? r.y : c == 2 ? r.z : r.w;
    bits |= (rng.keep16((l & 1u) ? (w >> 16) : (w & 0xffffu)) ? 1u : 0u) << j;
  }
  return bits;
}

__device__ __forceinline__ void load_bf16x16(const uint8_t* p, float (&v)[16]) {
  const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 16);
  float t[8];
  unpack_bf16x8(a, t);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = t[i];
  unpack_bf16x8(b, t);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[8 + i] = t[i];
}
__device__ __forceinline__ void store_bf16x16(uint8_t* p0, uint8_t* p1, const float (&v)[16]) {
  *reinterpret_cast<uint4*>(p0) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  *reinterpret_cast<uint4*>(p1) =
      make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
}
__device__ __forceinline__ float dot16(const float (&a)[16], const float (&b)[16]) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int c = 0; c < 16; c += 2) {
    s0 = fmaf(a[c], b[c], s0);
    s1 = fmaf(a[c + 1], b[c + 1], s1);
  }
  return s0 + s1;
}

// bf16 rows [N,64] -> operand tile with LAYOUT columns per row group: plain 16-byte copies (the four quarters of
// a warp read four adjacent 16-byte pieces of the same 8 rows; each quarter writes one 128-byte core matrix).
// ONES adds the [1 | 0...] chunk at column 64 (bias gradient through the GEMM).
template <int LAYOUT, bool ONES>
__device__ __forceinline__ void attn_fill_rows(uint8_t* tile, const __nv_bfloat16* __restrict__ src, int64_t row0, int64_t avail,
                                               int tid) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int c = tid + k * AT_THREADS;
    const int blk = c >> 5, l = c & 31;
    const int r = (blk >> 1) * 8 + (l & 7), j = (blk & 1) * 4 + (l >> 3);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < avail) v = __ldg(reinterpret_cast<const uint4*>(src + (row0 + r) * 64 + 8 * j));
    *reinterpret_cast<uint4*>(tile + tile_off(r, 8 * j, LAYOUT)) = v;
  }
  if (ONES && tid < 128) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (tid < avail) v.x = 0x00003f80u;     // bf16(1.0): column 64 = 1, columns 65..71 = 0
    *reinterpret_cast<uint4*>(tile + tile_off(tid, 64, LAYOUT)) = v;
  }
}

// =============================================================================================
// forward
// =============================================================================================
constexpr uint32_t AF_WQ = 0;                          // [64][64]   bf16  8 KB
constexpr uint32_t AF_WKV = AF_WQ + 64 * 64 * 2;       // [128][64]       16 KB
constexpr uint32_t AF_WO = AF_WKV + 128 * 64 * 2;      // [64][64]         8 KB
constexpr uint32_t AF_XU = AF_WO + 64 * 64 * 2;        // [128][64]       16 KB   xu tile, later the ctx tile
constexpr uint32_t AF_XP = AF_XU + 128 * 64 * 2;       // [128][64]       16 KB
constexpr uint32_t AF_KB = AF_XP + 128 * 64 * 2;       // [128][72]       18 KB   k rows (bf16) for the exchange
constexpr uint32_t AF_VB = AF_KB + 128 * AT_XS * 2;    // [128][72]       18 KB
constexpr uint32_t AF_BIAS = AF_VB + 128 * AT_XS * 2;  // bq | bk | bv | bo
constexpr uint32_t AF_TOTAL = AF_BIAS + 256 * 4;

struct AttnFwdArgs {
  const __nv_bfloat16 *xu, *xp;     // [N,64] bf16 rows: mlp_norm(user row), mlp_norm(item row)
  const float* dense;
  void* a_img;              // out: bf16 tile image [ceil(N/128)][128 x 64]
  int64_t N;
  DropoutRng rng;
};

__device__ __forceinline__ void attn_load_params(uint8_t* smem, uint32_t off_wq, uint32_t off_wkv, uint32_t off_wo, float* bias,
                                                 const float* __restrict__ P, int tid) {
  fill_tile_f32<64>(smem + off_wq, P + NCF_OFF(NCF_P_Q_W), 64, 0, 64, 64, tid, AT_THREADS);
  fill_tile_f32<64>(smem + off_wkv, P + NCF_OFF(NCF_P_K_W), 64, 0, 128, 128, tid, AT_THREADS);    // k_proj and v_proj are adjacent
  fill_tile_f32<64>(smem + off_wo, P + NCF_OFF(NCF_P_O_W), 64, 0, 64, 64, tid, AT_THREADS);
  if (tid < 256) {
    const int i = tid & 63;
    const int64_t o = tid < 64 ? NCF_OFF(NCF_P_Q_B) : tid < 128 ? NCF_OFF(NCF_P_K_B) : tid < 192 ? NCF_OFF(NCF_P_V_B) : NCF_OFF(NCF_P_O_B);
    bias[tid] = P[o + i];
  }
}

__global__ void __launch_bounds__(AT_THREADS, 2) attn_tc_fwd_kernel(AttnFwdArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, h = warp >> 2;
  const int row = q * 32 + lane;
  float* bias = reinterpret_cast<float*>(smem + AF_BIAS);
  attn_load_params(smem, AF_WQ, AF_WKV, AF_WO, bias, A.dense, tid);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  pdl_wait();      // prologue done: from here on the kernel reads what the previous kernels of the step wrote
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
  const uint32_t sWQ = smem_addr(smem + AF_WQ), sWKV = smem_addr(smem + AF_WKV), sWO = smem_addr(smem + AF_WO);
  const uint32_t sXU = smem_addr(smem + AF_XU), sXP = smem_addr(smem + AF_XP);
  uint32_t phase = 0;
  const int64_t Np = (A.N + 127) / 128 * 128;
  const int64_t ntiles = (Np + AT_RT - 1) / AT_RT;
  const int gl = row / AT_S, i_q = row - gl * AT_S, gb = gl * AT_S;   // interaction inside the tile, my position, its first row

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t n0 = tile * AT_RT;
    const int64_t n = n0 + row;
    const int64_t avail = max((int64_t)0, min((int64_t)AT_RT, A.N - n0));
    const bool live = row < avail;
    if (tid == 0) {       // pull the next tile's rows into L2 while this one is processed
      const int64_t nn0 = (tile + gridDim.x) * AT_RT;
      const int64_t nav = min((int64_t)AT_RT, A.N - nn0);
      if (nav > 0) {
        bulk_prefetch_l2(A.xu + nn0 * 64, (uint32_t)(nav * 128));
        bulk_prefetch_l2(A.xp + nn0 * 64, (uint32_t)(nav * 128));
      }
    }
    attn_fill_rows<64, false>(smem + AF_XU, A.xu, n0, avail, tid);
    attn_fill_rows<64, false>(smem + AF_XP, A.xp, n0, avail, tid);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_gemm(tmem + 0, sXU, 128, 64 * 16, 256, sWQ, 128, 64 * 16, 256, make_idesc(128, 64, false, false), 4, false);
      issue_gemm(tmem + 64, sXP, 128, 64 * 16, 256, sWKV, 128, 64 * 16, 256, make_idesc(128, 128, false, false), 4, false);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);                     // every warp polls for itself
    phase ^= 1;
    fence_after_sync();
    float qv[16];
    {
      float t[16];
      tmem_ld16(tmem + lane_addr + 64 + h * 16, t);
#pragma unroll
      for (int c = 0; c < 16; ++c) t[c] += bias[64 + h * 16 + c];
      uint8_t* p = smem + AF_KB + (row * AT_XS + h * 16) * 2;
      store_bf16x16(p, p + 16, t);
      tmem_ld16(tmem + lane_addr + 128 + h * 16, t);
#pragma unroll
      for (int c = 0; c < 16; ++c) t[c] += bias[128 + h * 16 + c];
      p = smem + AF_VB + (row * AT_XS + h * 16) * 2;
      store_bf16x16(p, p + 16, t);
      tmem_ld16(tmem + lane_addr + h * 16, qv);
#pragma unroll
      for (int c = 0; c < 16; ++c) qv[c] += bias[h * 16 + c];
    }
    __syncthreads();
    float o[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) o[c] = 0.f;
    if (live) {
      float p[AT_S], mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < AT_S; ++j) {
        float kk[16];
        load_bf16x16(smem + AF_KB + ((gb + j) * AT_XS + h * 16) * 2, kk);
        p[j] = dot16(qv, kk) * 0.25f;          // / sqrt(head_dim = 16)
        mx = fmaxf(mx, p[j]);
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < AT_S; ++j) {
        p[j] = expf(p[j] - mx);
        sum += p[j];
      }
      const float inv = 1.0f / sum;
      const uint64_t gidx = (uint64_t)(tile * (AT_RT / AT_S) + gl);
      const uint32_t keep = attn_keep5(A.rng, ((gidx * HEADS + h) * AT_S + i_q) * AT_S);
      const float ks = A.rng.thresh != 0u ? A.rng.scale : 1.f;
#pragma unroll
      for (int j = 0; j < AT_S; ++j) {
        const float pj = ((keep >> j) & 1u) ? p[j] * inv * ks : 0.f;
        float vv[16];
        load_bf16x16(smem + AF_VB + ((gb + j) * AT_XS + h * 16) * 2, vv);
#pragma unroll
        for (int c = 0; c < 16; ++c) o[c] = fmaf(pj, vv[c], o[c]);
      }
    }
    // ctx tile (bf16 A operand of the output projection) takes the place of the xu tile
    store_bf16x16(smem + AF_XU + tile_off(row, h * 16, 64), smem + AF_XU + tile_off(row, h * 16 + 8, 64), o);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_gemm(tmem + 192, sXU, 128, 64 * 16, 256, sWO, 128, 64 * 16, 256, make_idesc(128, 64, false, false), 4, false);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);                     // every warp polls for itself
    phase ^= 1;
    fence_after_sync();
    {
      float t[16];
      tmem_ld16(tmem + lane_addr + 192 + h * 16, t);
#pragma unroll
      for (int c = 0; c < 16; ++c) t[c] += bias[192 + h * 16 + c];
      if (row < AT_RT && n < Np) {     // rows in [N, Np) hold the bias: finite padding for the MLP tiles
        uint8_t* img = reinterpret_cast<uint8_t*>(A.a_img) + (n >> 7) * (128 * 64 * 2);
        const uint32_t r = (uint32_t)(n & 127);
        store_bf16x16(img + tile_off(r, h * 16, 64), img + tile_off(r, h * 16 + 8, 64), t);
      }
    }
    fence_before_sync();
    __syncthreads();     // TMEM and the tiles are free again
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// =============================================================================================
// backward
// =============================================================================================
constexpr int AT_CXL = 80;                                   // [X | 1 | 0...] tile layout width (bias gradient = ones column)
constexpr uint32_t AB_WQ = 0;
constexpr uint32_t AB_WKV = AB_WQ + 64 * 64 * 2;
constexpr uint32_t AB_WO = AB_WKV + 128 * 64 * 2;
constexpr uint32_t AB_DA = AB_WO + 64 * 64 * 2;              // [128][128-layout] 32 KB  da (cols 64..127 zero)
constexpr uint32_t AB_XU = AB_DA + 128 * 128 * 2;            // [128][80-layout]  20 KB
constexpr uint32_t AB_XP = AB_XU + 128 * AT_CXL * 2;
constexpr uint32_t AB_CTX = AB_XP + 128 * AT_CXL * 2;
constexpr uint32_t AB_EX = AB_CTX + 128 * AT_CXL * 2;        // exchange region, later the dq / dkv operand tiles
constexpr uint32_t AB_KB = AB_EX;
constexpr uint32_t AB_VB = AB_KB + 128 * AT_XS * 2;
constexpr uint32_t AB_QB = AB_VB + 128 * AT_XS * 2;
constexpr uint32_t AB_CB = AB_QB + 128 * AT_XS * 2;          // d ctx rows
constexpr uint32_t AB_PP = AB_CB + 128 * AT_XS * 2;          // p' (after dropout) [128][21] fp32
constexpr uint32_t AB_DS = AB_PP + 128 * AT_PS * 4;          // ds
constexpr uint32_t AB_EX_END = AB_DS + 128 * AT_PS * 4;
constexpr uint32_t AB_DQ = AB_EX;                            // [128][128-layout] 32 KB (cols 64..127 zero)
constexpr uint32_t AB_DKV = AB_DQ + 128 * 128 * 2;           // [128][128]        32 KB
static_assert(AB_DKV + 128 * 128 * 2 <= AB_EX_END, "operand tiles must fit the exchange region");
constexpr uint32_t AB_BIAS = AB_EX_END;
constexpr uint32_t AB_TOTAL = AB_BIAS + 256 * 4;
constexpr int AB_ACC = 3 * AT_CXL * 128;                     // weight-gradient accumulator words per CTA

struct AttnBwdArgs {
  const __nv_bfloat16 *xu, *xp;     // [N,64] bf16 rows
  const __nv_bfloat16* da;          // [N,64] bf16 rows: gradient wrt the block output
  const float* dense;
  float *dxu, *dxp;         // [N,64] out
  float* partial;           // [grid][AB_ACC] per-CTA weight-gradient sums
  int64_t N;
  DropoutRng rng;
};

// Register-staged [128 x 64] bf16 rows: the global loads of the NEXT tile are issued while the tensor core works
// on the current one; the shared-memory stores happen when the tile buffer is free.  Each thread owns 2 chunks of
// 8 columns (same mapping as attn_fill_rows).
struct AttnStage {
  uint4 v[2];
  __device__ __forceinline__ static void rc(int c, int& r, int& j) {
    const int blk = c >> 5, l = c & 31;
    r = (blk >> 1) * 8 + (l & 7);
    j = (blk & 1) * 4 + (l >> 3);
  }
  __device__ __forceinline__ void load(const __nv_bfloat16* __restrict__ src, int64_t row0, int64_t avail, int tid) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      int r, j;
      rc(tid + k * AT_THREADS, r, j);
      v[k] = make_uint4(0, 0, 0, 0);
      if (r < avail) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w) : "l"(src + (row0 + r) * 64 + 8 * j));
      }
    }
  }
  // LAYOUT columns per row group; ONES adds the [1 | 0...] chunk at column 64 (bias gradient)
  template <int LAYOUT, bool ONES>
  __device__ __forceinline__ void store(uint8_t* tile, int64_t avail, int tid) const {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      int r, j;
      rc(tid + k * AT_THREADS, r, j);
      *reinterpret_cast<uint4*>(tile + tile_off(r, 8 * j, LAYOUT)) = v[k];
    }
    if (ONES && tid < 128) {
      uint4 o = make_uint4(0, 0, 0, 0);
      if (tid < avail) o.x = 0x00003f80u;     // bf16(1.0): column 64 = 1, columns 65..71 = 0
      *reinterpret_cast<uint4*>(tile + tile_off(tid, 64, LAYOUT)) = o;
    }
  }
};

__global__ void __launch_bounds__(AT_THREADS, 1) attn_tc_bwd_kernel(AttnBwdArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, h = warp >> 2;
  const int row = q * 32 + lane;
  float* bias = reinterpret_cast<float*>(smem + AB_BIAS);
  float* s_pp = reinterpret_cast<float*>(smem + AB_PP);
  float* s_ds = reinterpret_cast<float*>(smem + AB_DS);
  // zero the operand tiles once: padding columns are never written again
  for (int i = tid; i < (int)(AB_EX - AB_DA) / 16; i += AT_THREADS) reinterpret_cast<uint4*>(smem + AB_DA)[i] = make_uint4(0, 0, 0, 0);
  attn_load_params(smem, AB_WQ, AB_WKV, AB_WO, bias, A.dense, tid);
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
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
  const uint32_t sWQ = smem_addr(smem + AB_WQ), sWKV = smem_addr(smem + AB_WKV), sWO = smem_addr(smem + AB_WO);
  const uint32_t sDA = smem_addr(smem + AB_DA), sXU = smem_addr(smem + AB_XU), sXP = smem_addr(smem + AB_XP);
  const uint32_t sCTX = smem_addr(smem + AB_CTX), sDQ = smem_addr(smem + AB_DQ), sDKV = smem_addr(smem + AB_DKV);
  constexpr uint32_t T_Q = 0, T_KV = 64, T_DC = 192, T_DXU = 0, T_DXP = 64, T_WQ = 256, T_WKV = 256 + AT_CXL, T_WO = 256 + 2 * AT_CXL;
  uint32_t phase = 0;
  bool first = true;
  const int64_t ntiles = (A.N + AT_RT - 1) / AT_RT;
  const int gl = row / AT_S, i_q = row - gl * AT_S, gb = gl * AT_S;

  AttnStage su, sp, sa;
  if ((int64_t)blockIdx.x < ntiles) {
    const int64_t n0 = (int64_t)blockIdx.x * AT_RT, avail = min((int64_t)AT_RT, A.N - n0);
    su.load(A.xu, n0, avail, tid);
    sp.load(A.xp, n0, avail, tid);
    sa.load(A.da, n0, avail, tid);
  }
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t n0 = tile * AT_RT;
    const int64_t n = n0 + row;
    const int64_t avail = max((int64_t)0, min((int64_t)AT_RT, A.N - n0));
    const bool live = row < avail;
    const int64_t nn0 = (tile + gridDim.x) * AT_RT, nav = max((int64_t)0, min((int64_t)AT_RT, A.N - nn0));
    if (tid == 0 && nav > 0) {      // the register prefetch below then finds the next tile in L2
      bulk_prefetch_l2(A.xu + nn0 * 64, (uint32_t)(nav * 128));
      bulk_prefetch_l2(A.xp + nn0 * 64, (uint32_t)(nav * 128));
      bulk_prefetch_l2(A.da + nn0 * 64, (uint32_t)(nav * 128));
    }
    su.store<AT_CXL, true>(smem + AB_XU, avail, tid);
    sp.store<AT_CXL, true>(smem + AB_XP, avail, tid);
    sa.store<128, false>(smem + AB_DA, avail, tid);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      // recompute q, k|v; d ctx = da . Wo
      issue_gemm(tmem + T_Q, sXU, 128, AT_CXL * 16, 256, sWQ, 128, 64 * 16, 256, make_idesc(128, 64, false, false), 4, false);
      issue_gemm(tmem + T_KV, sXP, 128, AT_CXL * 16, 256, sWKV, 128, 64 * 16, 256, make_idesc(128, 128, false, false), 4, false);
      issue_gemm(tmem + T_DC, sDA, 128, 128 * 16, 256, sWO, 64 * 16, 128, 2 * 64 * 16, make_idesc(128, 64, false, true), 4, false);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);                     // every warp polls for itself
    phase ^= 1;
    fence_after_sync();
    float qv[16], dc[16];
    {
      float t[16];
      tmem_ld16(tmem + lane_addr + T_KV + h * 16, t);
#pragma unroll
      for (int c = 0; c < 16; ++c) t[c] += bias[64 + h * 16 + c];
      uint8_t* p = smem + AB_KB + (row * AT_XS + h * 16) * 2;
      store_bf16x16(p, p + 16, t);
      tmem_ld16(tmem + lane_addr + T_KV + 64 + h * 16, t);
#pragma unroll
      for (int c = 0; c < 16; ++c) t[c] += bias[128 + h * 16 + c];
      p = smem + AB_VB + (row * AT_XS + h * 16) * 2;
      store_bf16x16(p, p + 16, t);
      tmem_ld16(tmem + lane_addr + T_Q + h * 16, qv);
#pragma unroll
      for (int c = 0; c < 16; ++c) qv[c] += bias[h * 16 + c];
      p = smem + AB_QB + (row * AT_XS + h * 16) * 2;
      store_bf16x16(p, p + 16, qv);
      tmem_ld16(tmem + lane_addr + T_DC + h * 16, dc);
      p = smem + AB_CB + (row * AT_XS + h * 16) * 2;
      store_bf16x16(p, p + 16, dc);
    }
    __syncthreads();
    // ---- as query: probabilities, ctx, ds, dq ------------------------------------------------------
    float dq[16];
    {
      float o[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        o[c] = 0.f;
        dq[c] = 0.f;
      }
      float pp[AT_S], ds[AT_S];
#pragma unroll
      for (int j = 0; j < AT_S; ++j) {
        pp[j] = 0.f;
        ds[j] = 0.f;
      }
      if (live) {
        float p[AT_S], dp[AT_S], mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < AT_S; ++j) {
          float kk[16];
          load_bf16x16(smem + AB_KB + ((gb + j) * AT_XS + h * 16) * 2, kk);
          p[j] = dot16(qv, kk) * 0.25f;
          mx = fmaxf(mx, p[j]);
          asm volatile("" ::: "memory");      // keep the shared-memory loads of later rows from piling up in registers
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < AT_S; ++j) {
          p[j] = expf(p[j] - mx);
          sum += p[j];
        }
        const float inv = 1.0f / sum;
        const uint64_t gidx = (uint64_t)(tile * (AT_RT / AT_S) + gl);
        const uint32_t keep = attn_keep5(A.rng, ((gidx * HEADS + h) * AT_S + i_q) * AT_S);
        const float ks = A.rng.thresh != 0u ? A.rng.scale : 1.f;
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < AT_S; ++j) {
          p[j] *= inv;
          const float kj = ((keep >> j) & 1u) ? ks : 0.f;
          pp[j] = p[j] * kj;
          float vv[16];
          load_bf16x16(smem + AB_VB + ((gb + j) * AT_XS + h * 16) * 2, vv);
#pragma unroll
          for (int c = 0; c < 16; ++c) o[c] = fmaf(pp[j], vv[c], o[c]);
          dp[j] = dot16(dc, vv) * kj;                      // dL/dp_ij
          dot = fmaf(p[j], dp[j], dot);
          asm volatile("" ::: "memory");
        }
#pragma unroll
        for (int j = 0; j < AT_S; ++j) {
          ds[j] = p[j] * (dp[j] - dot) * 0.25f;            // dL/ds_ij (scaled by 1/sqrt(16))
          float kk[16];
          load_bf16x16(smem + AB_KB + ((gb + j) * AT_XS + h * 16) * 2, kk);
#pragma unroll
          for (int c = 0; c < 16; ++c) dq[c] = fmaf(ds[j], kk[c], dq[c]);
          asm volatile("" ::: "memory");
        }
      }
#pragma unroll
      for (int j = 0; j < AT_S; ++j) {
        s_pp[row * AT_PS + h * AT_S + j] = pp[j];
        s_ds[row * AT_PS + h * AT_S + j] = ds[j];
      }
      store_bf16x16(smem + AB_CTX + tile_off(row, h * 16, AT_CXL), smem + AB_CTX + tile_off(row, h * 16 + 8, AT_CXL), o);
    }
    if (h == 0) {      // ones column of the ctx tile (bias gradient of out_proj)
      uint4 v = make_uint4(0, 0, 0, 0);
      if (live) v.x = 0x00003f80u;
      *reinterpret_cast<uint4*>(smem + AB_CTX + tile_off(row, 64, AT_CXL)) = v;
    }
    __syncthreads();
    // ---- as key / value: dk_j = sum_i ds_ij q_i, dv_j = sum_i p'_ij dctx_i --------------------------
    float dk[16], dv[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      dk[c] = 0.f;
      dv[c] = 0.f;
    }
    if (live) {
#pragma unroll
      for (int i = 0; i < AT_S; ++i) {
        const float dsi = s_ds[(gb + i) * AT_PS + h * AT_S + i_q], ppi = s_pp[(gb + i) * AT_PS + h * AT_S + i_q];
        float t[16];
        load_bf16x16(smem + AB_QB + ((gb + i) * AT_XS + h * 16) * 2, t);
#pragma unroll
        for (int c = 0; c < 16; ++c) dk[c] = fmaf(dsi, t[c], dk[c]);
        load_bf16x16(smem + AB_CB + ((gb + i) * AT_XS + h * 16) * 2, t);
#pragma unroll
        for (int c = 0; c < 16; ++c) dv[c] = fmaf(ppi, t[c], dv[c]);
        asm volatile("" ::: "memory");
      }
    }
    __syncthreads();     // the exchange arrays are dead: the operand tiles take their place
    {
      float z[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) z[c] = 0.f;
      store_bf16x16(smem + AB_DQ + tile_off(row, h * 16, 128), smem + AB_DQ + tile_off(row, h * 16 + 8, 128), dq);
      store_bf16x16(smem + AB_DQ + tile_off(row, 64 + h * 16, 128), smem + AB_DQ + tile_off(row, 64 + h * 16 + 8, 128), z);
      store_bf16x16(smem + AB_DKV + tile_off(row, h * 16, 128), smem + AB_DKV + tile_off(row, h * 16 + 8, 128), dk);
      store_bf16x16(smem + AB_DKV + tile_off(row, 64 + h * 16, 128), smem + AB_DKV + tile_off(row, 64 + h * 16 + 8, 128), dv);
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      // dxu = dq . Wq ; dxp = [dk|dv] . [Wk;Wv]          (A K-major on the 128-wide layout, B MN-major)
      issue_gemm(tmem + T_DXU, sDQ, 128, 128 * 16, 256, sWQ, 64 * 16, 128, 2 * 64 * 16, make_idesc(128, 64, false, true), 4, false);
      issue_gemm(tmem + T_DXP, sDKV, 128, 128 * 16, 256, sWKV, 64 * 16, 128, 2 * 64 * 16, make_idesc(128, 64, false, true), 8, false);
      // dW (+ bias column) += dY^T . [X | 1], accumulated over all tiles of this CTA
      issue_gemm(tmem + T_WQ, sDQ, 128 * 16, 128, 2 * 128 * 16, sXU, AT_CXL * 16, 128, 2 * AT_CXL * 16,
                 make_idesc(128, AT_CXL, true, true), 8, !first);
      issue_gemm(tmem + T_WKV, sDKV, 128 * 16, 128, 2 * 128 * 16, sXP, AT_CXL * 16, 128, 2 * AT_CXL * 16,
                 make_idesc(128, AT_CXL, true, true), 8, !first);
      issue_gemm(tmem + T_WO, sDA, 128 * 16, 128, 2 * 128 * 16, sCTX, AT_CXL * 16, 128, 2 * AT_CXL * 16,
                 make_idesc(128, AT_CXL, true, true), 8, !first);
      mma_commit(&bar);
    }
    first = false;
    if (nav > 0) {       // next tile's rows: in flight during the GEMMs and the dx stores
      su.load(A.xu, nn0, nav, tid);
      sp.load(A.xp, nn0, nav, tid);
      sa.load(A.da, nn0, nav, tid);
    }
    mbar_wait(&bar, phase);                     // every warp polls for itself
    phase ^= 1;
    fence_after_sync();
    {
      float t[16];
      tmem_ld16(tmem + lane_addr + T_DXU + h * 16, t);
      if (live) {
#pragma unroll
        for (int j = 0; j < 2; ++j)      // bf16 rows: K6 sums them per id in fp32
          *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(A.dxu) + n * 64 + h * 16 + 8 * j) =
              make_uint4(pack_bf16(t[8 * j], t[8 * j + 1]), pack_bf16(t[8 * j + 2], t[8 * j + 3]), pack_bf16(t[8 * j + 4], t[8 * j + 5]),
                         pack_bf16(t[8 * j + 6], t[8 * j + 7]));
      }
      tmem_ld16(tmem + lane_addr + T_DXP + h * 16, t);
      if (live) {
#pragma unroll
        for (int j = 0; j < 2; ++j)      // bf16 rows: K6 sums them per id in fp32
          *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(A.dxp) + n * 64 + h * 16 + 8 * j) =
              make_uint4(pack_bf16(t[8 * j], t[8 * j + 1]), pack_bf16(t[8 * j + 2], t[8 * j + 3]), pack_bf16(t[8 * j + 4], t[8 * j + 5]),
                         pack_bf16(t[8 * j + 6], t[8 * j + 7]));
      }
    }
    fence_before_sync();
    __syncthreads();
  }
  // flush: this CTA's accumulators go to its slice of the partial buffer (attn_wgrad_reduce_kernel adds them up);
  // CTAs without tiles store zeros.  Word index = column * 128 + lane.
  {
    float* part = A.partial + (int64_t)blockIdx.x * AB_ACC;
    for (int ch = 0; ch < 3 * AT_CXL / 16 / 4 + 1; ++ch) {     // 15 chunks of 16 columns over the 4 column groups h
      const int c0 = (ch * 4 + h) * 16;
      if (c0 >= 3 * AT_CXL) break;
      float v[16];
      if (!first) tmem_ld16(tmem + lane_addr + T_WQ + c0, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) part[(int64_t)(c0 + i) * 128 + row] = first ? 0.f : v[i];
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// dense_grad += sum over CTAs of the partial accumulators.  Columns [0,80) = dWq (lane = out feature, 64 =
// bias), [80,160) = d[Wk;Wv], [160,240) = dWo.
__global__ void __launch_bounds__(256) attn_wgrad_reduce_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ dg) {
  pdl_launch_dependents();
  pdl_wait();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= AB_ACC) return;
  const int col = e >> 7, lane = e & 127;
  const int which = col / AT_CXL, c = col % AT_CXL;
  if (c > 64 || (which != 1 && lane >= 64)) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += partial[(int64_t)p * AB_ACC + e];
  int64_t off;
  if (c < 64) off = (which == 0 ? NCF_OFF(NCF_P_Q_W) : which == 1 ? NCF_OFF(NCF_P_K_W) : NCF_OFF(NCF_P_O_W)) + (int64_t)lane * 64 + c;
  else off = (which == 0 ? NCF_OFF(NCF_P_Q_B) : which == 1 ? NCF_OFF(NCF_P_K_B) : NCF_OFF(NCF_P_O_B)) + lane;
  dg[off] += s;
}

int64_t attn_tc_partial_floats() { return (int64_t)num_sms() * AB_ACC; }

int attn_tc_forward(const ncf_run_cfg& cfg, const float* dense, int64_t N, TowerWs& w, cudaStream_t st) {
  if (N == 0) return NCF_OK;
  static bool configured = false;
  if (!configured) {
    NCF_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AF_TOTAL));
    configured = true;
  }
  AttnFwdArgs A{};
  A.xu = reinterpret_cast<const __nv_bfloat16*>(w.xu);      // bf16 rows (tower_bf16_rows)
  A.xp = reinterpret_cast<const __nv_bfloat16*>(w.xp);
  A.dense = dense;
  A.a_img = w.a_img;
  A.N = N;
  A.rng = make_rng(cfg, 0);
  const int grid = even_grid(attn_tc_tiles(N), (int64_t)tower_sms() * 2);
  NCF_CUDA(launch_pdl(PDL_ATTN_FWD, attn_tc_fwd_kernel, dim3(grid), dim3(AT_THREADS), AF_TOTAL, st, A));
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

// da = w.g64a -> dxu = w.g64b, dxp = w.g256 ([N,64]); accumulates the six attention parameter gradients
int attn_tc_backward(const ncf_run_cfg& cfg, const float* dense, float* dense_grad, int64_t N, TowerWs& w, cudaStream_t st,
                     int leave_sms, cudaStream_t reduce_st, cudaEvent_t done) {
  if (N == 0) return NCF_OK;
  static bool configured = false;
  if (!configured) {
    NCF_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AB_TOTAL));
    configured = true;
  }
  AttnBwdArgs A{};
  A.xu = reinterpret_cast<const __nv_bfloat16*>(w.xu);      // bf16 rows (tower_bf16_rows)
  A.xp = reinterpret_cast<const __nv_bfloat16*>(w.xp);
  A.da = reinterpret_cast<const __nv_bfloat16*>(w.g64a);
  A.dense = dense;
  A.dxu = w.g64b;
  A.dxp = w.g256;
  A.partial = w.at_partial;
  A.N = N;
  A.rng = make_rng(cfg, 0);
  const int grid = even_grid((N + AT_RT - 1) / AT_RT, std::max(1, tower_sms() - leave_sms));
  NCF_CUDA(launch_pdl(PDL_ATTN_BWD, attn_tc_bwd_kernel, dim3(grid), dim3(AT_THREADS), AB_TOTAL, st, A));
  NCF_LAUNCH_CHECK();
  cudaStream_t rst = st;
  if (reduce_st && done) {      // nothing before the dense Adam reads these gradients: off the stream the embedding backward waits on
    NCF_CUDA(cudaEventRecord(done, st));
    NCF_CUDA(cudaStreamWaitEvent(reduce_st, done, 0));
    rst = reduce_st;
  }
  NCF_CUDA(launch_pdl(PDL_ATTN_BWD, attn_wgrad_reduce_kernel, dim3((AB_ACC + 255) / 256), dim3(256), 0, rst, (const float*)w.at_partial, grid, dense_grad));
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

}  // namespace ncf
