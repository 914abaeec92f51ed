// Full-catalogue scoring + top-k with a tensor-core PRE-FILTER (app.py:43-77 at the config[4] scale).
//
// The exact kernel (ncf_score.cu) forms every logit LN(U_mf[u]) . P_hat[i] + g[i] with 64 fp32 FMAs; almost all
// of them lose against the user's running k-th best.  Here a tcgen05 bf16 GEMM (128 users x 256 items per tile,
// accumulators double-buffered in TMEM) produces APPROXIMATE logits with a rigorous error bound
//     |z_bf16 - z| <= 1.01 * 2^-7 * ||u||_2 * ||p_i||_2
// (bf16 keeps 8 significant bits: each operand is off by at most 2^-8 relative, a product by 2^-7 (1 + 2^-9); the
// products are exact in fp32; the fp32 accumulation of 64 terms - on the tensor core and in the exact kernel's FMA
// chain - adds less than 1e-5 relative; Cauchy-Schwarz turns sum |u_k p_k| into the two norms; 1 % covers the rest)
// and only pairs whose upper bound reaches the running threshold are re-scored EXACTLY - same fp32 FMA order, same
// sigmoid, same (score desc, index asc) key as the exact kernel - so the result is bit-identical to it while the
// CUDA cores touch a tiny fraction of the 10^13 pairs.  Item tiles arrive as ready operand images
// (ncf_item_image: bf16 canonical layout + g + 2^-7 ||p_i||) by one TMA bulk copy each.
#include "ncf_tower.cuh"
#include "ncf_umma.cuh"

namespace ncf {
using namespace umma;

constexpr int SC_THREADS = 512;            // 16 warps: TMEM lane quarter x accumulator column quarter
constexpr int SC_UT = 128;                 // users per CTA (M)
constexpr int SC_IT = 256;                 // items per tile (N)
constexpr int SC_KMAX = 128;               // running list length (k <= 128), same as the exact kernel
constexpr int SC_CAP = 512;                // list + candidates per user
constexpr int SC_K = 80;                   // GEMM depth: 64 embedding columns + [g_hi, g_lo, margin, slack] + 12 zero columns
constexpr uint32_t SC_IMG = SC_IT * SC_K * 2;               // bf16 operand image of one item tile
constexpr uint32_t SC_TILE_BYTES = SC_IMG;
constexpr float SC_EPS = 0.0078125f * 1.01f;   // 1.01 * 2^-7, see the bound above

constexpr uint32_t SCS_A = 0;                               // [128][80] bf16 image            20 KB
constexpr uint32_t SCS_U = SCS_A + SC_UT * SC_K * 2;        // [128][64] fp32 LN'd user rows   32 KB
constexpr int SC_NB = 3;                                    // item tile buffers: a load has a whole iteration to land
constexpr uint32_t SCS_B = SCS_U + SC_UT * 64 * 4;          // 3 x item tile                   102 KB
constexpr uint32_t SCS_SORT = SCS_B + SC_NB * SC_TILE_BYTES;    // 512 keys being merged         4 KB
constexpr uint32_t SCS_NU = SCS_SORT + SC_CAP * 8;          // ||u|| per user
constexpr uint32_t SCS_LTHR = SCS_NU + SC_UT * 4;           // logit pre-filter per user
constexpr uint32_t SCS_KTHR = SCS_LTHR + SC_UT * 4;         // key of the running k-th best
constexpr uint32_t SCS_CNT = SCS_KTHR + SC_UT * 8;
constexpr int SC_QCAP = 4096;                               // survivors queued for exact re-scoring (flushed when half full)
constexpr int SC_HIGH = SC_CAP - 64;                        // merge a list when it holds more than this
constexpr int SC_RCAP = SC_UT * SC_IT + SC_QCAP;            // keys whose candidate buffer was full: one tile + the queue
constexpr uint32_t SCS_QUEUE = SCS_CNT + SC_UT * 4;         // (row << 8 | column) per survivor
constexpr uint32_t SCS_TOTAL = SCS_QUEUE + SC_QCAP * 8;     // (row, item index) per survivor
constexpr int64_t SC_CTA_WS = (int64_t)SC_UT * SC_CAP * 8 + 2 * ((int64_t)SC_RCAP * 8 + SC_RCAP);   // candidates + 2 retry lists

__device__ __forceinline__ unsigned long long sc_key(float score, uint32_t idx) {
  return ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(0xffffffffu - idx);
}

// ---- item tile images ----------------------------------------------------------------------------------
// The per-item terms of the bound ride in the GEMM itself (columns 64..67 of the operand images):
//     bound(u, i) = sum_k bf16(u_k) bf16(p_ik)  +  1 * g_hi  +  1 * g_lo  +  up(||u||) * up(margin_i)  +  1 * up(slack_i)
// g = g_hi + g_lo up to 2^-18 |g| (two bf16 pieces); margin_i = 1.01 * 2^-7 ||p_i|| as before; slack_i = 2^-15 |g_i| + the
// split's remainder covers the fp32 accumulation of the extra terms; up() rounds to the next bf16 ABOVE, so every
// replacement errs upwards and the accumulator is an upper bound of the exact logit by construction.  The scan of a pair
// is then one subtraction (bound - threshold) and one funnel shift: no per-item loads in the epilogue at all.
__device__ __forceinline__ uint16_t bf16_bits_up(float x) {          // x >= 0: smallest bf16 >= x
  __nv_bfloat16 b = __float2bfloat16_rn(x);
  uint16_t bits = *reinterpret_cast<uint16_t*>(&b);
  if (__bfloat162float(b) < x) ++bits;
  return bits;
}
__device__ __forceinline__ uint16_t bf16_bits_rn(float x) {
  __nv_bfloat16 b = __float2bfloat16_rn(x);
  return *reinterpret_cast<uint16_t*>(&b);
}
// thread = (item, 8-column chunk); the 8 chunk threads of an item are adjacent lanes
__global__ void __launch_bounds__(256) item_image_kernel(const float* __restrict__ p_hat, const float* __restrict__ g, int64_t I,
                                                         uint8_t* __restrict__ img) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = t >> 3;
  const int j = (int)(t & 7);
  const int64_t ntiles = (I + SC_IT - 1) / SC_IT;
  if (i >= ntiles * SC_IT) return;
  float4 a = make_float4(0, 0, 0, 0), b = a;
  if (i < I) {
    a = ldg4(p_hat + i * 64 + 8 * j);
    b = ldg4(p_hat + i * 64 + 8 * j + 4);
  }
  float ss = f4_dot(a, a) + f4_dot(b, b);
  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
  ss += __shfl_xor_sync(0xffffffffu, ss, 4);
  uint8_t* tile = img + (i / SC_IT) * SC_TILE_BYTES;
  const uint32_t r = (uint32_t)(i % SC_IT);
  *reinterpret_cast<uint4*>(tile + tile_off(r, 8 * j, SC_K)) =
      make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  if (j == 0) {
    uint32_t w0, w1;
    if (i < I) {
      const float gi = __ldg(g + i);
      const uint16_t hi = bf16_bits_rn(gi);
      const float rem = gi - __uint_as_float((uint32_t)hi << 16);                 // exact in fp32
      const uint16_t lo = bf16_bits_rn(rem);
      const float rem2 = fabsf(rem - __uint_as_float((uint32_t)lo << 16));
      const uint16_t mg = bf16_bits_up(SC_EPS * sqrtf(ss) * 1.0001f);
      const uint16_t sl = bf16_bits_up(fabsf(gi) * 3.0517578125e-5f + rem2 + 1e-30f);
      w0 = (uint32_t)hi | ((uint32_t)lo << 16);
      w1 = (uint32_t)mg | ((uint32_t)sl << 16);
    } else {            // padding items can never pass the filter: a huge negative bias, no NaN
      w0 = 0xff00u;     // bf16(-1.7e38)
      w1 = 0u;
    }
    *reinterpret_cast<uint4*>(tile + tile_off(r, 64, SC_K)) = make_uint4(w0, w1, 0u, 0u);
    *reinterpret_cast<uint4*>(tile + tile_off(r, 72, SC_K)) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// descending bitonic sort of SC_CAP keys in shared memory by the whole CTA (one compare-exchange per thread and stage)
__device__ __forceinline__ void cta_sort_desc(unsigned long long* a, int tid) {
  for (int k = 2; k <= SC_CAP; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      if (tid < SC_CAP / 2) {
        const int i = ((tid & ~(j - 1)) << 1) | (tid & (j - 1));     // the tid-th index whose bit j is clear
        const int l = i | j;
        const unsigned long long x = a[i], y = a[l];
        const bool desc = (i & k) == 0;
        if (desc ? x < y : x > y) {
          a[i] = y;
          a[l] = x;
        }
      }
    }
  }
  __syncthreads();
}

// ---- warp-level descending sort of 512 keys (16 per lane; element e = lane * 16 + r) ---------------------------------
// A candidate-list merge used to be one 512-key bitonic network run by the WHOLE CTA (45 stages, a CTA barrier each, one
// user at a time: a quarter of the kernel's time).  Here a warp sorts a user's list on its own - register-local stages for
// partner distances below 16, shuffles above - so the 16 warps merge 16 users at once and no barrier is involved.
template <int J>
__device__ __forceinline__ void sort_local_stage(unsigned long long (&v)[16], int lane, int k) {
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    if ((r & J) == 0) {
      const bool desc = (((lane << 4) | r) & k) == 0;
      const unsigned long long a = v[r], b = v[r | J];
      const unsigned long long hi = a > b ? a : b, lo = a > b ? b : a;
      v[r] = desc ? hi : lo;
      v[r | J] = desc ? lo : hi;
    }
  }
}
__device__ __forceinline__ void warp_sort512_desc(unsigned long long (&v)[16], int lane) {
#pragma unroll 1
  for (int k = 2; k <= 512; k <<= 1) {
#pragma unroll 1
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 16) {
        const int lj = j >> 4;                                   // partner lane distance
        const bool lower = (lane & lj) == 0;                     // this lane holds the lower index of every pair
        const bool desc = ((lane << 4) & k) == 0;                // k >= 32 here: the block direction depends on the lane only
        const bool keep_max = lower == desc;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const unsigned long long o = __shfl_xor_sync(0xffffffffu, v[r], lj);
          v[r] = keep_max ? (v[r] > o ? v[r] : o) : (v[r] > o ? o : v[r]);
        }
      } else if (j == 8) sort_local_stage<8>(v, lane, k);
      else if (j == 4) sort_local_stage<4>(v, lane, k);
      else if (j == 2) sort_local_stage<2>(v, lane, k);
      else sort_local_stage<1>(v, lane, k);
    }
  }
}

struct ScoreTcArgs {
  const float* t_umf;
  const float* dense;
  const float* p_hat;      // [I,64] fp32 (exact re-scoring)
  const float* g;          // [I]
  const uint8_t* img;      // item tile images
  const int64_t* user_ids;
  int64_t n_users, I;
  int64_t rows_user;
  int32_t* status;
  int nsplit;
  unsigned long long* cand;   // [gridDim.x * gridDim.y][128][512]
  unsigned long long* part;   // [n_users][nsplit][128]
};

__global__ void __launch_bounds__(SC_THREADS, 1) score_tc_kernel(ScoreTcArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[SC_NB], accb[2];
  __shared__ uint32_t tmem_slot;
  __shared__ int s_qn, s_rn;
  __shared__ uint32_t s_mmask[SC_UT / 32];               // users whose list has to be merged
  unsigned long long* s_queue = reinterpret_cast<unsigned long long*>(smem + SCS_QUEUE);

  float* s_u = reinterpret_cast<float*>(smem + SCS_U);
  float* s_nu = reinterpret_cast<float*>(smem + SCS_NU);
  float* s_lthr = reinterpret_cast<float*>(smem + SCS_LTHR);
  unsigned long long* s_kthr = reinterpret_cast<unsigned long long*>(smem + SCS_KTHR);
  int* s_cnt = reinterpret_cast<int*>(smem + SCS_CNT);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = warp & 3, cq = warp >> 2;
  const int row = q * 32 + lane;
  const int64_t u0 = (int64_t)blockIdx.x * SC_UT;
  const int split = blockIdx.y;
  const int nu = (int)min((int64_t)SC_UT, A.n_users - u0);
  // this CTA's slice of the workspace: candidate buffers, then two retry lists (keys, rows) used alternately
  uint8_t* cta_ws = reinterpret_cast<uint8_t*>(A.cand) + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * SC_CTA_WS;
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(cta_ws);
  unsigned long long* retry_key[2];
  uint8_t* retry_row[2];
  retry_key[0] = cand + (int64_t)SC_UT * SC_CAP;
  retry_key[1] = retry_key[0] + SC_RCAP;
  retry_row[0] = reinterpret_cast<uint8_t*>(retry_key[1] + SC_RCAP);
  retry_row[1] = retry_row[0] + SC_RCAP;
  int rsel = 0;                                          // list that receives the overflowing pairs (uniform)

  // LN_mf of the tile's user rows: one warp per user, two-pass variance like nn.LayerNorm (same arithmetic as the
  // exact kernel); fp32 row for the exact re-scoring, bf16 operand image, ||u||
  for (int uu = warp; uu < SC_UT; uu += SC_THREADS / 32) {
    float x0 = 0.f, x1 = 0.f;
    if (uu < nu) {
      const int64_t uid = A.user_ids ? A.user_ids[u0 + uu] : u0 + uu;
      if (bad_id(uid, A.rows_user)) flag_status(A.status, NCF_STATUS_BAD_USER_ID);
      const float* r = A.t_umf + clamp_id(uid, A.rows_user) * D;
      x0 = r[lane];
      x1 = r[lane + 32];
    }
    float y0 = x0, y1 = x1;                  // A.dense == NULL: raw query rows (ncf_dot_topk), no LayerNorm
    if (A.dense) {
      const float mean = warp_sum(x0 + x1) * (1.0f / D);
      const float d0 = x0 - mean, d1 = x1 - mean;
      const float rstd = rsqrtf(warp_sum(d0 * d0 + d1 * d1) * (1.0f / D) + LN_EPS);
      y0 = fmaf(d0 * rstd, __ldg(A.dense + NCF_OFF(NCF_P_MF_NORM_W) + lane), __ldg(A.dense + NCF_OFF(NCF_P_MF_NORM_B) + lane));
      y1 = fmaf(d1 * rstd, __ldg(A.dense + NCF_OFF(NCF_P_MF_NORM_W) + lane + 32), __ldg(A.dense + NCF_OFF(NCF_P_MF_NORM_B) + lane + 32));
    }
    if (uu >= nu) {
      y0 = 0.f;
      y1 = 0.f;
    }
    s_u[uu * 64 + lane] = y0;
    s_u[uu * 64 + lane + 32] = y1;
    reinterpret_cast<__nv_bfloat16*>(smem + SCS_A + tile_off(uu, lane, SC_K))[0] = __float2bfloat16_rn(y0);
    reinterpret_cast<__nv_bfloat16*>(smem + SCS_A + tile_off(uu, lane + 32, SC_K))[0] = __float2bfloat16_rn(y1);
    const float nn = sqrtf(warp_sum(y0 * y0 + y1 * y1)) * 1.0001f;
    if (lane < 16) {      // columns 64..79: [1, 1, up(||u||), 1, 0 ...] = the coefficients of g_hi, g_lo, margin, slack
      uint16_t c = 0;
      if (uu < nu) c = (lane == 0 || lane == 1 || lane == 3) ? (uint16_t)0x3f80 : (lane == 2 ? bf16_bits_up(nn) : (uint16_t)0);
      reinterpret_cast<uint16_t*>(smem + SCS_A + tile_off(uu, 64 + lane, SC_K))[0] = c;
    }
    if (lane == 0) {
      s_nu[uu] = nn;
      s_lthr[uu] = uu < nu ? -INFINITY : INFINITY;      // rows without a user never pass
      s_kthr[uu] = 0ull;
      s_cnt[uu] = SC_KMAX;                              // slots [0,KMAX) = the running list (zeros = empty)
    }
  }
  for (int i = tid; i < SC_UT * SC_KMAX; i += SC_THREADS) cand[(i / SC_KMAX) * SC_CAP + (i % SC_KMAX)] = 0ull;
  if (tid == 0) {
    s_qn = 0;
    s_rn = 0;
    for (int i = 0; i < SC_UT / 32; ++i) s_mmask[i] = 0u;
    for (int i = 0; i < SC_NB; ++i) mbar_init(&full[i], 1);
    mbar_init(&accb[0], 1);
    mbar_init(&accb[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
  const uint32_t sA = smem_addr(smem + SCS_A);

  const int64_t ntile_all = (A.I + SC_IT - 1) / SC_IT;
  const int64_t per = (ntile_all + A.nsplit - 1) / A.nsplit;
  const int64_t t_begin = split * per, t_end = min(ntile_all, t_begin + per);
  const int64_t ntile = t_begin < t_end ? t_end - t_begin : 0;

  auto load_tile = [&](int64_t t) {                    // tid 0: tile t of this split into buffer t % SC_NB
    const int nb = (int)(t % SC_NB);
    mbar_arrive_expect_tx(&full[nb], SC_TILE_BYTES);
    bulk_g2s(smem + SCS_B + nb * SC_TILE_BYTES, A.img + (t_begin + t) * SC_TILE_BYTES, SC_TILE_BYTES, &full[nb]);
  };
  auto issue_mma = [&](int64_t t) {                    // tid 0: tile t -> accumulator t & 1
    const int nb = (int)(t % SC_NB), b = (int)(t & 1);
    mbar_wait(&full[nb], (uint32_t)((t / SC_NB) & 1));
    fence_after_sync();
    issue_gemm(tmem + 256 * b, sA, 128, SC_K * 16, 256, smem_addr(smem + SCS_B + nb * SC_TILE_BYTES), 128, SC_K * 16, 256,
               make_idesc(128, SC_IT, false, false), SC_K / 16, false);
    mma_commit(&accb[b]);
  };
  // merge the candidate buffer of one user into its running list: ONE WARP sorts the (at most 512) keys in registers,
  // keeps the best KMAX and refreshes the thresholds (dst != null: also emits the list - the final pass).
  auto warp_merge = [&](int uu, unsigned long long* dst) {
    const int cnt = min(s_cnt[uu], SC_CAP);
    unsigned long long* c = cand + (int64_t)uu * SC_CAP;
    unsigned long long v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int e = lane * 16 + r;
      v[r] = e < cnt ? __ldcg(c + e) : 0ull;
    }
    warp_sort512_desc(v, lane);
    if (lane < SC_KMAX / 16) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        c[lane * 16 + r] = v[r];
        if (dst) dst[lane * 16 + r] = v[r];
      }
    }
    const unsigned long long kth = __shfl_sync(0xffffffffu, v[15], SC_KMAX / 16 - 1);
    if (lane == 0) {
      s_cnt[uu] = SC_KMAX;
      s_kthr[uu] = kth;
      const float sk = __uint_as_float((uint32_t)(kth >> 32));
      float lt = -INFINITY;
      if (kth != 0ull && sk > 0.f && sk < 1.f) {          // same pre-filter as the exact kernel
        lt = logf(sk / (1.0f - sk));
        lt -= 1e-4f + 1e-4f * fabsf(lt);
      }
      s_lthr[uu] = lt;
    }
  };
  auto mark_merge = [&](int r) { atomicOr(&s_mmask[r >> 5], 1u << (r & 31)); };
  auto run_merges = [&]() {            // all threads; the marked users are dealt round-robin to the 16 warps
    __syncthreads();
    uint32_t m[SC_UT / 32];
#pragma unroll
    for (int wd = 0; wd < SC_UT / 32; ++wd) m[wd] = s_mmask[wd];
    __syncthreads();
    if (tid < SC_UT / 32) s_mmask[tid] = 0u;
    int nth = 0;
#pragma unroll
    for (int wd = 0; wd < SC_UT / 32; ++wd) {
      uint32_t mm = m[wd];
      while (mm) {
        const int uu = wd * 32 + __ffs(mm) - 1;
        mm &= mm - 1;
        if ((nth++ & (SC_THREADS / 32 - 1)) == warp) warp_merge(uu, nullptr);
      }
    }
    __syncthreads();
  };

  bool want_flush = false;      // this thread saw the queue pass its half mark or parked a key on the retry list
  // exact logit of (user row r, item i) with the exact kernel's arithmetic; pushes the key if it can enter.  A full
  // candidate buffer (possible while the thresholds are still loose) parks the key on the retry list, which is drained
  // after the merges.
  auto push_key = [&](int r, unsigned long long key) {
    const int pos = atomicAdd(&s_cnt[r], 1);
    if (pos < SC_CAP) {
      cand[(int64_t)r * SC_CAP + pos] = key;
      if (pos >= SC_HIGH) mark_merge(r);
    } else {
      atomicSub(&s_cnt[r], 1);
      mark_merge(r);
      want_flush = true;
      const int rp = atomicAdd(&s_rn, 1);     // < SC_RCAP per round (see the flush policy)
      retry_key[rsel][rp] = key;
      retry_row[rsel][rp] = (uint8_t)r;
    }
  };
  auto rescore = [&](int r, int64_t i) {
    const float* pr = A.p_hat + i * 64;
    const float* ur = s_u + r * 64;
    float4 p[16];
#pragma unroll
    for (int k4 = 0; k4 < 16; ++k4) p[k4] = ldg4(pr + 4 * k4);
    const float gi = __ldg(A.g + i);
    float acc = 0.f;
#pragma unroll
    for (int k4 = 0; k4 < 16; ++k4) {            // ascending k with fmaf: the exact kernel's order
      const float4 uv = *reinterpret_cast<const float4*>(ur + 4 * k4);
      acc = fmaf(uv.x, p[k4].x, acc);
      acc = fmaf(uv.y, p[k4].y, acc);
      acc = fmaf(uv.z, p[k4].z, acc);
      acc = fmaf(uv.w, p[k4].w, acc);
    }
    const float ze = acc + gi;
    if (ze >= s_lthr[r]) {
      const float sc = 1.0f / (1.0f + expf(-ze));
      const unsigned long long key = sc_key(sc, (uint32_t)i);
      if (key > s_kthr[r]) push_key(r, key);
    }
  };
  // drain the survivor queue (every thread takes one pair), merge the lists that filled up, retry the keys that found
  // their buffer full - until nothing is left
  auto flush = [&]() {
    __syncthreads();
    const int qn = min(s_qn, SC_QCAP);
    for (int e = tid; e < qn; e += SC_THREADS) {
      const unsigned long long v = s_queue[e];
      rescore((int)(v >> 32), (int64_t)(v & 0xffffffffull));
    }
    __syncthreads();
    if (tid == 0) s_qn = 0;
    for (;;) {
      run_merges();
      const int rn = s_rn;
      __syncthreads();                 // everybody has read rn before anyone moves on (and may park new keys)
      if (rn == 0) break;
      if (tid == 0) s_rn = 0;
      const int src = rsel;
      rsel ^= 1;
      __syncthreads();
      for (int e = tid; e < rn; e += SC_THREADS) {
        const int r = retry_row[src][e];
        const unsigned long long key = retry_key[src][e];
        if (key > s_kthr[r]) push_key(r, key);
      }
    }
  };

  if (tid == 0 && ntile > 0) {
    load_tile(0);
    if (ntile > 1) load_tile(1);
    issue_mma(0);
  }
  for (int64_t t = 0; t < ntile; ++t) {
    const int b = (int)(t & 1);
    if (tid == 0) {
      if (t + 2 < ntile) load_tile(t + 2);              // its buffer held tile t - 1: GEMM and epilogue are done
      if (t + 1 < ntile) issue_mma(t + 1);              // the other accumulator was drained by the previous epilogue
    }
    const float lthr = s_lthr[row];                      // changed by flushes only (CTA barriers on both sides)
    mbar_wait(&accb[b], (uint32_t)((t >> 1) & 1));       // every warp polls for itself: one CTA barrier per tile, below
    fence_after_sync();
    const int64_t base = (t_begin + t) * SC_IT;
    // both halves of this thread's 64 accumulator columns are requested before either is used
    float z[2][32];
    tmem_ld32_nowait(tmem + 256 * b + lane_addr + cq * 64, z[0]);
    tmem_ld32_nowait(tmem + 256 * b + lane_addr + cq * 64 + 32, z[1]);
    tmem_ld_wait();
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const int j0 = cq * 64 + ch * 32;
      // two instructions per pair: the accumulator IS the upper bound of the logit (bias, margin and slack ride in the
      // GEMM), so diff = bound - lthr and a funnel shift that appends diff's sign bit to a mask; four independent chains
      // of eight (a single 32-long chain is pure latency), joined so that element j ends up at bit 31 - j (set = miss)
      uint32_t m4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int c = 0; c < 4; ++c) m4[c] = __funnelshift_l(__float_as_uint(z[ch][8 * c + j] - lthr), m4[c], 1);
      }
      const uint32_t miss = (m4[0] << 24) | (m4[1] << 16) | (m4[2] << 8) | m4[3];
      uint32_t hit = ~miss;
      while (hit) {                                      // rare after the first tiles
        const int j = __clz(hit);                        // element j sits at bit 31 - j
        hit &= ~(0x80000000u >> j);
        const int64_t i = base + j0 + j;
        if (i >= A.I) continue;                          // padding of the last tile
        const int qp = atomicAdd(&s_qn, 1);
        if (qp >= SC_QCAP / 2) want_flush = true;
        if (qp < SC_QCAP) {      // re-scored at the flush: ask for its fp32 row now, so the flush finds it in L2
          s_queue[qp] = ((unsigned long long)row << 32) | (unsigned long long)i;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(A.p_hat + i * 64));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(A.p_hat + i * 64 + 32));
        } else {
          rescore(row, i);       // queue full: in place
        }
      }
    }
    fence_before_sync();
    // The one CTA barrier of a tile: accumulator b and the tile's g / margin are free after it, and it carries the
    // flush decision (a thread that pushed past the queue's half mark, or parked a key, votes yes).  Survivors wait
    // in the queue until then, so late in the run a flush - and the barriers it costs - happens once in many tiles.
    const int do_flush = __syncthreads_or((want_flush || t + 1 == ntile) ? 1 : 0);
    want_flush = false;
    if (do_flush) flush();
  }
  // final lists: merge every user once more and emit the best KMAX keys (one warp per user)
  __syncthreads();
  for (int uu = warp; uu < nu; uu += SC_THREADS / 32) warp_merge(uu, A.part + ((u0 + uu) * A.nsplit + split) * SC_KMAX);
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static int sc_splits(int64_t n_users, int64_t I) {
  const int64_t tiles = (n_users + SC_UT - 1) / SC_UT;
  int64_t want = std::max<int64_t>(1, (int64_t)num_sms() / tiles);          // one CTA per SM: whole waves only
  const int64_t ntile = (I + SC_IT - 1) / SC_IT;
  want = std::min<int64_t>(std::min<int64_t>(want, std::max<int64_t>(1, ntile / 16)), 32);    // 32 x 128 keys = the merge kernel's limit
  return (int)want;
}
}  // namespace ncf

using namespace ncf;

// defined in ncf_score.cu
namespace ncf {
int launch_topk_merge(const unsigned long long* part, int nsplit, int k, int64_t n_users, int64_t* idx, float* score, cudaStream_t st);
}

extern "C" int64_t ncf_item_image_bytes(int64_t I) { return ((std::max<int64_t>(I, 1) + SC_IT - 1) / SC_IT) * (int64_t)SC_TILE_BYTES; }

extern "C" int ncf_item_image(const float* p_hat, const float* g, int64_t I, void* img, void* stream) {
  NCF_REQUIRE(p_hat && g && img && I >= 1, "item_image: bad argument");
  const int64_t threads = ((I + SC_IT - 1) / SC_IT) * SC_IT * 8;
  item_image_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p_hat, g, I, static_cast<uint8_t*>(img));
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

extern "C" int64_t ncf_score_topk_tc_workspace_bytes(int64_t n_users, int64_t I, int32_t k) {
  (void)k;
  const int64_t n = std::max<int64_t>(n_users, 1);
  const int64_t tiles = (n + SC_UT - 1) / SC_UT;
  const int ns = sc_splits(n, I);
  return align_up(n * ns * SC_KMAX * 8, 256) + align_up(tiles * ns * SC_CTA_WS, 256);
}

namespace ncf {
int score_topk_impl(const ncf_tables* T, const float* dense, const float* p_hat, const float* g, const int64_t* user_ids,
                    int64_t n_users, int64_t I, int32_t k, int64_t* topk_idx, float* topk_score, void* workspace,
                    int64_t workspace_bytes, void* stream);
static int score_topk_tc_impl(const ncf_tables* T, const float* dense, const float* p_hat, const float* g, const void* img,
                              const int64_t* user_ids, int64_t n_users, int64_t I, int32_t k, int64_t* topk_idx,
                              float* topk_score, void* workspace, int64_t workspace_bytes, void* stream);
}
extern "C" int ncf_score_topk_tc(const ncf_tables* T, const float* dense, const float* p_hat, const float* g, const void* img,
                                 const int64_t* user_ids, int64_t n_users, int64_t I, int32_t k, int64_t* topk_idx,
                                 float* topk_score, void* workspace, int64_t workspace_bytes, void* stream) {
  NCF_REQUIRE(dense && user_ids, "score_topk_tc: null argument");
  return score_topk_tc_impl(T, dense, p_hat, g, img, user_ids, n_users, I, k, topk_idx, topk_score, workspace, workspace_bytes, stream);
}

// Exact top-k of sigmoid(q . v_i + bias_i) for raw 64-d query rows against I vectors (order: score descending, ties -> lowest
// index): the retrieval step of the serving design (cosine similarity when both sides are L2-normalised; the sigmoid is
// monotone and keeps the keys positive).  image: ncf_item_image(vectors, bias) or NULL (exact fp32 kernel only).
extern "C" int64_t ncf_dot_topk_workspace_bytes(int64_t n, int64_t I, int32_t k, int32_t with_image) {
  return with_image ? ncf_score_topk_tc_workspace_bytes(n, I, k) : ncf_score_topk_workspace_bytes(n, I, k);
}
extern "C" int ncf_dot_topk(const float* queries, int64_t n, const float* vectors, const float* bias, const void* image, int64_t I,
                            int32_t k, int64_t* topk_idx, float* topk_score, void* workspace, int64_t workspace_bytes, void* stream) {
  NCF_REQUIRE(queries && vectors && bias && n >= 0, "dot_topk: null argument");
  ncf_tables T{};
  T.w[0] = const_cast<float*>(queries);
  T.rows_user = std::max<int64_t>(n, 1);
  if (image) return score_topk_tc_impl(&T, nullptr, vectors, bias, image, nullptr, n, I, k, topk_idx, topk_score, workspace, workspace_bytes, stream);
  return score_topk_impl(&T, nullptr, vectors, bias, nullptr, n, I, k, topk_idx, topk_score, workspace, workspace_bytes, stream);
}

static int ncf::score_topk_tc_impl(const ncf_tables* T, const float* dense, const float* p_hat, const float* g, const void* img,
                                   const int64_t* user_ids, int64_t n_users, int64_t I, int32_t k, int64_t* topk_idx,
                                   float* topk_score, void* workspace, int64_t workspace_bytes, void* stream) {
  NCF_REQUIRE(T && p_hat && g && img && topk_idx && topk_score && workspace, "score_topk_tc: null argument");
  NCF_REQUIRE(k >= 1 && k <= SC_KMAX, "score_topk_tc: k=%d outside [1,%d]", k, SC_KMAX);
  NCF_REQUIRE(I >= 1 && I < ((int64_t)1 << 32) - 1, "score_topk_tc: bad catalogue size");
  if (n_users == 0) return NCF_OK;
  const int64_t need = ncf_score_topk_tc_workspace_bytes(n_users, I, k);
  if (workspace_bytes < need) {
    set_error("score_topk_tc: workspace %lld < %lld", (long long)workspace_bytes, (long long)need);
    return NCF_ERR_WORKSPACE;
  }
  const int ns = sc_splits(n_users, I);
  const int64_t tiles = (n_users + SC_UT - 1) / SC_UT;
  NCF_REQUIRE(tiles < ((int64_t)1 << 31), "score_topk_tc: too many users in one call");
  cudaStream_t st = (cudaStream_t)stream;
  static bool configured = false;
  if (!configured) {
    NCF_CUDA(cudaFuncSetAttribute(score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCS_TOTAL));
    configured = true;
  }
  ScoreTcArgs A{};
  A.t_umf = T->w[0];
  A.dense = dense;
  A.p_hat = p_hat;
  A.g = g;
  A.img = static_cast<const uint8_t*>(img);
  A.user_ids = user_ids;
  A.n_users = n_users;
  A.I = I;
  A.rows_user = T->rows_user;
  A.status = T->status;
  A.nsplit = ns;
  A.part = static_cast<unsigned long long*>(workspace);
  A.cand = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + align_up(n_users * ns * SC_KMAX * 8, 256));
  dim3 grid((unsigned)tiles, ns);
  score_tc_kernel<<<grid, SC_THREADS, SCS_TOTAL, st>>>(A);
  NCF_LAUNCH_CHECK();
  return launch_topk_merge(A.part, ns, k, n_users, topk_idx, topk_score, st);
}
