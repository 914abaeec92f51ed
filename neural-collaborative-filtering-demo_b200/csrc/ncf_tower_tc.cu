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
