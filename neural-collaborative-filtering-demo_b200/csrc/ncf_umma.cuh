// tcgen05 / TMEM / mbarrier primitives for sm_100a (inline PTX), and the shared-memory operand
// layout used by every tensor-core kernel of this library.
//
// Operand tiles are kept in the canonical NO-SWIZZLE ("interleaved") UMMA layout: 8x8 core matrices
// of 16-bit elements, each 128 contiguous bytes (8 rows x 16 B).  For a [R x C] bf16 tile
//     byte_off(r, c) = (r/8) * (C*16) + (c/8) * 128 + (r%8) * 16 + (c%8) * 2
// Read as a K-major operand (rows = M or N, cols = K):     LBO = 128,    SBO = C*16
// Read as an MN-major operand (cols = M or N, rows = K):   LBO = C*16,   SBO = 128
// so ONE image of a tile serves a GEMM and its transpose (forward / dgrad / wgrad) with no copy.
// (CuTe: cute/atom/mma_traits_sm100.hpp "canonical layouts", LayoutType::INTERLEAVE.)
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "ncf_common.cuh"

namespace ncf {
namespace umma {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__host__ __device__ constexpr uint32_t tile_off(uint32_t r, uint32_t c, uint32_t C) {
  return (r >> 3) * (C * 16u) + (c >> 3) * 128u + (r & 7u) * 16u + (c & 7u) * 2u;
}

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading /
// stride byte offsets (all >> 4), version = 1 (sm_100), no swizzle, base offset 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// 32-bit instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16, bf16 x bf16 -> fp32
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                         // c_format = F32
         | (1u << 7)                       // a_format = BF16
         | (1u << 10)                      // b_format = BF16
         | ((a_mn_major ? 1u : 0u) << 15)  // a_major
         | ((b_mn_major ? 1u : 0u) << 16)  // b_major
         | ((uint32_t)(N >> 3) << 17)      // n_dim
         | ((uint32_t)(M >> 4) << 24);     // m_dim
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// same with the A operand in TENSOR MEMORY (K-major, lane = row, two bf16 per 32-bit column: the low half is the
// even k); a K = 16 step covers 8 columns
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// all MMAs issued so far by this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (the tensor core's operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// TMEM allocation: one full warp; ncols power of two >= 32; the base address lands in *slot (smem)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (taddr.lane + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// the same load WITHOUT the wait: several loads can be in flight; tmem_ld_wait() before the first use of any of them
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
template <int W>
__device__ __forceinline__ void tmem_ldw(uint32_t taddr, float (&v)[W]) {
  if constexpr (W == 32) tmem_ld32(taddr, v);
  else tmem_ld16(taddr, v);
}

// registers -> TMEM, same 32 lanes x W columns shape as tmem_ld*; tmem_st_wait() before anyone reads them back
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
               :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])), "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
               : "memory");
}
__device__ __forceinline__ void tmem_st16u(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                  "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
               : "memory");
}
template <int W>
__device__ __forceinline__ void tmem_stw(uint32_t taddr, const float (&v)[W]) {
  if constexpr (W == 32) tmem_st32(taddr, v);
  else tmem_st16(taddr, v);
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- mbarrier ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t a = smem_addr(bar);
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  }
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (1-D, multiple of 16 bytes), completion counted on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}

// ask the memory system to bring [src, src + bytes) (16-byte multiples) into L2; no completion tracking
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// ---- bf16 packing --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ void unpack_bf16x8(uint4 q, float (&v)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// ---- operand tiles and GEMM issue ---------------------------------------------------------------
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


// GEMM with the A operand in tensor memory (see mma_bf16_ts): `ksteps` MMAs of K = 16
__device__ __forceinline__ void issue_gemm_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_addr, uint32_t b_lbo, uint32_t b_sbo,
                                              uint32_t b_step, uint32_t idesc, int ksteps, bool accumulate_first) {
  for (int k = 0; k < ksteps; ++k)
    mma_bf16_ts(tmem_d, tmem_a + 8 * k, make_desc(b_addr + k * b_step, b_lbo, b_sbo), idesc, accumulate_first || k > 0);
}

}  // namespace umma
}  // namespace ncf
