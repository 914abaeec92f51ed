// Shared device/host helpers for libncf_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "ncf_b200.h"

namespace ncf {

constexpr int D = NCF_D;
constexpr int H1 = NCF_H1, H2 = NCF_H2, H3 = NCF_H3;
constexpr int HEADS = NCF_HEADS, HD = NCF_D / NCF_HEADS;
constexpr int TDIM = NCF_TDIM;
constexpr int K0 = NCF_D + NCF_TDIM;  // mlp.0 in_features = 96
constexpr float LN_EPS = 1e-5f;

// ---- error plumbing ---------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define NCF_CUDA(call)                                                          \
  do {                                                                          \
    cudaError_t e__ = (call);                                                   \
    if (e__ != cudaSuccess) return ::ncf::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)
extern unsigned long long g_launches;   // kernels launched by this library (bench.py's gpu_launches)
#define NCF_LAUNCH_CHECK()          \
  do {                              \
    ++::ncf::g_launches;            \
    NCF_CUDA(cudaGetLastError());   \
  } while (0)
#define NCF_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::ncf::set_error(__VA_ARGS__);           \
      return NCF_ERR_ARG;                      \
    }                                          \
  } while (0)
#define NCF_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != NCF_OK) return rc__; \
  } while (0)

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------
// The tower kernels of one step form a chain of persistent grids, each with a prologue (weight images, parameters, tensor
// memory) that depends on nothing the chain computes.  Launched with cudaLaunchAttributeProgrammaticStreamSerialization,
// a kernel's CTAs take over an SM as soon as the previous kernel's CTA on it exits, run the prologue, and only then wait
// (griddepcontrol.wait) for the previous grid to have completed: the tail of one kernel overlaps the prologue of the
// next.  Rules kept by every kernel that uses this: nothing written, and nothing another kernel of the step produces
// read, before pdl_wait().  NCF_PDL=0 launches everything the ordinary way (the instructions are no-ops then).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
// which launches carry the attribute (NCF_PDL = bit mask, A/B switch): 1 attention forward, 2 MLP forward, 4 loss + head
// backward, 8 MLP backward, 16 MLP weight gradients + reduce, 32 attention backward + reduce.  Default 60: the BACKWARD
// chain only - during the forward the auxiliary stream's id sort and dense-equivalent sweep live on the SM resources the
// gather / attention kernels leave free, and an early-resident MLP forward CTA (200 KB of shared memory) would take them.
enum { PDL_ATTN_FWD = 1, PDL_MLP_FWD = 2, PDL_HEAD = 4, PDL_MLP_BWD = 8, PDL_MLP_WGRAD = 16, PDL_ATTN_BWD = 32 };
inline int pdl_mask() {
  static int m = -1;
  if (m < 0) {
    const char* e = getenv("NCF_PDL");
    m = e ? atoi(e) : 0;
  }
  return m;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int which, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_mask() & which) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- dense flat layout --------------------------------------------------------------------
struct DenseLayout {
  int64_t off[NCF_P_COUNT];
  int64_t size[NCF_P_COUNT];
  int64_t total;
};
constexpr DenseLayout make_layout() {
  DenseLayout L{};
  const int64_t sizes[NCF_P_COUNT] = {
      D, D, D, D,
      D * D, D * D, D * D, D * D, D, D, D, D,
      (int64_t)H1 * K0, H1, H1, H1,
      (int64_t)H2 * H1, H2, H2, H2,
      (int64_t)H3 * H2, H3, H3, H3,
      D, 1, H3, 1, 2, 1};
  int64_t o = 0;
  for (int i = 0; i < NCF_P_COUNT; ++i) {
    L.off[i] = o;
    L.size[i] = sizes[i];
    o += (sizes[i] + 3) / 4 * 4;  // keep every tensor 16-byte aligned
  }
  L.total = o;
  return L;
}
constexpr DenseLayout kLayout = make_layout();   // host-side table
// compile-time offsets usable in device code
template <int ID>
struct DenseOff {
  static constexpr int64_t value = make_layout().off[ID];
};
#define NCF_OFF(id) (::ncf::DenseOff<id>::value)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// ncf_set_sm_reserve: SMs the persistent tower kernels (one or two CTAs per SM that own its registers / shared memory)
// leave free, so that a communication kernel queued on another stream (NCCL) starts at once instead of at the end of
// the tower kernel that happens to be running.  Process-wide; 0 by default.
inline int& sm_reserve() {
  static int r = getenv("NCF_SM_RESERVE") ? atoi(getenv("NCF_SM_RESERVE")) : 0;
  return r;
}
inline int tower_sms() { return std::max(1, num_sms() - sm_reserve()); }
// Persistent kernels walk their tiles round-robin: `ctas` CTAs need ceil(ntiles / ctas) rounds, and the smallest grid with
// that round count finishes at the same time while leaving the other SMs to whatever waits on another stream (the id sort
// and the dense-equivalent sweep of the auxiliary stream).  2,560 tiles: 143 CTAs instead of 148, 18 rounds either way.
inline int even_grid(int64_t ntiles, int64_t ctas) {
  static const bool on = !(getenv("NCF_EVEN_GRID") && getenv("NCF_EVEN_GRID")[0] == '0');     // A/B switch
  ctas = std::max<int64_t>(1, std::min(ntiles, ctas));
  if (!on || ntiles <= 0) return (int)ctas;
  const int64_t rounds = (ntiles + ctas - 1) / ctas;
  return (int)((ntiles + rounds - 1) / rounds);
}

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// bump allocator over the caller-provided workspace
struct Carver {
  char* base;
  int64_t used = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(int64_t count) {
    used = align_up(used, 256);
    T* p = base ? reinterpret_cast<T*>(base + used) : nullptr;
    used += count * (int64_t)sizeof(T);
    return p;
  }
};

// ---- device helpers -------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the 16 lanes of a half warp (lanes keep their half)
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// caller-supplied table index -> always inside [0, rows): see "Id validation" in ncf_b200.h
__device__ __forceinline__ int64_t clamp_id(int64_t id, int64_t rows) { return id < 0 ? 0 : (id >= rows ? rows - 1 : id); }
__device__ __forceinline__ bool bad_id(int64_t id, int64_t rows) { return (unsigned long long)id >= (unsigned long long)rows; }
__device__ __forceinline__ void flag_status(int32_t* status, int word) {
  if (status) *reinterpret_cast<volatile int32_t*>(status + word) = 1;      // plain store: every writer stores 1
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// 4 consecutive values of a [N,64] row: fp32 (16 B) or, for the tcgen05 attention block, bf16 (8 B)
__device__ __forceinline__ void st_row4(float* base, int64_t n, int c, float4 v, bool bf16_rows) {
  if (bf16_rows) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(base) + n * 64 + c) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
  } else {
    *reinterpret_cast<float4*>(base + n * 64 + c) = v;
  }
}
// four columns c..c+3 of row n of a [*,64] row array held as fp32 or (bf16_rows) as bf16 (see st_row4)
__device__ __forceinline__ float4 ld_row4(const float* base, int64_t n, int c, bool bf16_rows) {
  if (bf16_rows) {
    const uint2 q = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(base) + n * 64 + c));
    return make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u), __uint_as_float(q.y << 16),
                       __uint_as_float(q.y & 0xffff0000u));
  }
  return __ldg(reinterpret_cast<const float4*>(base + n * 64 + c));
}
__device__ __forceinline__ float4 f4_fma(float a, float4 x, float4 y) {
  return make_float4(fmaf(a, x.x, y.x), fmaf(a, x.y, y.y), fmaf(a, x.z, y.z), fmaf(a, x.w, y.w));
}
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float f4_dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float f4_hsum(float4 a) { return (a.x + a.y) + (a.z + a.w); }

// Philox4x32-10 (Salmon et al. 2011): counter-based, so every dropout element is a pure function
// of (seed, step, site, element index) - the backward and the test-side mask dump regenerate it.
template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) { return philox4x32<10>(ctr, key); }
// Dropout stream: one Philox call yields 8 x 16-bit lanes = the keep decisions of 8 consecutive
// elements (element idx uses call idx >> 3, lane idx & 7).  The low 15 bits of a lane are the uniform
// draw: keep iff (lane & 0x7fff) >= thresh (p quantised to 1/32768; kept values are scaled by 1/(1-p)
// with the nominal p like nn.Dropout).  15-bit draws leave a guard bit per lane, so the two decisions of
// a 32-bit Philox word come out of ONE subtraction (drop_signs) - the tensor-core MLP stores them in the
// sign bits of its saved (non-negative) ReLU outputs and the backward never regenerates the stream.
struct DropoutRng {
  uint2 key;
  uint32_t step_lo, step_hi_site;
  uint32_t thresh;  // 15-bit threshold; 0 = dropout off
  float scale;      // 1/(1-p)
  __device__ __forceinline__ uint4 draw8(uint64_t group) const {
    // Philox4x32-7: the smallest round count that passes BigCrush (Salmon et al. 2011, table 2); the dropout draw is
    // ~30 % of the MLP forward epilogue's instructions, and three rounds less is a tenth of that kernel
    return philox4x32<7>(make_uint4((uint32_t)group, (uint32_t)(group >> 32), step_lo, step_hi_site), key);
  }
  __device__ __forceinline__ bool keep16(uint32_t lane16) const { return (lane16 & 0x7fffu) >= thresh; }
  __device__ __forceinline__ float mask(float x, uint32_t lane16) const { return keep16(lane16) ? x * scale : 0.f; }
  // 0x8000 in each half of the result whose lane is DROPPED (w = one Philox word = two lanes)
  __device__ __forceinline__ uint32_t drop_signs(uint32_t w) const {
    const uint32_t d = ((w & 0x7fff7fffu) | 0x80008000u) - (thresh * 0x10001u);
    return ~d & 0x80008000u;
  }
  __device__ __forceinline__ bool keep(uint64_t idx) const {
    const uint4 r = draw8(idx >> 3);
    const uint32_t c = ((uint32_t)idx >> 1) & 3u;
    const uint32_t w = c == 0 ? r.x : c == 1 ? r.y : c == 2 ? r.z : r.w;
    return keep16(((uint32_t)idx & 1u) ? (w >> 16) : (w & 0xffffu));
  }
  // 8 consecutive elements starting at e0 (multiple of 8)
  __device__ __forceinline__ void apply8(uint64_t e0, float* v) const {
    const uint4 r = draw8(e0 >> 3);
    v[0] = mask(v[0], r.x & 0xffffu); v[1] = mask(v[1], r.x >> 16);
    v[2] = mask(v[2], r.y & 0xffffu); v[3] = mask(v[3], r.y >> 16);
    v[4] = mask(v[4], r.z & 0xffffu); v[5] = mask(v[5], r.z >> 16);
    v[6] = mask(v[6], r.w & 0xffffu); v[7] = mask(v[7], r.w >> 16);
  }
  // 4 consecutive elements starting at e0 (multiple of 4)
  __device__ __forceinline__ void apply4(uint64_t e0, float* v) const {
    const uint4 r = draw8(e0 >> 3);
    const bool hi = (e0 & 4) != 0;
    const uint32_t a = hi ? r.z : r.x, b = hi ? r.w : r.y;
    v[0] = mask(v[0], a & 0xffffu); v[1] = mask(v[1], a >> 16);
    v[2] = mask(v[2], b & 0xffffu); v[3] = mask(v[3], b >> 16);
  }
  // 2 consecutive elements starting at e0 (multiple of 2)
  __device__ __forceinline__ void apply2(uint64_t e0, float* v) const {
    const uint4 r = draw8(e0 >> 3);
    const uint32_t c = ((uint32_t)e0 >> 1) & 3u;
    const uint32_t w = c == 0 ? r.x : c == 1 ? r.y : c == 2 ? r.z : r.w;
    v[0] = mask(v[0], w & 0xffffu); v[1] = mask(v[1], w >> 16);
  }
};
__host__ __device__ inline DropoutRng make_rng(const ncf_run_cfg& cfg, int site) {
  DropoutRng r;
  r.key = make_uint2((uint32_t)cfg.seed, (uint32_t)(cfg.seed >> 32));
  r.step_lo = (uint32_t)cfg.step;
  r.step_hi_site = ((uint32_t)(cfg.step >> 32) & 0x00ffffffu) | ((uint32_t)site << 24);
  const double p = (cfg.training && cfg.dropout_p > 0.f) ? (double)cfg.dropout_p : 0.0;
  double t = p * 32768.0 + 0.5;
  r.thresh = t >= 32767.0 ? 32767u : (uint32_t)t;
  if (p == 0.0) r.thresh = 0u;
  r.scale = (float)(1.0 / (1.0 - p));   // nn.Dropout: kept values / (1-p)
  return r;
}
#endif  // __CUDACC__

}  // namespace ncf
