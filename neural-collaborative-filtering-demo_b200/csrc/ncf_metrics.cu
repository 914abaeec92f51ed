// Ranking metrics of the reference (src/utils/metrics.py:110-265) on the device that holds the scores.
//
//   rank_metrics_kernel   HR@K / NDCG@K / MRR@K / MAP@K for every K in one pass over the [groups, M] score matrix:
//                         one warp per group; the only thing the four metrics need is the RANK of each positive inside
//                         its group (0-based position in the stable descending order: #scores above + #equal scores at a
//                         lower index), so nothing is sorted.
//   auc_*                 sklearn.roc_auc_score == Mann-Whitney U with ties counted 1/2.  The SMALLER class is sorted
//                         (a bitonic network over its order-preserving 32-bit keys), every element of the larger class
//                         binary-searches it; wins and ties are counted in 64-bit integers, so the result is exact and
//                         independent of the order of the atomics.
#include "ncf_common.cuh"

namespace ncf {

constexpr int RM_MAX_K = 8;
struct RankArgs {
  const float* scores;
  const float* targets;
  int64_t groups;
  int M;
  int n_k;
  int k[RM_MAX_K];
  double* out;      // [n_k][4] sums over the groups: hit, ndcg, mrr, ap
};

__global__ void __launch_bounds__(256) rank_metrics_kernel(RankArgs A) {
  __shared__ double s_acc[RM_MAX_K * 4];
  if (threadIdx.x < RM_MAX_K * 4) s_acc[threadIdx.x] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float hit[RM_MAX_K], dcg[RM_MAX_K], mrr[RM_MAX_K], ap[RM_MAX_K];
#pragma unroll
  for (int q = 0; q < RM_MAX_K; ++q) hit[q] = dcg[q] = mrr[q] = ap[q] = 0.f;
  for (int64_t g = warp; g < A.groups; g += nwarps) {
    const float* s = A.scores + g * A.M;
    const float* t = A.targets + g * A.M;
    // number of positives of the group (ideal DCG) and, per K, the per-group sums
    int npos = 0;
    for (int j = lane; j < A.M; j += 32) npos += t[j] == 1.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) npos += __shfl_xor_sync(0xffffffffu, npos, o);
    float g_dcg[RM_MAX_K], g_ap[RM_MAX_K];
    int g_first[RM_MAX_K], g_rel[RM_MAX_K];
#pragma unroll
    for (int q = 0; q < RM_MAX_K; ++q) {
      g_dcg[q] = g_ap[q] = 0.f;
      g_first[q] = 0x7fffffff;
      g_rel[q] = 0;
    }
    for (int p = 0; p < A.M; ++p) {            // every positive of the group (usually exactly one)
      if (t[p] != 1.0f) continue;               // uniform across the warp
      const float sp = s[p];
      int above = 0, pos_above = 0;             // rank of p, and how many POSITIVES rank at or above it (precision)
      for (int j = lane; j < A.M; j += 32) {
        const float sj = s[j];
        const bool before = sj > sp || (sj == sp && j < p);
        above += before;
        pos_above += before && t[j] == 1.0f;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        above += __shfl_xor_sync(0xffffffffu, above, o);
        pos_above += __shfl_xor_sync(0xffffffffu, pos_above, o);
      }
#pragma unroll
      for (int q = 0; q < RM_MAX_K; ++q) {
        if (q >= A.n_k) break;
        const int kk = min(A.k[q], A.M);
        if (above < kk) {
          g_dcg[q] += 1.0f / log2f((float)above + 2.0f);
          g_ap[q] += (float)(pos_above + 1) / (float)(above + 1);
          g_first[q] = min(g_first[q], above);
          g_rel[q] += 1;
        }
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < RM_MAX_K; ++q) {
        if (q >= A.n_k) break;
        const int kk = min(A.k[q], A.M);
        float idcg = 0.f;
        for (int i = 0; i < min(kk, npos); ++i) idcg += 1.0f / log2f((float)i + 2.0f);
        hit[q] += g_rel[q] > 0 ? 1.f : 0.f;
        dcg[q] += idcg > 0.f ? g_dcg[q] / idcg : 0.f;
        mrr[q] += g_rel[q] > 0 ? 1.0f / (float)(g_first[q] + 1) : 0.f;
        ap[q] += g_rel[q] > 0 ? g_ap[q] / (float)g_rel[q] : 0.f;
      }
    }
  }
  if (lane == 0) {
    for (int q = 0; q < A.n_k; ++q) {
      atomicAdd(&s_acc[q * 4 + 0], (double)hit[q]);
      atomicAdd(&s_acc[q * 4 + 1], (double)dcg[q]);
      atomicAdd(&s_acc[q * 4 + 2], (double)mrr[q]);
      atomicAdd(&s_acc[q * 4 + 3], (double)ap[q]);
    }
  }
  __syncthreads();
  if (threadIdx.x < A.n_k * 4) atomicAdd(A.out + threadIdx.x, s_acc[threadIdx.x]);
}

// ---- AUC ---------------------------------------------------------------------------------------------------------
// order-preserving key of a float (handles negative scores too): larger float <=> larger key
__device__ __forceinline__ uint32_t float_key(float x) {
  uint32_t b = __float_as_uint(x);
  if (b == 0x80000000u) b = 0u;        // -0.0 == +0.0
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
// counts[0] = positives, counts[1] = negatives, counts[2] = accuracy hits at the threshold
__global__ void auc_count_kernel(const float* __restrict__ targets, const float* __restrict__ scores, int64_t n, float threshold,
                                 unsigned long long* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = i < n;
  const bool pos = in && targets[i] > 0.5f;
  const bool acc = in && ((scores[i] >= threshold) == pos);
  const unsigned mp = __ballot_sync(0xffffffffu, pos), mi = __ballot_sync(0xffffffffu, in), ma = __ballot_sync(0xffffffffu, acc);
  if ((threadIdx.x & 31) == 0) {
    if (mp) atomicAdd(counts + 0, (unsigned long long)__popc(mp));
    if (mi & ~mp) atomicAdd(counts + 1, (unsigned long long)__popc(mi & ~mp));
    if (ma) atomicAdd(counts + 2, (unsigned long long)__popc(ma));
  }
}
// keys of the smaller class, compacted (order irrelevant: they get sorted); cursor = counts[3]
__global__ void auc_compact_kernel(const float* __restrict__ targets, const float* __restrict__ scores, int64_t n,
                                   const unsigned long long* __restrict__ counts, uint32_t* __restrict__ keys,
                                   unsigned long long* __restrict__ cursor, int64_t cap) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool small_is_pos = counts[0] <= counts[1];
  const bool pos = targets[i] > 0.5f;
  if (pos == small_is_pos) {
    const unsigned long long at = atomicAdd(cursor, 1ull);
    if (at < (unsigned long long)cap) keys[at] = float_key(scores[i]);      // beyond the caller's capacity hint: reported by finish
  }
}
__global__ void auc_pad_kernel(const unsigned long long* __restrict__ counts, uint32_t* __restrict__ keys, int64_t padded) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long m = min(counts[0], counts[1]);
  if (i < padded && (unsigned long long)i >= m) keys[i] = 0xffffffffu;       // above every real key: sorts to the end
}
// one compare-exchange stage of the ascending bitonic network
__global__ void auc_bitonic_kernel(uint32_t* __restrict__ keys, int64_t padded, int64_t j, int64_t k) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= padded / 2) return;
  const int64_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
  const int64_t l = i | j;
  const uint32_t x = keys[i], y = keys[l];
  const bool asc = (i & k) == 0;
  if (asc ? x > y : x < y) {
    keys[i] = y;
    keys[l] = x;
  }
}
// every element of the LARGER class against the sorted smaller class: out[0] += 2 * (pairs the positive wins) + ties
__global__ void auc_search_kernel(const float* __restrict__ targets, const float* __restrict__ scores, int64_t n,
                                  const unsigned long long* __restrict__ counts, const uint32_t* __restrict__ keys,
                                  unsigned long long* __restrict__ twice_u, int64_t cap) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long mine = 0;
  const bool small_is_pos = counts[0] <= counts[1];
  const long long m = (long long)min(min(counts[0], counts[1]), (unsigned long long)cap);
  if (i < n && m > 0) {
    const bool pos = targets[i] > 0.5f;
    if (pos != small_is_pos) {
      const uint32_t key = float_key(scores[i]);
      long long lo = 0, hi = m;                 // lower bound: first index with keys[idx] >= key
      while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (keys[mid] < key) lo = mid + 1;
        else hi = mid;
      }
      const long long below = lo;
      hi = m;                                   // upper bound: first index with keys[idx] > key
      while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (keys[mid] <= key) lo = mid + 1;
        else hi = mid;
      }
      const long long ties = lo - below, above = m - lo;
      // this element is a NEGATIVE when the sorted class holds the positives: positives above it win
      const long long wins = small_is_pos ? above : below;
      mine = 2ull * (unsigned long long)wins + (unsigned long long)ties;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(twice_u, mine);
}
__global__ void auc_finish_kernel(const unsigned long long* __restrict__ counts, int64_t n, int64_t cap, double* __restrict__ out) {
  const double p = (double)counts[0], q = (double)counts[1];
  out[0] = (counts[0] == 0 || counts[1] == 0) ? nan("") : (double)counts[4] / (2.0 * p * q);     // sklearn: undefined for one class
  if (min(counts[0], counts[1]) > (unsigned long long)cap) out[0] = -1.0;      // the smaller class did not fit the capacity hint
  out[1] = n > 0 ? (double)counts[2] / (double)n : nan("");
  out[2] = p;
  out[3] = q;
}

static int64_t next_pow2(int64_t x) {
  int64_t p = 1;
  while (p < x) p <<= 1;
  return p;
}
}  // namespace ncf

using namespace ncf;

extern "C" int ncf_rank_metrics(const float* scores, const float* targets, int64_t groups, int32_t M, const int32_t* k_values,
                                int32_t n_k, double* out, void* stream) {
  NCF_REQUIRE(scores && targets && k_values && out && groups >= 0 && M >= 1, "rank_metrics: bad argument");
  NCF_REQUIRE(n_k >= 1 && n_k <= RM_MAX_K, "rank_metrics: between 1 and %d values of K", RM_MAX_K);
  cudaStream_t st = (cudaStream_t)stream;
  NCF_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * 4 * n_k, st));
  if (groups == 0) return NCF_OK;
  RankArgs A{};
  A.scores = scores;
  A.targets = targets;
  A.groups = groups;
  A.M = M;
  A.n_k = n_k;
  for (int q = 0; q < n_k; ++q) {
    NCF_REQUIRE(k_values[q] >= 1, "rank_metrics: K must be positive");
    A.k[q] = k_values[q];
  }
  A.out = out;
  const int grid = (int)std::min<int64_t>((groups + 7) / 8, (int64_t)num_sms() * 8);
  rank_metrics_kernel<<<grid, 256, 0, st>>>(A);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}

static int64_t auc_cap(int64_t n, int64_t small_class_cap) {
  // the smaller class has at most n / 2 elements; the caller may know a tighter bound (one positive per ranking group)
  const int64_t bound = small_class_cap > 0 ? std::min<int64_t>(small_class_cap, n / 2 + 1) : n / 2 + 1;
  return next_pow2(std::max<int64_t>(bound, 2));
}
extern "C" int64_t ncf_auc_workspace_bytes(int64_t n, int64_t small_class_cap) {
  return align_up(auc_cap(n, small_class_cap) * 4, 256) + 256;       // sort buffer (padded to a power of two) + counters
}

extern "C" int ncf_auc(const float* scores, const float* targets, int64_t n, int64_t small_class_cap, float threshold, double* out,
                       void* workspace, int64_t workspace_bytes, void* stream) {
  NCF_REQUIRE(scores && targets && out && workspace && n >= 0, "auc: bad argument");
  NCF_REQUIRE(workspace_bytes >= ncf_auc_workspace_bytes(n, small_class_cap), "auc: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t cap = auc_cap(n, small_class_cap);
  uint32_t* keys = static_cast<uint32_t*>(workspace);
  unsigned long long* counts = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + align_up(cap * 4, 256));
  NCF_CUDA(cudaMemsetAsync(counts, 0, 64, st));
  const unsigned grid = (unsigned)std::max<int64_t>((n + 255) / 256, 1);
  if (n > 0) {
    auc_count_kernel<<<grid, 256, 0, st>>>(targets, scores, n, threshold, counts);
    NCF_LAUNCH_CHECK();
    auc_compact_kernel<<<grid, 256, 0, st>>>(targets, scores, n, counts, keys, counts + 3, cap);
    NCF_LAUNCH_CHECK();
    // The size of the smaller class is known on the device only; the network is sized for its upper bound n/2 (padding
    // keys sort to the end).  Stages: log2(cap) * (log2(cap) + 1) / 2 launches of cap / 2 threads.
    auc_pad_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, st>>>(counts, keys, cap);
    NCF_LAUNCH_CHECK();
    for (int64_t k = 2; k <= cap; k <<= 1)
      for (int64_t j = k >> 1; j > 0; j >>= 1) {
        auc_bitonic_kernel<<<(unsigned)((cap / 2 + 255) / 256), 256, 0, st>>>(keys, cap, j, k);
        NCF_LAUNCH_CHECK();
      }
    auc_search_kernel<<<grid, 256, 0, st>>>(targets, scores, n, counts, keys, counts + 4, cap);
    NCF_LAUNCH_CHECK();
  }
  auc_finish_kernel<<<1, 1, 0, st>>>(counts, n, cap, out);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}
