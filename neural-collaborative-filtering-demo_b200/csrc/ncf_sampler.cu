// Input pipeline on the device: the negative sampler of SheetzDataset.__getitem__ /
// _sample_negative (data_prep.py:134-161, 181-212) fused with collate_recommender_batch (:230-320).
// One thread per sample row writes the key-major id columns and the targets directly:
//   row b*S+0   = the positive interaction            target 1
//   row b*S+s   = a negative for the same user        target 0
// Negative = draw from the inverse-popularity distribution (CDF search on a Philox uniform, what
// np.random.choice(p=w) does), reject the positive and the user's history (binary search in a sorted CSR
// row), at most 10 tries; then uniform over the complement (rejection on uniform draws, bounded, then a
// scan from a random start - the same distribution as np.random.choice(valid_negatives)).
#include "ncf_common.cuh"

namespace ncf {

__device__ __forceinline__ double u01(uint32_t a, uint32_t b) {      // 53-bit uniform in [0,1)
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ bool in_history(const int64_t* __restrict__ hist_off, const int64_t* __restrict__ hist_items,
                                           int64_t user, int64_t item) {
  if (!hist_off) return false;
  int64_t lo = hist_off[user], hi = hist_off[user + 1];
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int64_t v = hist_items[mid];
    if (v == item) return true;
    if (v < item) lo = mid + 1; else hi = mid;
  }
  return false;
}

__global__ void __launch_bounds__(256) sample_batch_kernel(const int64_t* __restrict__ pos_user, const int64_t* __restrict__ pos_item,
                                                           int64_t B, int S, const double* __restrict__ cdf, int64_t I,
                                                           const int64_t* __restrict__ hist_off,
                                                           const int64_t* __restrict__ hist_items, uint64_t seed, uint64_t step,
                                                           int64_t* __restrict__ user_ids, int64_t* __restrict__ item_ids,
                                                           float* __restrict__ targets) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= B * S) return;
  const int64_t b = n / S;
  const int s = (int)(n % S);
  const int64_t user = pos_user[b], pos = pos_item[b];
  user_ids[n] = user;
  targets[n] = s == 0 ? 1.0f : 0.0f;
  if (s == 0) {
    item_ids[n] = pos;
    return;
  }
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const double total = cdf[I - 1];
  int64_t chosen = -1;
  // 10 inverse-popularity draws (data_prep.py:143-151); 4 Philox calls give 16 words = 8 uniforms
  for (int call = 0; call < 3 && chosen < 0; ++call) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)(n >> 32), (uint32_t)step, (uint32_t)(step >> 32) ^ (0x51u + call)), key);
    const uint4 r2 = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)(n >> 32), (uint32_t)step, (uint32_t)(step >> 32) ^ (0xA1u + call)), key);
    const uint32_t w[8] = {r.x, r.y, r.z, r.w, r2.x, r2.y, r2.z, r2.w};
    for (int t = 0; t < 4 && chosen < 0; ++t) {
      if (call * 4 + t >= 10) break;
      const double u = u01(w[2 * t], w[2 * t + 1]) * total;
      int64_t lo = 0, hi = I - 1;                      // first index with cdf[idx] > u  (searchsorted side='right')
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (cdf[mid] > u) hi = mid; else lo = mid + 1;
      }
      if (lo != pos && !in_history(hist_off, hist_items, user, lo)) chosen = lo;
    }
  }
  // fallback: uniform over the complement (data_prep.py:153-161)
  for (int call = 0; call < 16 && chosen < 0; ++call) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)(n >> 32), (uint32_t)step, (uint32_t)(step >> 32) ^ (0xF00u + call)), key);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    for (int t = 0; t < 4 && chosen < 0; ++t) {
      const int64_t c = (int64_t)(((uint64_t)w[t] * (uint64_t)I) >> 32);
      if (c != pos && !in_history(hist_off, hist_items, user, c)) chosen = c;
    }
  }
  if (chosen < 0) {       // almost everything is in the history: scan from a random start
    const uint4 r = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)(n >> 32), (uint32_t)step, (uint32_t)(step >> 32) ^ 0xFFFFu), key);
    const int64_t start = (int64_t)(((uint64_t)r.x * (uint64_t)I) >> 32);
    for (int64_t k = 0; k < I && chosen < 0; ++k) {
      const int64_t c = (start + k) % I;
      if (c != pos && !in_history(hist_off, hist_items, user, c)) chosen = c;
    }
    if (chosen < 0) chosen = (pos + 1 + start % (I > 1 ? I - 1 : 1)) % I;   // no valid negative at all: any item != positive
  }
  item_ids[n] = chosen;
}

}  // namespace ncf

using namespace ncf;

extern "C" int ncf_sample_batch(const int64_t* pos_user, const int64_t* pos_item, int64_t B, int32_t S, const double* cdf,
                                int64_t I, const int64_t* hist_off, const int64_t* hist_items, uint64_t seed, uint64_t step,
                                int64_t* user_ids, int64_t* item_ids, float* targets, void* stream) {
  NCF_REQUIRE(pos_user && pos_item && cdf && user_ids && item_ids && targets, "sample_batch: null argument");
  NCF_REQUIRE(B >= 0 && S >= 1 && I >= 1, "sample_batch: bad sizes");
  NCF_REQUIRE((hist_off == nullptr) == (hist_items == nullptr), "sample_batch: history needs both offsets and items");
  if (B == 0) return NCF_OK;
  const int64_t n = B * S;
  sample_batch_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pos_user, pos_item, B, S, cdf, I, hist_off,
                                                                                     hist_items, seed, step, user_ids, item_ids,
                                                                                     targets);
  NCF_LAUNCH_CHECK();
  return NCF_OK;
}
