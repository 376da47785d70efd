// Counter-based dropout masks (statistical parity with nn.Dropout: keep with probability 1 - p,
// survivors scaled by 1 / (1 - p); reference call sites src/model.py:195-201 -- the nn.Dropout
// between c_fc and QuickGELU --, src/mmbt.py:56,82 and the hidden / attention-probability dropouts
// of pytorch_pretrained_bert's BertModel).  The decision for element `idx` of the tensor a site
// drops is a pure function of (seed, site, idx), so the backward REGENERATES the mask instead of
// storing it, and the CPU oracle (oracle/dropout.py) restates the same integer arithmetic: masks
// are compared bit for bit in the tests, the arithmetic around them at the usual tolerance.
//
//   h = idx * 0x9E3779B1 + lo(site_seed);  h ^= h >> 16;  h *= 0x7FEB352D;  h ^= h >> 15;
//   h += hi(site_seed);                    h *= 0x846CA68B;  h ^= h >> 16;      (all mod 2^32)
//   keep  <=>  h >= floor(p * 2^32)
//   site_seed = splitmix64(seed + 0x9E3779B97F4A7C15 * (site + 1))
#pragma once
#include <cstdint>

namespace mmu {
namespace dropout {

__host__ __device__ inline unsigned long long site_seed(unsigned long long seed, unsigned int site) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (static_cast<unsigned long long>(site) + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__host__ __device__ inline unsigned int threshold(float p) {
  if (!(p > 0.f)) return 0u;
  const double t = static_cast<double>(p) * 4294967296.0;
  return t >= 4294967295.0 ? 4294967295u : static_cast<unsigned int>(t);
}

__host__ __device__ __forceinline__ unsigned int hash(unsigned int idx, unsigned int lo, unsigned int hi) {
  unsigned int h = idx * 0x9E3779B1u + lo;
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h += hi;
  h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}

// Parameters of one dropout site as the kernels carry them (thresh == 0: dropout off).
struct Site {
  unsigned int lo, hi, thresh;
  float scale;  // 1 / (1 - p)
  __host__ __device__ __forceinline__ bool on() const { return thresh != 0u; }
  __host__ __device__ __forceinline__ bool keep(unsigned int idx) const { return hash(idx, lo, hi) >= thresh; }
  // multiplier applied to element idx: 0 or 1 / (1 - p)
  __host__ __device__ __forceinline__ float mult(unsigned int idx) const { return keep(idx) ? scale : 0.f; }
};

__host__ __device__ inline Site make_site(float p, unsigned long long seed, unsigned int site) {
  Site s;
  const unsigned long long ss = site_seed(seed, site);
  s.lo = static_cast<unsigned int>(ss);
  s.hi = static_cast<unsigned int>(ss >> 32);
  s.thresh = threshold(p);
  s.scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  return s;
}

}  // namespace dropout
}  // namespace mmu
