// rng.cuh — counter-based random numbers for the sampler.
//
// Philox4x32-10 (Salmon et al., SC'11). Every variate of every election is a pure function of
//   key     = (seed_lo, seed_hi ^ election id)
//   counter = (node key lo, node key hi, stream tag, attempt)
// where the node key is a 64-bit hash chain over the node's prefix (root -> child by candidate).
// Nothing depends on which thread, block, launch or GPU evaluates a node, which is what makes
// 1/2/4/8-GPU totals bit-identical and lets different kernels reproduce each other's draws.
// Host and device share this file so the tests can check the generator against known answers.
#pragma once
#include <cstdint>

#include "dtree_types.h"

namespace dtb {

struct U4 {
  uint32_t x, y, z, w;
};

DT_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

DT_HD U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    U4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

DT_HD uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

constexpr uint64_t kRootNodeKey = 0x6a09e667f3bcc909ull;
DT_HD uint64_t child_node_key(uint64_t parent, uint32_t cand) {
  return mix64(parent + 0x9e3779b97f4a7c15ull * (uint64_t)(cand + 1));
}

// Stream tags (counter word 2): which variate of the node a Philox block feeds.
enum : uint32_t {
  kStreamGamma = 0x01000000u,     // + slot: Marsaglia-Tsang attempts for that outcome
  kStreamBinomial = 0x02000000u,  // + slot: binomial split of that outcome
  kStreamUrn = 0x03000000u,       // Polya-urn draws (small counts); counter word 3 = ball / 4
  kStreamOneHot = 0x04000000u,    // all-zero Dirichlet parameters: uniform one-hot (distributions.cpp:69-77)
  kStreamTie = 0x05000000u,       // IRV tie-break of a round (node key = 0, word 3 = round)
};

struct RngKey {
  uint32_t k0, k1;
};
DT_HD RngKey election_key(uint64_t seed, uint64_t election) {
  RngKey k;
  k.k0 = (uint32_t)seed ^ (uint32_t)(election >> 32) * 0x85ebca6bu;
  k.k1 = (uint32_t)(seed >> 32) ^ (uint32_t)election;
  return k;
}

// Out of line on the device: every variate calls it, and one copy instead of one per call site keeps the
// kernels' hot code inside the instruction cache (they were I-cache bound with it inlined).
#if defined(__CUDACC__)
static __device__ __noinline__ U4 node_random(RngKey k, uint64_t node_key, uint32_t stream, uint32_t attempt) {
#else
inline U4 node_random(RngKey k, uint64_t node_key, uint32_t stream, uint32_t attempt) {
#endif
  U4 c;
  c.x = (uint32_t)node_key;
  c.y = (uint32_t)(node_key >> 32);
  c.z = stream;
  c.w = attempt;
  return philox4x32_10(c, k.k0, k.k1);
}

// 23-bit uniform in (0,1): (k + 1/2) 2^-23 is exactly representable, so neither 0 nor 1 occurs.
DT_HD float u01f(uint32_t r) { return ((float)(r >> 9) + 0.5f) * 1.1920928955078125e-7f; }
// 32-bit uniform in (0,1) as a double.
DT_HD double u01d(uint32_t r) { return ((double)r + 0.5) * 2.3283064365386963e-10; }
// uniform integer in [0, n)
DT_HD uint32_t uniform_below(uint32_t r, uint32_t n) { return mulhi32(r, n); }

}  // namespace dtb
