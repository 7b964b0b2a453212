// debug.cu — test hooks (include/dtree_b200_debug.h): the device variate generators on their own.
#include <cuda_runtime.h>

#include "dtree_b200.h"
#include "dtree_b200_debug.h"
#include "host_tree.h"
#include "launch_shape.h"
#include "variates.cuh"

using namespace dtb;

namespace {

__global__ void philox_kernel(U4 c, uint32_t k0, uint32_t k1, U4 *out) { *out = philox4x32_10(c, k0, k1); }

template <bool EXACT>
__global__ void gamma_kernel(double alpha_d, int log_space, uint64_t seed, uint32_t n, float *out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  RngKey key = election_key(seed, i >> 5);
  uint64_t nk = child_node_key(kRootNodeKey, i & 31);
  uint32_t att = 0;
  if (alpha_d >= 4194304.0) {  // what a tree node does with an outcome that observed >= 2^22 ballots
    const uint32_t count = (uint32_t)alpha_d;
    const float prior = (float)(alpha_d - (double)count);
    out[i] = log_space ? outcome_gamma<true>(count, prior, key, nk, 0, att) : outcome_gamma<false>(count, prior, key, nk, 0, att);
    return;
  }
  const float alpha = (float)alpha_d;
  out[i] = log_space ? gamma_variate<true, EXACT>(alpha, key, nk, 0, att) : gamma_variate<false, EXACT>(alpha, key, nk, 0, att);
}

template <bool EXACT>
__global__ void binomial_kernel(uint32_t trials, float p, float q, uint64_t seed, uint32_t n, uint32_t *out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  RngKey key = election_key(seed, i >> 5);
  uint64_t nk = child_node_key(kRootNodeKey, i & 31);
  uint32_t att = 0;
  // same convention as the sampler: pass the smaller of (p, q), flip the result for the other
  bool flip = q < p;
  uint32_t k = binomial_small_p<EXACT>(trials, flip ? q : p, key, nk, 1, att);
  out[i] = flip ? trials - k : k;
}

__global__ void binv_law_kernel(uint32_t trials, float p, uint32_t *counts) {
  __shared__ uint32_t h[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
  __syncthreads();
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < (1u << 23); j += gridDim.x * blockDim.x)
    atomicAdd(&h[min(binv(trials, p, j << 9), 255u)], 1u);  // u01f keeps the top 23 bits
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    if (h[i]) atomicAdd(&counts[i], h[i]);
}

int cuda_fail(const char *what, cudaError_t e) {
  (void)what;
  return e == cudaSuccess ? DTREE_OK : DTREE_ERR_CUDA;
}

}  // namespace

extern "C" {

int dtree_debug_lane_shape(uint64_t batches, uint32_t sms, uint32_t warps_per_sm, uint64_t fit_warps, uint32_t forced_wpb,
                           uint32_t *warps_per_block, uint32_t *grid) {
  if (!warps_per_block || !grid) return DTREE_ERR_ARG;
  const LaneShape s = lane_shape(batches, sms, warps_per_sm, fit_warps, forced_wpb);
  *warps_per_block = s.warps_per_block;
  *grid = s.grid;
  return DTREE_OK;
}

int dtree_debug_philox(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4], int on_device) {
  U4 c{counter[0], counter[1], counter[2], counter[3]};
  U4 r;
  if (!on_device) {
    r = philox4x32_10(c, key[0], key[1]);
  } else {
    U4 *d = nullptr;
    cudaError_t e = cudaMalloc((void **)&d, sizeof(U4));
    if (e != cudaSuccess) return DTREE_ERR_CUDA;
    philox_kernel<<<1, 1>>>(c, key[0], key[1], d);
    e = cudaMemcpy(&r, d, sizeof(U4), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return DTREE_ERR_CUDA;
  }
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
  return DTREE_OK;
}

uint64_t dtree_debug_hash_seed(const char *seed, size_t seed_len) { return hash_seed(seed, seed_len); }

int dtree_debug_gamma_ex(double alpha, int log_space, uint64_t seed, uint32_t n, float *out, int device, int exact_only) {
  if (!(alpha > 0.0) || !out) return DTREE_ERR_ARG;
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return DTREE_ERR_CUDA;
  float *d = nullptr;
  cudaError_t e = cudaMalloc((void **)&d, (size_t)(n ? n : 1) * sizeof(float));
  if (e != cudaSuccess) return DTREE_ERR_CUDA;
  if (exact_only) gamma_kernel<true><<<(n + 255) / 256, 256>>>(alpha, log_space, seed, n, d);
  else gamma_kernel<false><<<(n + 255) / 256, 256>>>(alpha, log_space, seed, n, d);
  e = cudaMemcpy(out, d, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return cuda_fail("gamma", e);
}

int dtree_debug_gamma(double alpha, int log_space, uint64_t seed, uint32_t n, float *out, int device) {
  return dtree_debug_gamma_ex(alpha, log_space, seed, n, out, device, 0);
}

int dtree_debug_binomial_ex(uint32_t trials, double p, uint64_t seed, uint32_t n, uint32_t *out, int device, int exact_only) {
  if (!(p >= 0.0 && p <= 1.0) || !out) return DTREE_ERR_ARG;
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return DTREE_ERR_CUDA;
  uint32_t *d = nullptr;
  cudaError_t e = cudaMalloc((void **)&d, (size_t)(n ? n : 1) * sizeof(uint32_t));
  if (e != cudaSuccess) return DTREE_ERR_CUDA;
  if (exact_only) binomial_kernel<true><<<(n + 255) / 256, 256>>>(trials, (float)p, (float)(1.0 - p), seed, n, d);
  else binomial_kernel<false><<<(n + 255) / 256, 256>>>(trials, (float)p, (float)(1.0 - p), seed, n, d);
  e = cudaMemcpy(out, d, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return cuda_fail("binomial", e);
}

int dtree_debug_binomial(uint32_t trials, double p, uint64_t seed, uint32_t n, uint32_t *out, int device) {
  return dtree_debug_binomial_ex(trials, p, seed, n, out, device, 0);
}

int dtree_debug_binv_law(uint32_t trials, double p, uint32_t counts[256], int device) {
  if (!(p > 0.0 && p <= 0.5) || !counts || !((double)trials * p < 30.0)) return DTREE_ERR_ARG;
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return DTREE_ERR_CUDA;
  uint32_t *d = nullptr;
  cudaError_t e = cudaMalloc((void **)&d, 256 * sizeof(uint32_t));
  if (e != cudaSuccess) return DTREE_ERR_CUDA;
  e = cudaMemset(d, 0, 256 * sizeof(uint32_t));
  if (e == cudaSuccess) {
    binv_law_kernel<<<1024, 256>>>(trials, (float)p, d);
    e = cudaMemcpy(counts, d, 256 * sizeof(uint32_t), cudaMemcpyDeviceToHost);
  }
  cudaFree(d);
  return cuda_fail("binv", e);
}

}  // extern "C"
