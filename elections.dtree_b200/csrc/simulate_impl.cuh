// simulate_impl.cuh — included by simulate_w{1,2,3}.cu, which instantiate the kernels for one
// key width each so that the widths compile in parallel.
#pragma once
#include "launch.h"

namespace dtb {

// Dynamic shared memory of simulate_kernel: one stack of single-ballot frames per warp (walk_small_items).
template <int W>
static size_t simulate_smem_bytes(int block) {
  return (size_t)(block / 32) * WalkBudget<W>::kFused * sizeof(Item<W>);
}

// The three phase kernels over a wave of cfg.grid elections, block b on election b in scratch slot b. The walkers and
// the IRV stage run with large blocks (an election is finished sooner, so the tail of a wave is shorter).
#ifndef DTB_WALK_BLOCK
#define DTB_WALK_BLOCK 1024
#endif
#ifndef DTB_IRV_BLOCK
#define DTB_IRV_BLOCK 768
#endif
constexpr int kWalkBlock = DTB_WALK_BLOCK, kIrvBlock = DTB_IRV_BLOCK;
template <int W>
static size_t walk_smem_bytes() {
  return (size_t)(kWalkBlock / 32) * (WalkBudget<W>::kPhased + WalkBudget<W>::kPhased / 2) * sizeof(Item<W>);
}

template <int W, int BLOCK, bool LOG>
static cudaError_t launch_phased_one(const SimArgs &a, const LaunchCfg &cfg) {
  const size_t wsm = walk_smem_bytes<W>();
  cudaError_t e = cudaFuncSetAttribute(sim_walk_kernel<W, kWalkBlock>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsm);
  if (e != cudaSuccess) return e;
  const size_t ism = (size_t)(kIrvBlock / 32) * kPrivTallyWords * sizeof(uint32_t);
  e = cudaFuncSetAttribute(sim_irv_kernel<W, kIrvBlock>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ism);
  if (e != cudaSuccess) return e;
  sim_big_kernel<W, BLOCK, LOG><<<cfg.grid, BLOCK, 0, cfg.stream>>>(a);
  sim_walk_kernel<W, kWalkBlock><<<cfg.grid, kWalkBlock, wsm, cfg.stream>>>(a);
  sim_irv_kernel<W, kIrvBlock><<<cfg.grid, kIrvBlock, ism, cfg.stream>>>(a);
  return cudaGetLastError();
}

template <int W, bool LOG>
static cudaError_t launch_phased_block(const SimArgs &a, const LaunchCfg &cfg) {
  switch (cfg.block) {
    case 32: return launch_phased_one<W, 32, LOG>(a, cfg);
    case 64: return launch_phased_one<W, 64, LOG>(a, cfg);
    case 128: return launch_phased_one<W, 128, LOG>(a, cfg);
    case 256: return launch_phased_one<W, 256, LOG>(a, cfg);
    default: return cudaErrorInvalidValue;
  }
}

template <int W>
cudaError_t launch_simulate_phased(const SimArgs &a, const LaunchCfg &cfg) {
  return cfg.log_gamma ? launch_phased_block<W, true>(a, cfg) : launch_phased_block<W, false>(a, cfg);
}

template <int W, int BLOCK, bool LOG>
static cudaError_t launch_one(const SimArgs &a, const LaunchCfg &cfg) {
  size_t smem = simulate_smem_bytes<W>(BLOCK);
  cudaError_t e = cudaFuncSetAttribute(simulate_kernel<W, BLOCK, LOG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  simulate_kernel<W, BLOCK, LOG><<<cfg.grid, BLOCK, smem, cfg.stream>>>(a);
  return cudaGetLastError();
}

template <int W, bool LOG>
static cudaError_t launch_block(const SimArgs &a, const LaunchCfg &cfg) {
  switch (cfg.block) {
    case 32: return launch_one<W, 32, LOG>(a, cfg);
    case 64: return launch_one<W, 64, LOG>(a, cfg);
    case 128: return launch_one<W, 128, LOG>(a, cfg);
    case 256: return launch_one<W, 256, LOG>(a, cfg);
    default: return cudaErrorInvalidValue;
  }
}

template <int W>
cudaError_t launch_simulate(const SimArgs &a, const LaunchCfg &cfg) {
  return cfg.log_gamma ? launch_block<W, true>(a, cfg) : launch_block<W, false>(a, cfg);
}

template <int W, int BLOCK, bool LOG>
static int occ_one(uint32_t nc) {
  (void)nc;
  size_t smem = simulate_smem_bytes<W>(BLOCK);
  cudaFuncSetAttribute(simulate_kernel<W, BLOCK, LOG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, simulate_kernel<W, BLOCK, LOG>, BLOCK, smem) != cudaSuccess)
    return 0;
  return n;
}

template <int W>
int simulate_blocks_per_sm(uint32_t nc, int block, bool lg) {
  switch (block) {
    case 32: return lg ? occ_one<W, 32, true>(nc) : occ_one<W, 32, false>(nc);
    case 64: return lg ? occ_one<W, 64, true>(nc) : occ_one<W, 64, false>(nc);
    case 128: return lg ? occ_one<W, 128, true>(nc) : occ_one<W, 128, false>(nc);
    case 256: return lg ? occ_one<W, 256, true>(nc) : occ_one<W, 256, false>(nc);
    default: return 0;
  }
}

// dtree_irv: one block runs socialChoiceIRV over caller-supplied entries.
template <int W>
__global__ void __launch_bounds__(256) irv_only_kernel(const IrvArgs ia) {
  __shared__ SharedFixed<256> sm;
  __shared__ uint32_t priv[256 / 32][kPrivTallyWords];
  Scratch<W> sc;
  sc.bind(ia.scratch, 0, ia.cap_entries);
  SimArgs a;  // only the fields irv_election reads
  a.P.nc = ia.nc;
  a.n_obs = 0;
  a.n_winners = 0;  // 0: never stop early, dtree_irv wants the whole order
  a.elim_orders = nullptr;
  a.obs_key = nullptr;
  a.obs_cnt = nullptr;
  a.obs_first = nullptr;
  a.obs_tally0 = nullptr;
  uint32_t stat[kStatCount];
#pragma unroll
  for (int i = 0; i < kStatCount; ++i) stat[i] = 0;
  if (threadIdx.x == 0) sm.n_entries = ia.n_entries;
  __syncthreads();
  const RngKey key = election_key(ia.seed, 0);
  irv_election<W, 256>(a, sc, sm, priv[threadIdx.x >> 5], key, ia.nc - 1, stat);
  if (threadIdx.x == 0) {
    uint32_t standing = ~sm.eliminated & (ia.nc >= 32 ? 0xffffffffu : ((1u << ia.nc) - 1u));
    for (uint32_t r = 0; r + 1 < ia.nc; ++r) ia.order_out[r] = sm.order[r];
    ia.order_out[ia.nc - 1] = (uint8_t)(__ffs(standing) - 1);
    *ia.tie_rounds_out = sm.tie_rounds;
  }
}

template <int W>
cudaError_t launch_irv_only(const IrvArgs &a, cudaStream_t stream) {
  irv_only_kernel<W><<<1, 256, 0, stream>>>(a);
  return cudaGetLastError();
}

#define DTREE_INSTANTIATE(W)                                                          \
  template cudaError_t launch_simulate<W>(const SimArgs &, const LaunchCfg &);        \
  template cudaError_t launch_simulate_phased<W>(const SimArgs &, const LaunchCfg &); \
  template int simulate_blocks_per_sm<W>(uint32_t, int, bool);                        \
  template cudaError_t launch_irv_only<W>(const IrvArgs &, cudaStream_t);

}  // namespace dtb
