// deferred.cu — instantiates deferred_kernel (see deferred.cuh).
#include "deferred.cuh"
#include "launch.h"

namespace dtb {

template <int BLOCK, bool LOG>
static cudaError_t launch_d(const SimArgs &a, const LaunchCfg &cfg) {
  deferred_kernel<BLOCK, LOG><<<cfg.grid, BLOCK, 0, cfg.stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_deferred(const SimArgs &a, const LaunchCfg &cfg) {
  switch (cfg.block) {
    case 64: return cfg.log_gamma ? launch_d<64, true>(a, cfg) : launch_d<64, false>(a, cfg);
    case 128: return cfg.log_gamma ? launch_d<128, true>(a, cfg) : launch_d<128, false>(a, cfg);
    case 256: return cfg.log_gamma ? launch_d<256, true>(a, cfg) : launch_d<256, false>(a, cfg);
    default: return cudaErrorInvalidValue;
  }
}

uint64_t deferred_scratch_bytes(uint32_t cap_items, uint32_t cap_queue, uint32_t cap_big) {
  return DeferredScratch::bytes(cap_items, cap_queue, cap_big);
}

template <int BLOCK, bool LOG>
static int occ_d() {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, deferred_kernel<BLOCK, LOG>, BLOCK, 0) != cudaSuccess) return 0;
  return n;
}

int deferred_blocks_per_sm(int block, bool lg) {
  switch (block) {
    case 64: return lg ? occ_d<64, true>() : occ_d<64, false>();
    case 128: return lg ? occ_d<128, true>() : occ_d<128, false>();
    case 256: return lg ? occ_d<256, true>() : occ_d<256, false>();
    default: return 0;
  }
}

}  // namespace dtb
