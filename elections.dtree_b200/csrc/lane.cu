// lane.cu — instantiates lane_kernel (see lane.cuh).
#include "lane.cuh"
#include "launch.h"

namespace dtb {

static uint32_t lane_tile_leaves(const DeviceParams &P) {
  const uint32_t d0 = P.nc + (P.min_depth == 0u ? 1u : 0u);  // outcomes of the root, the widest node
  uint32_t dp = 2;
  while (dp < d0) dp <<= 1;
  return dp;
}

static size_t lane_smem_bytes_per_warp(const DeviceParams &P) { return LaneShared::bytes(P.nc, lane_tile_leaves(P)); }

uint64_t lane_scratch_bytes(uint32_t cap_items, uint32_t cap_queue, uint32_t cap_big) {
  return LaneScratch::bytes(cap_items, cap_queue, cap_big);
}

template <bool LOG>
static cudaError_t prepare(size_t smem) {
  return cudaFuncSetAttribute(lane_kernel<kLaneMaxBlock, LOG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

cudaError_t launch_lane(const SimArgs &a, const LaunchCfg &cfg) {
  if (cfg.block < 32 || cfg.block > kLaneMaxBlock || cfg.block % 32) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(cfg.block / 32) * lane_smem_bytes_per_warp(a.P);
  cudaError_t e = cfg.log_gamma ? prepare<true>(smem) : prepare<false>(smem);
  if (e != cudaSuccess) return e;
  if (cfg.log_gamma) lane_kernel<kLaneMaxBlock, true><<<cfg.grid, cfg.block, smem, cfg.stream>>>(a);
  else lane_kernel<kLaneMaxBlock, false><<<cfg.grid, cfg.block, smem, cfg.stream>>>(a);
  return cudaGetLastError();
}

int lane_warps_per_sm(const DeviceParams &P, bool lg) {
  // asked in blocks of one warp: the answer is then limited by registers and shared memory per warp alone
  const size_t smem = lane_smem_bytes_per_warp(P);
  if ((lg ? prepare<true>(smem) : prepare<false>(smem)) != cudaSuccess) return 0;
  int n = 0;
  cudaError_t e = lg ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lane_kernel<kLaneMaxBlock, true>, 32, smem)
                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lane_kernel<kLaneMaxBlock, false>, 32, smem);
  if (e != cudaSuccess) return 0;
  return n < kLaneMaxBlock / 32 ? n : kLaneMaxBlock / 32;
}

}  // namespace dtb
