// lane.cu — instantiates lane_kernel (see lane.cuh).
#include "lane.cuh"
#include "launch.h"

namespace dtb {

cudaError_t launch_lane(const SimArgs &a, const LaunchCfg &cfg) {
  if (cfg.block != 128) return cudaErrorInvalidValue;
  if (cfg.log_gamma) lane_kernel<128, true><<<cfg.grid, 128, 0, cfg.stream>>>(a);
  else lane_kernel<128, false><<<cfg.grid, 128, 0, cfg.stream>>>(a);
  return cudaGetLastError();
}

int lane_blocks_per_sm(bool lg) {
  int n = 0;
  cudaError_t e = lg ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lane_kernel<128, true>, 128, 0)
                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lane_kernel<128, false>, 128, 0);
  return e == cudaSuccess ? n : 0;
}

}  // namespace dtb
