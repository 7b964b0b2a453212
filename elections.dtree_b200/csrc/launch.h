// launch.h — host-callable entry points of the kernels, one translation unit per key width W.
#pragma once
#include <cuda_runtime.h>

#include "simulate.cuh"

namespace dtb {

struct LaunchCfg {
  int block;        // threads per block: one of kBlockSizes
  bool log_gamma;   // small-alpha regime (Gamma variates carried as logarithms)
  int grid;
  cudaStream_t stream;
};

constexpr int kBlockSizes[] = {32, 64, 128, 256};
constexpr int kNumBlockSizes = 4;

// Launches simulate_kernel<W, block, log>. Returns cudaErrorInvalidValue for an unsupported block size.
template <int W>
cudaError_t launch_simulate(const SimArgs &a, const LaunchCfg &cfg);
// The same elections as three launches (sim_big_kernel, sim_walk_kernel, sim_irv_kernel): cfg.grid = elections of the
// wave, one scratch slot each.
template <int W>
cudaError_t launch_simulate_phased(const SimArgs &a, const LaunchCfg &cfg);
// Resident blocks per SM for that instantiation (occupancy API).
template <int W>
int simulate_blocks_per_sm(uint32_t nc, int block, bool log_gamma);

// deferred_kernel (deferred.cuh): one warp per election, `block` threads per block (64, 128 or 256).
cudaError_t launch_deferred(const SimArgs &a, const LaunchCfg &cfg);
int deferred_blocks_per_sm(int block, bool log_gamma);
uint64_t deferred_scratch_bytes(uint32_t cap_items, uint32_t cap_queue, uint32_t cap_big);  // per warp

// lane_kernel (lane.cuh): one thread per election, 32 elections per warp. Compiled for blocks of up to kLaneMaxBlock
// threads (one block fills an SM's register file at 80 registers per thread); launched with any multiple of 32 up to that.
// A job that fits the machine in one pass runs as ONE block per SM of as many warps as it takes (warps of a block start
// together and stay close in the instruction stream: measured 6 % faster than six 128-thread blocks per SM).
#ifndef DTB_LANE_MAX_BLOCK
#define DTB_LANE_MAX_BLOCK 768
#endif
constexpr int kLaneMaxBlock = DTB_LANE_MAX_BLOCK;
constexpr int kLaneBlock = 128;  // the unit the tuning key blocks_per_sm counts in (round-1 block size)
cudaError_t launch_lane(const SimArgs &a, const LaunchCfg &cfg);
int lane_warps_per_sm(const DeviceParams &P, bool log_gamma);  // resident warps per SM (registers, shared memory)
uint64_t lane_scratch_bytes(uint32_t cap_items, uint32_t cap_queue, uint32_t cap_big);  // per election

// Stand-alone IRV over entries already placed in block 0's scratch (dtree_irv).
struct IrvArgs {
  uint32_t nc;
  uint32_t n_entries;
  uint64_t seed;
  uint8_t *scratch;
  uint32_t cap_entries;
  uint8_t *order_out;       // [nc] device
  uint32_t *tie_rounds_out; // [1] device
};
template <int W>
cudaError_t launch_irv_only(const IrvArgs &a, cudaStream_t stream);

}  // namespace dtb
