// launch_shape.h — how many blocks of how many warps a lane_kernel job is launched with (host arithmetic only, so
// that the CPU tests can check it: tests/test_abi_cpu.py).
#pragma once
#include <algorithm>
#include <cstdint>

namespace dtb {

struct LaneShape {
  uint32_t warps_per_block;  // 1 .. warps_per_sm
  uint32_t grid;             // blocks
};

// batches: units of work a warp claims (elections / elections per warp); sms: multiprocessors; warps_per_sm: resident
// warps an SM can hold of this kernel; fit_warps: warps whose scratch fits the memory budget; forced_wpb: tuning key
// lane_block_warps (0: automatic).
// Automatic: ONE block per SM of as many warps as the job needs to run in a single pass — the warps of a block start
// together and stay close in the instruction stream — at most warps_per_sm; a larger job runs on full blocks whose
// warps claim batch after batch. Never more warps than have scratch.
inline LaneShape lane_shape(uint64_t batches, uint32_t sms, uint32_t warps_per_sm, uint64_t fit_warps, uint32_t forced_wpb) {
  batches = std::max<uint64_t>(batches, 1);
  sms = std::max<uint32_t>(sms, 1);
  warps_per_sm = std::max<uint32_t>(warps_per_sm, 1);
  fit_warps = std::max<uint64_t>(fit_warps, 1);
  uint64_t wpb = std::min<uint64_t>(warps_per_sm, std::max<uint64_t>(1, (batches + sms - 1) / sms));
  if (forced_wpb) wpb = std::min<uint64_t>(warps_per_sm, forced_wpb);
  wpb = std::max<uint64_t>(1, std::min(wpb, fit_warps));
  const uint64_t resident = (uint64_t)sms * (warps_per_sm / wpb), fit = std::max<uint64_t>(1, fit_warps / wpb);
  const uint64_t need = (batches + wpb - 1) / wpb;
  LaneShape s;
  s.warps_per_block = (uint32_t)wpb;
  s.grid = (uint32_t)std::max<uint64_t>(1, std::min(std::min(resident, fit), need));
  return s;
}

}  // namespace dtb
