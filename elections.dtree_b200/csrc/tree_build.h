// tree_build.h — builds the level-ordered CSR tree (dtree_types.h: DeviceTree) and the packed observed
// multiset ON THE DEVICE from the log of observed ballots. Replaces, for the device mirror, the host walk
// of IRVNode::update (irv_node.cpp:176-221) + DirichletTree::update's std::map (dirichlet_tree.h:182-193):
// the result is array-for-array identical to HostTree::flatten, which is pinned to the reference.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "devbuf.h"
#include "dtree_types.h"

namespace dtb {

// Everything the device holds about one tree.
struct DeviceTreeStore {
  // the mirror the sampling kernels read
  DevBuf<uint32_t> node_base, node_total, slot_count, obs_cnt, obs_tally0;
  DevBuf<int32_t> slot_child;
  DevBuf<uint8_t> slot_cand, obs_first;
  DevBuf<uint64_t> obs_key;
  uint32_t n_nodes = 0, n_slots = 0, n_obs = 0;
  // ballot log as uploaded so far (CSR, absolute offsets) and per-ballot slot sequences: entry j of a
  // ballot is the child-slot index its j-th preference takes at depth j under the swap rule
  DevBuf<uint8_t> log_prefs, log_slots;
  DevBuf<uint64_t> log_off;
  uint64_t log_ballots = 0, log_bytes = 0;
  // build-time only
  DevBuf<uint64_t> node_key;
  DevBuf<uint32_t> cursor, block_sums;
  DevBuf<uint32_t> level_tab;   // per depth {first node, nodes, first slot, -}, then {nodes, slots, observed entries, -}
  uint32_t *h_tab = nullptr;    // pinned copy of level_tab
  void release_all();
};

struct BuildStats {
  uint64_t h2d_bytes = 0;
  uint32_t launches = 0, host_syncs = 0;
};

// Appends ballots [S.log_ballots, n_ballots) of the host log (prefs/offsets as given to dtree_update,
// already validated) to the device log and rebuilds the mirror. Returns a CUDA error; `too_big` is set
// when the tree needs more than 2^32-16 slots. When the arrays sized by a priori bounds (a level has at most
// min(#ballots, parents x children) nodes) fit `bound_budget_bytes`, the whole build is ONE cooperative kernel
// (grid-wide barriers between the phases of every level) or, with allow_cooperative = false, the same phases as ~4
// launches per level enqueued without a host round trip; otherwise the host reads every level's size to allocate
// exactly. uniform_len: every ballot of the log has that many preferences (UINT32_MAX: lengths differ): the offsets are
// then written by a kernel instead of being uploaded (8 bytes per ballot of host-to-device traffic saved).
cudaError_t build_tree_on_device(DeviceTreeStore &S, uint32_t nc, const uint8_t *h_prefs, const uint64_t *h_offsets,
                                 uint64_t n_ballots, cudaStream_t stream, BuildStats *stats, bool *too_big,
                                 uint64_t bound_budget_bytes = 1ull << 30, bool allow_cooperative = true,
                                 uint32_t uniform_len = 0xFFFFFFFFu);

}  // namespace dtb
