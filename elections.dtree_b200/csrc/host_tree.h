// host_tree.h — host-side state of one Dirichlet-tree: parameters, the observed-count trie
// and its level-ordered CSR flattening. Replaces DirichletTree<IRVNode,...> + IRVParameters
// (dirichlet_tree.h:24-157, irv_node.h:22-154) on the host; sampling never runs here.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "dtree_types.h"

namespace dtb {

struct HostParams {
  uint32_t nc = 0, min_depth = 0, max_depth = 0;
  double a0 = 1.0;
  bool vd = false;
  std::vector<double> depth_factors;  // irv_node.cpp:13-25; stale after set_min_depth (irv_node.h:127)
  void recompute_depth_factors();
  // Checks a CSR batch of ballots (candidate indices) the way the Rcpp layer does before it updates
  // (R_tree.cpp:24-43,139-151). On failure returns false with `err` set.
  bool validate(const uint8_t *prefs, const uint64_t *offsets, uint64_t n, uint64_t *n_short, uint64_t *depth_mask,
                std::string &err) const;
};

// CSR image of the trie, ready for upload, plus what the exports need.
struct FlatTree {
  std::vector<uint32_t> node_base;
  std::vector<uint32_t> node_total;  // per node: sum of its child-slot counts
  std::vector<uint32_t> slot_count;
  std::vector<int32_t> slot_child;
  std::vector<uint8_t> slot_cand;
  std::vector<uint8_t> node_depth;
  std::vector<uint32_t> node_parent_slot;  // global slot index in the parent that leads here (root: UINT32_MAX)
  std::vector<uint32_t> level_first;       // first node id of each depth, plus n_nodes at the end
  // Observed multiset derived from the counts: ballots in CSR + counts.
  std::vector<uint8_t> obs_prefs;
  std::vector<uint64_t> obs_offsets;
  std::vector<uint32_t> obs_counts;
  uint32_t n_nodes() const { return (uint32_t)node_depth.size(); }
  // prefix (candidate sequence) of node i, by walking parent slots
  void node_prefix(uint32_t node, std::vector<uint8_t> &out) const;
};

class HostTree {
 public:
  HostParams P;

  explicit HostTree(const HostParams &p);
  void reset();
  // Inserts ballots [first, n) of a validated CSR batch (IRVNode::update once per ballot).
  void insert_ballots(const uint8_t *prefs, const uint64_t *offsets, uint64_t first, uint64_t n);
  uint64_t n_nodes() const { return depth_.size(); }

  void flatten(FlatTree &out) const;

 private:
  // Arena trie. Node i: depth_[i]; K = nc - depth; counts_[base_[i] .. +K] (last = stop slot);
  // child_[base_[i] .. +K-1] arena index or -1. Slot j of a node is candidate path[depth + j]
  // under the swap rule of irv_node.cpp:203-220.
  std::vector<uint8_t> depth_;
  std::vector<uint64_t> base_;
  std::vector<uint32_t> counts_;
  std::vector<int32_t> child_;
  int32_t new_node(uint32_t depth);
  void insert(const uint8_t *b, uint32_t len);
};

uint64_t hash_seed(const char *seed, size_t len);

}  // namespace dtb
