// host_tree.cpp — see host_tree.h.
#include "host_tree.h"

#include <algorithm>
#include <cstring>

namespace dtb {

// irv_node.cpp:13-25: F[max_depth-1] = 1; F[d-1] = F[d] * (#outcomes of a node at depth d).
void HostParams::recompute_depth_factors() {
  depth_factors.assign(max_depth, 0.0);
  double f = 1.0;
  for (int d = (int)max_depth - 1; d >= 0; --d) {
    depth_factors[d] = f;
    f *= (double)(nc - (uint32_t)d + ((uint32_t)d >= min_depth ? 1u : 0u));
  }
}

HostTree::HostTree(const HostParams &p) : P(p) {
  P.recompute_depth_factors();
  reset();
}

void HostTree::reset() {
  depth_.clear();
  base_.clear();
  counts_.clear();
  child_.clear();
  new_node(0);
}

int32_t HostTree::new_node(uint32_t depth) {
  uint32_t K = P.nc - depth;
  depth_.push_back((uint8_t)depth);
  base_.push_back(counts_.size());
  counts_.resize(counts_.size() + K + 1, 0u);
  child_.resize(child_.size() + K + 1, -1);
  return (int32_t)depth_.size() - 1;
}

// IRVNode::update, irv_node.cpp:176-221, iteratively. `path` starts as the identity and
// receives the reference's swaps, which fixes the slot <-> candidate mapping.
void HostTree::insert(const uint8_t *b, uint32_t len) {
  uint8_t path[kMaxCandidates];
  for (uint32_t i = 0; i < P.nc; ++i) path[i] = (uint8_t)i;
  int32_t cur = 0;
  for (;;) {
    uint32_t depth = depth_[cur], K = P.nc - depth;
    uint64_t base = base_[cur];
    if (depth == len) {  // :191-194 the ballot stops here
      counts_[base + K] += 1;
      return;
    }
    uint32_t i = depth;
    while (path[i] != b[depth]) ++i;  // :203-205
    uint32_t slot = i - depth;
    counts_[base + slot] += 1;
    if (K == 2) return;  // :210 the last two levels are never materialised
    int32_t c = child_[base + slot];
    if (c < 0) {
      c = new_node(depth + 1);
      child_[base + slot] = c;
    }
    std::swap(path[depth], path[i]);
    cur = c;
  }
}

bool HostParams::validate(const uint8_t *prefs, const uint64_t *offsets, uint64_t n, uint64_t *n_short,
                          uint64_t *depth_mask, std::string &err) const {
  const HostParams &P = *this;
  uint64_t shorts = 0, mask = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if (offsets[i + 1] < offsets[i]) { err = "ballot offsets must be non-decreasing"; return false; }
    uint64_t len = offsets[i + 1] - offsets[i];
    if (len > P.nc) { err = "ballot ranks more candidates than exist"; return false; }
    uint32_t seen = 0;
    for (uint64_t j = offsets[i]; j < offsets[i + 1]; ++j) {
      uint32_t c = prefs[j];
      if (c >= P.nc) { err = "Unknown candidate encountered in ballot!"; return false; }  // R_tree.cpp:30-31
      if (seen >> c & 1u) { err = "ballot ranks a candidate more than once"; return false; }
      seen |= 1u << c;
    }
    if (len > 0 && len < P.min_depth) ++shorts;  // R_tree.cpp:139
    mask |= 1ull << len;                          // R_tree.cpp:151
  }
  if (n_short) *n_short = shorts;
  if (depth_mask) *depth_mask = mask;
  return true;
}

void HostTree::insert_ballots(const uint8_t *prefs, const uint64_t *offsets, uint64_t first, uint64_t n) {
  for (uint64_t i = first; i < n; ++i) insert(prefs + offsets[i], (uint32_t)(offsets[i + 1] - offsets[i]));
}

void FlatTree::node_prefix(uint32_t node, std::vector<uint8_t> &out) const {
  out.resize(node_depth[node]);
  uint32_t cur = node;
  while (node_depth[cur] > 0) {
    uint32_t ps = node_parent_slot[cur];
    out[node_depth[cur] - 1] = slot_cand[ps];
    // parent = the node owning slot ps: last node whose base <= ps
    uint32_t parent = (uint32_t)(std::upper_bound(node_base.begin(), node_base.end(), ps) - node_base.begin()) - 1;
    cur = parent;
  }
}

// Breadth-first renumbering: nodes of one depth contiguous, siblings adjacent, so that a
// frontier processed in order reads slot_count with unit stride. One pass over a queue of arena
// ids; every output array is sized up front (the CSR has exactly as many slots as the arena).
void HostTree::flatten(FlatTree &F) const {
  const uint32_t nc = P.nc;
  const size_t n = depth_.size(), total_slots = counts_.size();
  F = FlatTree();
  F.node_base.resize(n + 1);
  F.node_total.resize(n);
  F.node_depth.resize(n);
  F.node_parent_slot.resize(n);
  F.slot_count.resize(total_slots);
  F.slot_child.resize(total_slots);
  F.slot_cand.resize(total_slots);
  F.obs_offsets.push_back(0);
  std::vector<int32_t> order(n);      // flat id -> arena id, filled as children are discovered
  std::vector<uint8_t> paths(n * nc);  // candidate permutation of every node (irv_node.cpp:219 swaps)
  order[0] = 0;
  for (uint32_t i = 0; i < nc; ++i) paths[i] = (uint8_t)i;
  F.node_parent_slot[0] = UINT32_MAX;
  size_t head = 0, tail = 1;
  uint32_t fbase = 0, cur_depth = UINT32_MAX;
  while (head < tail) {
    const size_t i = head++;
    const int32_t arena = order[i];
    const uint32_t depth = depth_[arena], K = nc - depth;
    const uint64_t base = base_[arena];
    const uint8_t *p = &paths[i * nc];
    if (depth != cur_depth) {
      F.level_first.push_back((uint32_t)i);
      cur_depth = depth;
    }
    F.node_base[i] = fbase;
    F.node_depth[i] = (uint8_t)depth;
    uint32_t child_total = 0;
    for (uint32_t j = 0; j < K; ++j) {
      const uint32_t cnt = counts_[base + j];
      child_total += cnt;
      F.slot_count[fbase + j] = cnt;
      F.slot_cand[fbase + j] = p[depth + j];
      const int32_t c = child_[base + j];
      if (c >= 0) {
        const size_t id = tail++;
        order[id] = c;
        uint8_t *q = &paths[id * nc];
        std::memcpy(q, p, nc);
        std::swap(q[depth], q[depth + j]);
        F.node_parent_slot[id] = fbase + j;
        F.slot_child[fbase + j] = (int32_t)id;
      } else {
        F.slot_child[fbase + j] = -1;
      }
    }
    const uint32_t stop = counts_[base + K];
    F.slot_count[fbase + K] = stop;
    F.slot_cand[fbase + K] = 0xFF;
    F.slot_child[fbase + K] = -1;
    F.node_total[i] = child_total;
    // Observed ballots that END at this node: the stop slot, and at K == 2 (never expanded,
    // irv_node.cpp:210) also the two child slots, which stand for prefix+cand (with or
    // without the implied last preference).
    if (stop) {
      F.obs_prefs.insert(F.obs_prefs.end(), p, p + depth);
      F.obs_offsets.push_back(F.obs_prefs.size());
      F.obs_counts.push_back(stop);
    }
    if (K == 2)
      for (uint32_t j = 0; j < 2; ++j)
        if (counts_[base + j]) {
          F.obs_prefs.insert(F.obs_prefs.end(), p, p + depth);
          F.obs_prefs.push_back(p[depth + j]);
          F.obs_offsets.push_back(F.obs_prefs.size());
          F.obs_counts.push_back(counts_[base + j]);
        }
    fbase += K + 1;
  }
  F.node_base[n] = fbase;
  F.level_first.push_back((uint32_t)n);
}

// FNV-1a over the seed bytes, then a splitmix64 finaliser: same string => same 64-bit key,
// which is all the reference's seed strings promise (R/dtree.R:682-684, test-determinism.R).
uint64_t hash_seed(const char *seed, size_t len) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < len; ++i) {
    h ^= (uint8_t)seed[i];
    h *= 0x100000001b3ull;
  }
  h += 0x9e3779b97f4a7c15ull;
  h = (h ^ (h >> 30)) * 0xbf58476d1ce4e5b9ull;
  h = (h ^ (h >> 27)) * 0x94d049bb133111ebull;
  return h ^ (h >> 31);
}

}  // namespace dtb
