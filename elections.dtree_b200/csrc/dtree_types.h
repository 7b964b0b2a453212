// dtree_types.h — plain data shared by the host tree, the C ABI and the kernels.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define DT_HD __host__ __device__ __forceinline__
#else
#define DT_HD inline
#endif

namespace dtb {

constexpr int kMaxCandidates = 32;
constexpr int kPrefBits = 5;        // one candidate index
constexpr int kPrefsPerWord = 12;   // 60 of 64 bits used
constexpr int kMaxKeyWords = 3;     // ceil(31 / 12)

// A ranked ballot (or tree-node prefix) packed 5 bits per preference, first preference in
// the low bits of word 0. A sampled ballot never has more than n_candidates-1 preferences
// (irv_node.cpp:46: the last preference is implied), so W = ceil((nc-1)/12) words suffice.
DT_HD int key_words_for(uint32_t nc) { return nc <= 13 ? 1 : (nc <= 25 ? 2 : 3); }

template <int W>
struct BallotKey {
  uint64_t w[W];
};

template <int W>
DT_HD uint32_t key_get(const BallotKey<W> &k, int j) {
  return (uint32_t)(k.w[j / kPrefsPerWord] >> (kPrefBits * (j % kPrefsPerWord))) & 31u;
}
template <int W>
DT_HD void key_set(BallotKey<W> &k, int j, uint32_t cand) {
  k.w[j / kPrefsPerWord] |= (uint64_t)cand << (kPrefBits * (j % kPrefsPerWord));
}

// Sampling parameters as the kernels see them (IRVParameters, irv_node.h:22-154).
struct DeviceParams {
  uint32_t nc;
  uint32_t min_depth;
  uint32_t max_depth;
  uint32_t leaf_depth;     // max_depth - 1 in unsigned arithmetic (irv_node.cpp:136), UINT32_MAX if max_depth == 0
  float a_prior[kMaxCandidates];  // a0 * (vd ? depthFactor[depth] : 1) per depth (irv_node.cpp:35,115)
  int small_alpha;         // any a_prior[d] < 1: Gamma variates are carried in log space
  float binv_max;          // binomials with n p below this are drawn by inversion, the others by BTRS (>= 10)
};

// Level-ordered CSR mirror of the observed-count tree (IRVNode, tree_node.h:36-58).
// Node i owns slots [node_base[i], node_base[i+1]): K = nc - depth child slots, then the
// "ballot stops here" slot. Nodes of one depth are contiguous; siblings are adjacent.
struct DeviceTree {
  const uint32_t *node_base;   // [n_nodes + 1]
  const uint32_t *node_total;  // [n_nodes] sum of the node's K child-slot counts (stop slot excluded)
  const uint32_t *slot_count;  // observed counts (IRVNode::as)
  const int32_t *slot_child;   // child node id, -1 = uninitialised (IRVNode::children) / stop slot
  const uint8_t *slot_cand;    // candidate of the slot (path-dependent, irv_node.cpp:219), 0xFF = stop slot
  uint32_t n_nodes;
};

}  // namespace dtb
