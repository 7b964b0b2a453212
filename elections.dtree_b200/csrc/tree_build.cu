// tree_build.cu — see tree_build.h. Level-synchronous construction: at depth d every observed ballot adds
// one count to the slot its d-th preference (or its end) selects in its current node; slots that
// received a count and may have children (irv_node.cpp:210-215) are numbered by a prefix sum, which
// places the children of a level in breadth-first order — the order HostTree::flatten produces.
#include "tree_build.h"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>

namespace dtb {

void DeviceTreeStore::release_all() {
  node_base.release(); node_total.release(); slot_count.release(); obs_cnt.release(); obs_tally0.release();
  slot_child.release(); slot_cand.release(); obs_first.release(); obs_key.release();
  log_prefs.release(); log_slots.release(); log_off.release();
  node_key.release(); cursor.release(); block_sums.release(); level_tab.release();
  if (h_tab) cudaFreeHost(h_tab);
  h_tab = nullptr;
  n_nodes = n_slots = n_obs = 0;
  log_ballots = log_bytes = 0;
}

namespace {

#ifndef DTB_SCAN_BLOCK
#define DTB_SCAN_BLOCK 512
#endif
constexpr int kScanBlock = DTB_SCAN_BLOCK;
#ifndef DTB_SCAN_ITEMS
#define DTB_SCAN_ITEMS 1
#endif
constexpr int kScanItems = DTB_SCAN_ITEMS;
constexpr int kScanTile = kScanBlock * kScanItems;
constexpr uint32_t kDone = 0xFFFFFFFFu;
// Level table (device memory, so that no level has to wait for the host): row d = {first node, nodes, first slot, -}
// of depth d; row kTotRow = {nodes, slots, observed entries, -} of the whole tree.
constexpr int kTotRow = kMaxCandidates + 1;
constexpr int kTabWords = 4 * (kMaxCandidates + 2);

struct TreePtrs {
  uint32_t *node_base, *node_total, *slot_count;
  int32_t *slot_child;
  uint8_t *slot_cand;
  uint64_t *node_key;
};

// Slot index of every preference of ballots [first, n): position of the candidate among the not yet
// ranked ones, where ranking a candidate swaps it with the one at the front (irv_node.cpp:203-220).
__global__ void slot_sequences_kernel(const uint8_t *__restrict__ prefs, const uint64_t *__restrict__ off,
                                      uint8_t *__restrict__ slots, uint64_t first, uint64_t n, uint32_t nc) {
  for (uint64_t b = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t lo = off[b];
    const uint32_t len = (uint32_t)(off[b + 1] - lo);
    const uint32_t L = min(len, nc - 1u);
    uint8_t path[kMaxCandidates];
    for (uint32_t i = 0; i < nc; ++i) path[i] = (uint8_t)i;
    for (uint32_t d = 0; d < L; ++d) {
      const uint8_t c = prefs[lo + d];
      uint32_t i = d;
      while (i < nc - 1u && path[i] != c) ++i;
      slots[lo + d] = (uint8_t)(i - d);
      path[i] = path[d];
      path[d] = c;
    }
    for (uint32_t d = L; d < len; ++d) slots[lo + d] = 0;
  }
}

__global__ void init_root_kernel(TreePtrs T, uint32_t *obs_tally0, uint32_t *tab, uint32_t nc, int W) {
  const uint32_t j = threadIdx.x;
  for (uint32_t k = j; k < (uint32_t)kTabWords; k += blockDim.x) tab[k] = 0;
  if (j <= nc) {
    T.slot_count[j] = 0;
    T.slot_child[j] = -1;
    T.slot_cand[j] = j < nc ? (uint8_t)j : 0xFF;
  }
  if (j < (uint32_t)kMaxCandidates) obs_tally0[j] = 0;
  if (j < (uint32_t)W) T.node_key[j] = 0;
  __syncthreads();
  if (j == 0) {
    T.node_base[0] = 0;
    T.node_base[1] = nc + 1;
    tab[1] = 1;  // depth 0: node 0, one node, slots from 0
    tab[4 * kTotRow + 0] = 1;
    tab[4 * kTotRow + 1] = nc + 1;
  }
}

// One count per ballot still walking at `depth`; equal targets inside a warp are added once.
__global__ void count_level_kernel(TreePtrs T, const uint64_t *__restrict__ off, const uint8_t *__restrict__ slots,
                                   uint32_t *__restrict__ cursor, uint64_t n, uint32_t nc, uint32_t depth) {
  const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t idx = kDone;
  if (b < n) {
    const uint32_t cur = depth == 0 ? 0u : cursor[b];
    if (cur != kDone) {
      const uint32_t node = depth == 0 ? 0u : (uint32_t)T.slot_child[cur];
      const uint32_t base = T.node_base[node], K = nc - depth;
      const uint64_t lo = off[b];
      const uint32_t len = (uint32_t)(off[b + 1] - lo);
      if (len == depth) {  // irv_node.cpp:191-194
        idx = base + K;
        cursor[b] = kDone;
      } else {
        idx = base + slots[lo + depth];
        cursor[b] = K == 2 ? kDone : idx;  // irv_node.cpp:210
      }
    } else if (depth == 0) {
      cursor[b] = kDone;
    }
  }
  const unsigned peers = __match_any_sync(0xFFFFFFFFu, idx);
  if (idx != kDone && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&T.slot_count[idx], (uint32_t)__popc(peers));
}

// --- prefix sums over "how many things does element i produce" ---------------------------------

// Slot i of a level (slots [slot_lo, slot_lo + n), K + 1 per node) gets a child when it is a child slot
// with a count.
struct LevelFlag {
  const uint32_t *slot_count;
  const uint32_t *row;  // the level's row of the table
  uint32_t per_node;    // K + 1
  __device__ uint32_t operator()(uint32_t i) const {
    return ((i % per_node) != per_node - 1u && slot_count[row[2] + i] > 0u) ? 1u : 0u;
  }
};

// Node i contributes its distinct observed ballots: those that stop there and, where the tree ends
// (two candidates left), those that went on to either of them.
struct ObservedFlag {
  const uint32_t *node_base, *slot_count;
  __device__ uint32_t operator()(uint32_t i) const {
    const uint32_t base = node_base[i], K = node_base[i + 1] - base - 1u;
    uint32_t f = slot_count[base + K] > 0u;
    if (K == 2u) f += (slot_count[base] > 0u) + (slot_count[base + 1] > 0u);
    return f;
  }
};

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total) {
  __shared__ uint32_t warp_sums[kScanBlock / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  uint32_t before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < kScanBlock / 32; ++w) {
    const uint32_t s = warp_sums[w];
    if (w < warp) before += s;
    all += s;
  }
  __syncthreads();
  *total = all;
  return before + x - v;
}

// The number of elements of a scan is count[0] * mul, read on the device; blocks beyond it have nothing to do.
template <typename F>
__global__ void __launch_bounds__(kScanBlock)
    tile_sums_kernel(F f, const uint32_t *__restrict__ count, uint32_t mul, uint32_t *__restrict__ block_sums) {
  const uint32_t n = count[0] * mul;
  if (blockIdx.x * (uint32_t)kScanTile >= n) return;
  const uint32_t first = blockIdx.x * (uint32_t)kScanTile + threadIdx.x * (uint32_t)kScanItems;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (first + k < n) s += f(first + k);
  uint32_t total;
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// Exclusive scan of the tile sums in place (one block). The total goes into the table: for a level (depth >= 0)
// it is the number of nodes of the next depth, whose row and the tree totals are filled in; for the observed
// entries (depth < 0) it is their number.
__global__ void __launch_bounds__(kScanBlock)
    scan_sums_kernel(uint32_t *__restrict__ block_sums, const uint32_t *count, uint32_t mul, uint32_t *tab, int depth,
                     uint32_t per_node) {
  const uint32_t n = count[0] * mul;
  const uint32_t nb = (n + kScanTile - 1) / kScanTile;
  uint32_t carry = 0;
  for (uint32_t start = 0; start < nb; start += kScanBlock) {
    const uint32_t i = start + threadIdx.x;
    const uint32_t v = i < nb ? block_sums[i] : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(v, &total);
    if (i < nb) block_sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) {
    uint32_t *tot = tab + 4 * kTotRow;
    if (depth >= 0) {
      const uint32_t *row = tab + 4 * depth;
      uint32_t *next = tab + 4 * (depth + 1);
      next[0] = row[0] + row[1];
      next[1] = carry;
      next[2] = row[2] + row[1] * per_node;
      tot[0] = next[0] + carry;
      tot[1] = next[2] + carry * (per_node - 1u);  // a child has one slot less than its parent
    } else {
      tot[2] = carry;
    }
  }
}

// Numbers the children of a level and initialises them: slots, candidates (the parent's order with the
// taken candidate swapped to the front, irv_node.cpp:219), prefix key.
// A new node for slot j of parent pn (its slots start at pbase): its candidates are the parent's with the chosen one
// swapped to the front and dropped (irv_node.cpp:219). The parent's candidates are read before anything is written, so
// the loads are in flight together instead of alternating with stores to the same array.
__device__ __forceinline__ void write_child(const TreePtrs &T, uint32_t slot, uint32_t pbase, uint32_t j, uint32_t K,
                                            uint32_t child, uint32_t cbase, uint32_t parent, uint32_t depth, int W) {
  uint8_t pc[kMaxCandidates + 1];
#pragma unroll 8
  for (uint32_t q = 0; q < K; ++q) pc[q] = T.slot_cand[pbase + q];
  uint64_t pk[kMaxKeyWords];
  for (int w = 0; w < W; ++w) pk[w] = T.node_key[(size_t)parent * W + w];
  T.slot_child[slot] = (int32_t)child;
  T.node_base[child] = cbase;
  for (uint32_t q = 0; q + 1 < K; ++q) {
    T.slot_cand[cbase + q] = (q + 1 == j) ? pc[0] : pc[q + 1];
    T.slot_child[cbase + q] = -1;
    T.slot_count[cbase + q] = 0;
  }
  T.slot_cand[cbase + K - 1] = 0xFF;
  T.slot_child[cbase + K - 1] = -1;
  T.slot_count[cbase + K - 1] = 0;
  const uint64_t cand = pc[j];
  for (int w = 0; w < W; ++w) {
    uint64_t key = pk[w];
    if ((int)(depth / kPrefsPerWord) == w) key |= cand << (kPrefBits * (depth % kPrefsPerWord));
    T.node_key[(size_t)child * W + w] = key;
  }
}

__global__ void __launch_bounds__(kScanBlock)
    create_children_kernel(TreePtrs T, LevelFlag f, const uint32_t *__restrict__ tab, const uint32_t *__restrict__ block_sums,
                           uint32_t depth, int W) {
  const uint32_t *row = tab + 4 * depth, *next = row + 4;
  const uint32_t n = row[1] * f.per_node, level_first_node = row[0];
  const uint32_t first_new_node = next[0], n_new = next[1], new_slot_base = next[2];
  if (blockIdx.x == 0 && threadIdx.x == 0) T.node_base[first_new_node + n_new] = new_slot_base + n_new * (f.per_node - 1u);
  if (blockIdx.x * (uint32_t)kScanTile >= n) return;
  const uint32_t first = blockIdx.x * (uint32_t)kScanTile + threadIdx.x * (uint32_t)kScanItems;
  uint32_t flags = 0, s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (first + k < n && f(first + k)) {
      flags |= 1u << k;
      ++s;
    }
  uint32_t total;
  uint32_t rank = block_exclusive_scan(s, &total) + block_sums[blockIdx.x];
  const uint32_t K = f.per_node - 1u;  // children of a parent; a child has K - 1 children + stop = K slots
  for (int k = 0; k < kScanItems; ++k) {
    if (!(flags >> k & 1u)) continue;
    const uint32_t i = first + k, pn = i / f.per_node, j = i % f.per_node;
    const uint32_t slot = row[2] + i, pbase = row[2] + pn * f.per_node;
    const uint32_t child = first_new_node + rank, cbase = new_slot_base + rank * K;
    ++rank;
    write_child(T, slot, pbase, j, K, child, cbase, level_first_node + pn, depth, W);
  }
}

// Also the first-preference tallies of the observed ballots: slot c of the root is candidate c, and every ballot with a
// first preference was counted there (one copy of nc words instead of one contended atomic per observed entry).
__global__ void node_totals_kernel(TreePtrs T, const uint32_t *__restrict__ tab, uint32_t *obs_tally0, uint32_t nc) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nc) obs_tally0[i] = T.slot_count[i];
  if (i >= tab[4 * kTotRow]) return;
  const uint32_t base = T.node_base[i], K = T.node_base[i + 1] - base - 1u;
  uint32_t s = 0;
  for (uint32_t j = 0; j < K; ++j) s += T.slot_count[base + j];
  T.node_total[i] = s;
}

struct ObsOut {
  uint64_t *key;
  uint32_t *cnt, *tally0;
  uint8_t *first;
};

__device__ void write_observed(const TreePtrs &T, const ObsOut &O, uint32_t at, uint32_t node, uint32_t depth, int extra_cand,
                               uint32_t count, uint32_t nc, int W) {
  BallotKey<kMaxKeyWords> k;
  for (int w = 0; w < kMaxKeyWords; ++w) k.w[w] = w < W ? T.node_key[(size_t)node * W + w] : 0ull;
  uint32_t L = depth;
  if (extra_cand >= 0) key_set<kMaxKeyWords>(k, (int)L++, (uint32_t)extra_cand);
  uint32_t first = 0xFFu;
  if (L > 0) {
    first = key_get<kMaxKeyWords>(k, 0);
    for (uint32_t j = L; j < nc - 1u; ++j) key_set<kMaxKeyWords>(k, (int)j, first);  // capi.cu pack_ballot
  }
  for (int w = 0; w < W; ++w) O.key[(size_t)at * W + w] = k.w[w];
  O.cnt[at] = count;
  O.first[at] = (uint8_t)first;
}

__global__ void __launch_bounds__(kScanBlock)
    write_observed_kernel(TreePtrs T, ObservedFlag f, const uint32_t *__restrict__ tab, const uint32_t *__restrict__ block_sums,
                          ObsOut O, uint32_t nc, int W) {
  const uint32_t n_nodes = tab[4 * kTotRow];
  if (blockIdx.x * (uint32_t)kScanTile >= n_nodes) return;
  const uint32_t first = blockIdx.x * (uint32_t)kScanTile + threadIdx.x * (uint32_t)kScanItems;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (first + k < n_nodes) s += f(first + k);
  uint32_t total;
  uint32_t at = block_exclusive_scan(s, &total) + block_sums[blockIdx.x];
  if (s == 0) return;
  for (int k = 0; k < kScanItems; ++k) {
    const uint32_t i = first + k;
    if (i >= n_nodes) break;
    const uint32_t base = T.node_base[i], K = T.node_base[i + 1] - base - 1u, depth = nc - K;
    const uint32_t stop = T.slot_count[base + K];
    if (stop) write_observed(T, O, at++, i, depth, -1, stop, nc, W);
    if (K == 2u)
      for (uint32_t j = 0; j < 2u; ++j) {
        const uint32_t c = T.slot_count[base + j];
        if (c) write_observed(T, O, at++, i, depth, (int)T.slot_cand[base + j], c, nc, W);
      }
  }
}


// --- the whole build as ONE cooperative kernel --------------------------------------------------------------------
// The level-synchronous build above needs ~4 launches per tree level (39 for the 10-candidate benchmark tree), each a
// few microseconds of work: the rebuild was launch-latency-bound. Here the same phases run inside one persistent
// grid with grid-wide barriers between them (the level table lives in device memory already, so no phase ever needed
// the host). Used whenever the arrays sized by a priori bounds fit the budget; the multi-launch path remains for
// trees whose bounds do not (the host then reads every level's size to allocate exactly).
namespace cg = cooperative_groups;

struct CoopArgs {
  TreePtrs T;
  const uint8_t *prefs;
  const uint64_t *off;
  uint8_t *slots;
  uint32_t *cursor, *block_sums, *tab, *obs_tally0;
  ObsOut O;
  uint64_t first_new, n_ballots;
  uint32_t nc;
  int W;
};

template <typename F>
__device__ __forceinline__ void tile_sum_phase(F f, uint32_t n, uint32_t *block_sums) {
  const uint32_t n_tiles = (n + kScanTile - 1) / kScanTile;
  for (uint32_t vb = blockIdx.x; vb < n_tiles; vb += gridDim.x) {
    const uint32_t first = vb * (uint32_t)kScanTile + threadIdx.x * (uint32_t)kScanItems;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
      if (first + k < n) s += f(first + k);
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) block_sums[vb] = total;
  }
}

// Sum of block_sums[0..nb), to every thread of the calling block.
__device__ __forceinline__ uint32_t sum_tiles(const uint32_t *block_sums, uint32_t nb) {
  uint32_t v = 0;
  for (uint32_t i = threadIdx.x; i < nb; i += kScanBlock) v += block_sums[i];
  uint32_t total;
  block_exclusive_scan(v, &total);
  return total;
}

#ifndef DTB_COUNT_UNROLL
#define DTB_COUNT_UNROLL 1  // 2 and 3 measured: 0.188 / 0.184 ms against 0.174
#endif
constexpr int kCountUnroll = DTB_COUNT_UNROLL;
constexpr int kCountBins = 8192;  // outcome slots of a level counted in shared memory (32 KB)

// Profiling build (-DDTB_BUILD_CLOCKS): block 0 prints the time between the grid-wide barriers of a few rebuilds.
#ifdef DTB_BUILD_CLOCKS
__device__ int g_build_prints = 0;
#define DTB_STAMP(tag)                                                                   \
  do {                                                                                   \
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_ts < 96) {                              \
      unsigned long long t_;                                                             \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                             \
      ts[n_ts] = t_;                                                                     \
      tg[n_ts++] = (tag);                                                                \
    }                                                                                    \
  } while (0)
#else
#define DTB_STAMP(tag)
#endif

__global__ void __launch_bounds__(kScanBlock) build_tree_coop_kernel(const CoopArgs A) {
  __shared__ uint32_t s_bins[kCountBins];
  cg::grid_group grid = cg::this_grid();
  const TreePtrs &T = A.T;
  const uint32_t nc = A.nc, tid = threadIdx.x;
  const uint64_t n = A.n_ballots;
  const uint64_t gsize = (uint64_t)gridDim.x * kScanBlock, gtid = (uint64_t)blockIdx.x * kScanBlock + tid;
  uint32_t *tab = A.tab;
#ifdef DTB_BUILD_CLOCKS
  unsigned long long ts[96];
  int tg[96], n_ts = 0;
#endif
  DTB_STAMP(0);

  // phase 0: slot sequences of the new ballots (slot_sequences_kernel) and the root (init_root_kernel)
  for (uint64_t b = A.first_new + gtid; b < n; b += gsize) {
    const uint64_t lo = A.off[b];
    const uint32_t len = (uint32_t)(A.off[b + 1] - lo);
    const uint32_t L = min(len, nc - 1u);
    uint8_t path[kMaxCandidates];
    for (uint32_t i = 0; i < nc; ++i) path[i] = (uint8_t)i;
    for (uint32_t d = 0; d < L; ++d) {
      const uint8_t c = A.prefs[lo + d];
      uint32_t i = d;
      while (i < nc - 1u && path[i] != c) ++i;
      A.slots[lo + d] = (uint8_t)(i - d);
      path[i] = path[d];
      path[d] = c;
    }
    for (uint32_t d = L; d < len; ++d) A.slots[lo + d] = 0;
  }
  if (blockIdx.x == 0) {
    for (uint32_t k = tid; k < (uint32_t)kTabWords; k += kScanBlock) tab[k] = 0;
    if (tid <= nc) {
      T.slot_count[tid] = 0;
      T.slot_child[tid] = -1;
      T.slot_cand[tid] = tid < nc ? (uint8_t)tid : 0xFF;
    }
    if (tid < (uint32_t)kMaxCandidates) A.obs_tally0[tid] = 0;
    if (tid < (uint32_t)A.W) T.node_key[tid] = 0;
    __syncthreads();
    if (tid == 0) {
      T.node_base[0] = 0;
      T.node_base[1] = nc + 1;
      tab[1] = 1;
      tab[4 * kTotRow + 0] = 1;
      tab[4 * kTotRow + 1] = nc + 1;
    }
  }
  grid.sync();
  DTB_STAMP(1);

  for (uint32_t depth = 0;; ++depth) {
    const uint32_t K = nc - depth;
    // count_level_kernel: one count per ballot still walking. Near the root all ballots hit a handful of slots and
    // global atomics on them serialise (0.3 of the 0.4 ms of a 200 000-ballot rebuild), so while the level's slots
    // fit, every block counts in shared memory and adds each touched slot once. The loop bound is block-uniform, so
    // whole warps reach the match.
    const uint32_t lvl_base = depth == 0 ? 0u : tab[4 * depth + 2];
    const uint32_t lvl_slots = (depth == 0 ? 1u : tab[4 * depth + 1]) * (K + 1);
    const bool privatise = lvl_slots <= (uint32_t)kCountBins;
    if (privatise) {
      for (uint32_t i = tid; i < lvl_slots; i += kScanBlock) s_bins[i] = 0;
      __syncthreads();
    }
    // kCountUnroll ballots per thread and step: all their loads (cursor -> child -> node base, offsets, slot) are issued
    // before anything is stored, so the dependent-load chains of a thread's ballots overlap instead of following
    // each other.
    for (uint64_t b0 = (uint64_t)blockIdx.x * kScanBlock; b0 < n; b0 += kCountUnroll * gsize) {
      uint32_t idx[kCountUnroll], next_cur[kCountUnroll];
#pragma unroll
      for (int u = 0; u < kCountUnroll; ++u) {
        const uint64_t b = b0 + (uint64_t)u * gsize + tid;
        idx[u] = kDone;
        next_cur[u] = kDone;
        if (b < n) {
          const uint32_t cur = depth == 0 ? 0u : A.cursor[b];
          if (cur != kDone) {
            const uint32_t node = depth == 0 ? 0u : (uint32_t)T.slot_child[cur];
            const uint32_t base = T.node_base[node];
            const uint64_t lo = A.off[b];
            const uint32_t len = (uint32_t)(A.off[b + 1] - lo);
            if (len == depth) {  // irv_node.cpp:191-194
              idx[u] = base + K;
            } else {
              idx[u] = base + A.slots[lo + depth];
              next_cur[u] = K == 2 ? kDone : idx[u];  // irv_node.cpp:210
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kCountUnroll; ++u) {
        const uint64_t b = b0 + (uint64_t)u * gsize + tid;
        if (b < n && (idx[u] != kDone || depth == 0)) A.cursor[b] = next_cur[u];
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, idx[u]);
        if (idx[u] != kDone && (tid & 31) == (unsigned)(__ffs(peers) - 1)) {
          if (privatise) atomicAdd(&s_bins[idx[u] - lvl_base], (uint32_t)__popc(peers));
          else atomicAdd(&T.slot_count[idx[u]], (uint32_t)__popc(peers));
        }
      }
    }
    if (privatise) {
      __syncthreads();
      for (uint32_t i = tid; i < lvl_slots; i += kScanBlock)
        if (s_bins[i]) atomicAdd(&T.slot_count[lvl_base + i], s_bins[i]);
    }
    if (K <= 2 || n == 0) break;  // irv_node.cpp:210: a node with two candidates left has no children
    grid.sync();
    DTB_STAMP(10 + (int)depth);  // count
    const uint32_t *row = tab + 4 * depth;
    const uint32_t level_nodes = row[1];
    if (level_nodes == 0) break;
    const uint32_t n_items = level_nodes * (K + 1);
    const uint32_t n_tiles = (n_items + kScanTile - 1) / kScanTile;
    LevelFlag f{T.slot_count, row, K + 1};
    // (Counting the new children per tile inside the count phase - the thread whose atomic finds a slot at 0 adds one
    // to its tile - would save this phase and its barrier, but at the deep levels nearly every ballot opens a slot and
    // those extra atomics land on a few hundred counters: 0.174 -> 0.189 ms. Measured and dropped.)
    uint32_t *sums = A.block_sums;
    tile_sum_phase(f, n_items, sums);
    grid.sync();
    DTB_STAMP(100 + (int)depth);  // tile sums
    // scan_sums_kernel without its own barrier: there are at most a few hundred tile sums, so every block adds up
    // the ones it needs itself - all of them for the size of the next level, those before a tile for that tile's
    // first rank - and only block 0 writes the level table (read again after the next barrier).
    const uint32_t carry = sum_tiles(sums, n_tiles);
    const uint32_t level_first_node = row[0], first_new_node = row[0] + row[1], n_new = carry;
    const uint32_t new_slot_base = row[2] + row[1] * (K + 1);
    if (blockIdx.x == 0 && tid == 0) {
      uint32_t *tot = tab + 4 * kTotRow, *next = tab + 4 * (depth + 1);
      next[0] = first_new_node;
      next[1] = carry;
      next[2] = new_slot_base;
      tot[0] = first_new_node + carry;
      tot[1] = new_slot_base + carry * K;  // a child has one slot less than its parent
      T.node_base[first_new_node + n_new] = new_slot_base + n_new * K;
    }
    {  // create_children_kernel
      for (uint32_t vb = blockIdx.x; vb < n_tiles; vb += gridDim.x) {
        const uint32_t first = vb * (uint32_t)kScanTile + tid * (uint32_t)kScanItems;
        uint32_t flags = 0, s = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k)
          if (first + k < n_items && f(first + k)) {
            flags |= 1u << k;
            ++s;
          }
        const uint32_t before = sum_tiles(sums, vb);
        uint32_t total;
        uint32_t rank = block_exclusive_scan(s, &total) + before;
        for (int k = 0; k < kScanItems; ++k) {
          if (!(flags >> k & 1u)) continue;
          const uint32_t i = first + k, pn = i / (K + 1), j = i % (K + 1);
          const uint32_t slot = row[2] + i, pbase = row[2] + pn * (K + 1);
          const uint32_t child = first_new_node + rank, cbase = new_slot_base + rank * K;
          ++rank;
          write_child(T, slot, pbase, j, K, child, cbase, level_first_node + pn, depth, A.W);
        }
      }
    }
    grid.sync();
    DTB_STAMP(300 + (int)depth);  // children
  }
  grid.sync();
  DTB_STAMP(400);
  // node totals and the packed observed multiset (node_totals_kernel, tile/scan, write_observed_kernel)
  const uint32_t n_nodes = tab[4 * kTotRow];
  if (gtid < nc) A.obs_tally0[gtid] = T.slot_count[gtid];  // first-preference tallies = the root's slot counts
  for (uint64_t i = gtid; i < n_nodes; i += gsize) {
    const uint32_t base = T.node_base[i], K = T.node_base[i + 1] - base - 1u;
    uint32_t s = 0;
    for (uint32_t j = 0; j < K; ++j) s += T.slot_count[base + j];
    T.node_total[i] = s;
  }
  ObservedFlag of{T.node_base, T.slot_count};
  tile_sum_phase(of, n_nodes, A.block_sums);
  grid.sync();
  DTB_STAMP(401);
  const uint32_t node_tiles = (n_nodes + kScanTile - 1) / kScanTile;
  if (blockIdx.x == 0) {  // the number of observed entries; each block sums the tiles before its own itself
    const uint32_t carry = sum_tiles(A.block_sums, node_tiles);
    if (tid == 0) tab[4 * kTotRow + 2] = carry;
  }
  for (uint32_t vb = blockIdx.x; vb < node_tiles; vb += gridDim.x) {
    const uint32_t first = vb * (uint32_t)kScanTile + tid * (uint32_t)kScanItems;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
      if (first + k < n_nodes) s += of(first + k);
    const uint32_t before = sum_tiles(A.block_sums, vb);
    uint32_t total;
    uint32_t at = block_exclusive_scan(s, &total) + before;
    if (s == 0) continue;
    for (int k = 0; k < kScanItems; ++k) {
      const uint32_t i = first + k;
      if (i >= n_nodes) break;
      const uint32_t base = T.node_base[i], K = T.node_base[i + 1] - base - 1u, depth = nc - K;
      const uint32_t stop = T.slot_count[base + K];
      if (stop) write_observed(T, A.O, at++, i, depth, -1, stop, nc, A.W);
      if (K == 2u)
        for (uint32_t j = 0; j < 2u; ++j) {
          const uint32_t c = T.slot_count[base + j];
          if (c) write_observed(T, A.O, at++, i, depth, (int)T.slot_cand[base + j], c, nc, A.W);
        }
    }
  }
#ifdef DTB_BUILD_CLOCKS
  DTB_STAMP(403);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int k = atomicAdd(&g_build_prints, 1);
    if (k >= 20 && k < 22)
      for (int i = 1; i < n_ts; ++i) printf("build %d phase %d: %llu ns\n", k, tg[i], ts[i] - ts[i - 1]);
  }
#endif
}

#define BUILD_TRY(expr)              \
  do {                               \
    cudaError_t e_ = (expr);         \
    if (e_ != cudaSuccess) return e_; \
  } while (0)

// Co-resident blocks per SM of the cooperative build kernel (0: cooperative launch not available).
int coop_blocks_per_sm() {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0, coop = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, build_tree_coop_kernel, kScanBlock, 0) != cudaSuccess)
      n = 0;
    cached = n;
  }
  return cached;
}

TreePtrs ptrs(DeviceTreeStore &S) {
  return TreePtrs{S.node_base.p, S.node_total.p, S.slot_count.p, S.slot_child.p, S.slot_cand.p, S.node_key.p};
}

cudaError_t grow_nodes(DeviceTreeStore &S, size_t nodes, size_t keep_nodes, int W, cudaStream_t s) {
  BUILD_TRY(S.node_base.grow(nodes + 1, keep_nodes + 1, s));
  BUILD_TRY(S.node_key.grow(nodes * W, keep_nodes * W, s));
  return cudaSuccess;
}

cudaError_t grow_slots(DeviceTreeStore &S, size_t slots, size_t keep, cudaStream_t s) {
  BUILD_TRY(S.slot_count.grow(slots, keep, s));
  BUILD_TRY(S.slot_child.grow(slots, keep, s));
  BUILD_TRY(S.slot_cand.grow(slots, keep, s));
  return cudaSuccess;
}

}  // namespace

// off[i] = i * len for i in [first, last]: the offsets of a log whose ballots all have `len` preferences.
__global__ void uniform_offsets_kernel(uint64_t *off, uint64_t first, uint64_t last, uint32_t len) {
  const uint64_t i = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= last) off[i] = i * len;
}

cudaError_t build_tree_on_device(DeviceTreeStore &S, uint32_t nc, const uint8_t *h_prefs, const uint64_t *h_offsets,
                                 uint64_t n_ballots, cudaStream_t stream, BuildStats *stats, bool *too_big,
                                 uint64_t bound_budget_bytes, bool allow_cooperative, uint32_t uniform_len) {
  const int W = key_words_for(nc);
  BuildStats st;
  if (too_big) *too_big = false;
  if (!S.h_tab) BUILD_TRY(cudaHostAlloc((void **)&S.h_tab, kTabWords * sizeof(uint32_t), cudaHostAllocDefault));
  BUILD_TRY(S.level_tab.reserve(kTabWords));
  uint32_t *tab = S.level_tab.p;
  auto read_table = [&]() -> cudaError_t {  // the one kind of host round trip of a build
    BUILD_TRY(cudaMemcpyAsync(S.h_tab, tab, kTabWords * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    ++st.host_syncs;
    return cudaStreamSynchronize(stream);
  };

  // 1. append the new part of the ballot log and compute its slot sequences
  if (n_ballots < S.log_ballots) S.log_ballots = S.log_bytes = 0;  // the log was reset
  uint64_t first_new = n_ballots;  // cooperative path: the kernel computes the slot sequences of [first_new, n_ballots)
  bool slots_pending = false;
  if (n_ballots > S.log_ballots) {
    const uint64_t old_b = S.log_ballots, old_bytes = S.log_bytes, bytes = h_offsets[n_ballots];
    BUILD_TRY(S.log_prefs.grow(bytes, old_bytes, stream));
    BUILD_TRY(S.log_slots.grow(bytes, old_bytes, stream));
    BUILD_TRY(S.log_off.grow(n_ballots + 1, old_b + 1, stream));
    if (bytes > old_bytes)
      BUILD_TRY(cudaMemcpyAsync(S.log_prefs.p + old_bytes, h_prefs + old_bytes, bytes - old_bytes, cudaMemcpyHostToDevice, stream));
    if (uniform_len != 0xFFFFFFFFu) {
      const uint64_t cnt = n_ballots - old_b + 1;
      uniform_offsets_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, stream>>>(S.log_off.p, old_b, n_ballots, uniform_len);
      ++st.launches;
      st.h2d_bytes += bytes - old_bytes;
    } else {
      BUILD_TRY(cudaMemcpyAsync(S.log_off.p + old_b, h_offsets + old_b, (n_ballots - old_b + 1) * sizeof(uint64_t),
                                cudaMemcpyHostToDevice, stream));
      st.h2d_bytes += (bytes - old_bytes) + (n_ballots - old_b + 1) * sizeof(uint64_t);
    }
    first_new = old_b;
    slots_pending = true;
    S.log_ballots = n_ballots;
    S.log_bytes = bytes;
  }

  // 2. a priori bounds: depth d + 1 has at most min(#ballots, nodes(d) * children) nodes
  uint64_t ub_nodes[kMaxCandidates + 1] = {0};
  ub_nodes[0] = 1;
  uint64_t ub_total = 1, ub_slots = nc + 1;
  for (uint32_t d = 0; nc - d > 2; ++d) {
    ub_nodes[d + 1] = std::min<uint64_t>(n_ballots, ub_nodes[d] * (nc - d));
    ub_total += ub_nodes[d + 1];
    ub_slots += ub_nodes[d + 1] * (nc - d);  // a node of depth d + 1 has nc - d - 1 children + the stop slot
  }
  const uint64_t ub_bytes = ub_total * (8 + 8ull * W) + ub_slots * 9 + n_ballots * (8ull * W + 5);
  const bool bounded = ub_bytes <= bound_budget_bytes && ub_slots < 0xFFFFFFF0ull;

  // 3. the whole build in one cooperative launch when the bounds fit (no host round trip until the end)
  if (bounded && allow_cooperative && coop_blocks_per_sm() > 0) {
    BUILD_TRY(grow_nodes(S, ub_total, 0, W, stream));
    BUILD_TRY(grow_slots(S, ub_slots, 0, stream));
    BUILD_TRY(S.node_total.grow(ub_total, 0, stream));
    BUILD_TRY(S.obs_tally0.reserve(kMaxCandidates));
    BUILD_TRY(S.cursor.grow(std::max<uint64_t>(n_ballots, 1), 0, stream));
    uint64_t max_items = std::max<uint64_t>(ub_total, 1);
    for (uint32_t d = 0; nc - d > 2; ++d) max_items = std::max<uint64_t>(max_items, ub_nodes[d] * (nc - d + 1));
    BUILD_TRY(S.block_sums.grow((max_items + kScanTile - 1) / kScanTile + 1, 0, stream));
    const uint64_t obs_cap = std::max<uint64_t>(n_ballots, 1);  // every entry stands for at least one ballot
    BUILD_TRY(S.obs_key.grow((size_t)obs_cap * W, 0, stream));
    BUILD_TRY(S.obs_cnt.grow(obs_cap, 0, stream));
    BUILD_TRY(S.obs_first.grow(obs_cap, 0, stream));
    CoopArgs A;
    A.T = ptrs(S);
    A.prefs = S.log_prefs.p;
    A.off = S.log_off.p;
    A.slots = S.log_slots.p;
    A.cursor = S.cursor.p;
    A.block_sums = S.block_sums.p;
    A.tab = tab;
    A.obs_tally0 = S.obs_tally0.p;
    A.O = ObsOut{S.obs_key.p, S.obs_cnt.p, S.obs_tally0.p, S.obs_first.p};
    A.first_new = first_new;
    A.n_ballots = n_ballots;
    A.nc = nc;
    A.W = W;
    int dev = 0, sms = 0;
    BUILD_TRY(cudaGetDevice(&dev));
    BUILD_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // One block per SM at most: a grid-wide barrier costs more the more blocks take part, and the phases are short
    // chains of dependent loads rather than bandwidth (measured on the 200 000-ballot benchmark tree: 0.40 ms with
    // 148 blocks, 0.44 with 296, 0.49 with 592; 0.41 as 38 separate launches).
    const uint64_t want = std::max<uint64_t>(1, (std::max<uint64_t>(n_ballots, max_items / 4) + 2 * kScanBlock - 1) / (2 * kScanBlock));
    const int grid = (int)std::min<uint64_t>(want, (uint64_t)std::min(coop_blocks_per_sm(), 1) * sms);
    void *args[] = {&A};
    BUILD_TRY(cudaLaunchCooperativeKernel((const void *)build_tree_coop_kernel, dim3(grid), dim3(kScanBlock), args, 0, stream));
    ++st.launches;
    BUILD_TRY(read_table());
    S.n_nodes = S.h_tab[4 * kTotRow];
    S.n_slots = S.h_tab[4 * kTotRow + 1];
    S.n_obs = S.h_tab[4 * kTotRow + 2];
    if (stats) *stats = st;
    return cudaSuccess;
  }
  if (slots_pending) {
    const uint64_t fresh = n_ballots - first_new;
    const int grid = (int)std::min<uint64_t>((fresh + 127) / 128, 1u << 16);
    slot_sequences_kernel<<<grid, 128, 0, stream>>>(S.log_prefs.p, S.log_off.p, S.log_slots.p, first_new, n_ballots, nc);
    ++st.launches;
  }

  // 3b. root (multi-launch path)
  if (bounded) {
    BUILD_TRY(grow_nodes(S, ub_total, 0, W, stream));
    BUILD_TRY(grow_slots(S, ub_slots, 0, stream));
    BUILD_TRY(S.node_total.grow(ub_total, 0, stream));
  } else {
    size_t guess = std::max<size_t>({(size_t)S.n_nodes + S.n_nodes / 4, (size_t)std::min<uint64_t>(n_ballots, 1u << 24), 64});
    BUILD_TRY(grow_nodes(S, guess, 0, W, stream));
    BUILD_TRY(grow_slots(S, std::max<size_t>((size_t)S.n_slots + S.n_slots / 4, guess * 4 + nc + 1), 0, stream));
  }
  BUILD_TRY(S.obs_tally0.reserve(kMaxCandidates));
  BUILD_TRY(S.cursor.grow(std::max<uint64_t>(n_ballots, 1), 0, stream));
  init_root_kernel<<<1, 64, 0, stream>>>(ptrs(S), S.obs_tally0.p, tab, nc, W);
  ++st.launches;

  // 4. one level at a time
  uint32_t n_nodes = 1, n_slots = nc + 1;  // exact path: what the arrays hold so far
  const int ballot_grid = (int)((n_ballots + 255) / 256);
  for (uint32_t depth = 0;; ++depth) {
    const uint32_t K = nc - depth;
    if (n_ballots > 0) {
      count_level_kernel<<<ballot_grid, 256, 0, stream>>>(ptrs(S), S.log_off.p, S.log_slots.p, S.cursor.p, n_ballots, nc, depth);
      ++st.launches;
    }
    if (K <= 2 || n_ballots == 0) break;  // irv_node.cpp:210: a node with two candidates left has no children
    uint32_t level_nodes;
    if (bounded) {
      level_nodes = (uint32_t)ub_nodes[depth];
    } else {
      BUILD_TRY(read_table());
      level_nodes = S.h_tab[4 * depth + 1];
    }
    if (level_nodes == 0) break;
    const uint32_t nb = (uint32_t)(((uint64_t)level_nodes * (K + 1) + kScanTile - 1) / kScanTile);
    BUILD_TRY(S.block_sums.grow(nb, 0, stream));
    LevelFlag f{S.slot_count.p, tab + 4 * depth, K + 1};
    tile_sums_kernel<<<nb, kScanBlock, 0, stream>>>(f, tab + 4 * depth + 1, K + 1, S.block_sums.p);
    scan_sums_kernel<<<1, kScanBlock, 0, stream>>>(S.block_sums.p, tab + 4 * depth + 1, K + 1, tab, (int)depth, K + 1);
    st.launches += 2;
    if (!bounded) {
      BUILD_TRY(read_table());
      const uint32_t n_new = S.h_tab[4 * (depth + 1) + 1];
      if (n_new == 0) break;
      const uint64_t want_slots = (uint64_t)n_slots + (uint64_t)n_new * K;
      if (want_slots >= 0xFFFFFFF0ull) {
        if (too_big) *too_big = true;
        return cudaSuccess;
      }
      BUILD_TRY(grow_nodes(S, (size_t)n_nodes + n_new, n_nodes, W, stream));
      BUILD_TRY(grow_slots(S, want_slots, n_slots, stream));
      n_nodes += n_new;
      n_slots = (uint32_t)want_slots;
      f.slot_count = S.slot_count.p;
    }
    create_children_kernel<<<nb, kScanBlock, 0, stream>>>(ptrs(S), f, tab, S.block_sums.p, depth, W);
    ++st.launches;
  }

  // 5. per-node totals and the observed multiset
  uint32_t nodes_launch = (uint32_t)std::min<uint64_t>(ub_total, 0xFFFFFFFFull);
  if (!bounded) {
    BUILD_TRY(read_table());
    nodes_launch = S.h_tab[4 * kTotRow];
    BUILD_TRY(S.node_total.grow(nodes_launch, 0, stream));
  }
  node_totals_kernel<<<(nodes_launch + 255) / 256, 256, 0, stream>>>(ptrs(S), tab, S.obs_tally0.p, nc);
  const uint32_t nb = (nodes_launch + kScanTile - 1) / kScanTile;
  BUILD_TRY(S.block_sums.grow(nb, 0, stream));
  ObservedFlag of{S.node_base.p, S.slot_count.p};
  tile_sums_kernel<<<nb, kScanBlock, 0, stream>>>(of, tab + 4 * kTotRow, 1u, S.block_sums.p);
  scan_sums_kernel<<<1, kScanBlock, 0, stream>>>(S.block_sums.p, tab + 4 * kTotRow, 1u, tab, -1, 0u);
  st.launches += 3;
  uint64_t obs_cap = std::max<uint64_t>(n_ballots, 1);  // every entry stands for at least one ballot
  if (!bounded) {
    BUILD_TRY(read_table());
    obs_cap = std::max<uint32_t>(S.h_tab[4 * kTotRow + 2], 1u);
  }
  BUILD_TRY(S.obs_key.grow((size_t)obs_cap * W, 0, stream));
  BUILD_TRY(S.obs_cnt.grow(obs_cap, 0, stream));
  BUILD_TRY(S.obs_first.grow(obs_cap, 0, stream));
  ObsOut O{S.obs_key.p, S.obs_cnt.p, S.obs_tally0.p, S.obs_first.p};
  write_observed_kernel<<<nb, kScanBlock, 0, stream>>>(ptrs(S), of, tab, S.block_sums.p, O, nc, W);
  ++st.launches;
  BUILD_TRY(cudaGetLastError());
  BUILD_TRY(read_table());
  S.n_nodes = S.h_tab[4 * kTotRow];
  S.n_slots = S.h_tab[4 * kTotRow + 1];
  S.n_obs = S.h_tab[4 * kTotRow + 2];
  if (stats) *stats = st;
  return cudaSuccess;
}

}  // namespace dtb
