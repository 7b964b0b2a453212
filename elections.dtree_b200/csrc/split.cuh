// split.cuh — pieces shared by the two simulation kernels (simulate.cuh: materialise every ballot;
// deferred.cuh: expand only what IRV asks for): the frontier item, kernel arguments, the per-warp
// shared-memory tile and the Dirichlet-multinomial split of a chunk of items.
//
// Every draw made for a node is a pure function of (seed, election, node key, node parameters,
// count), never of which kernel, warp or chunk evaluates it. That is what makes the two kernels
// produce bit-identical elections.
#pragma once
#include <cstdint>

#include "dtree_types.h"
#include "rng.cuh"
#include "variates.cuh"

namespace dtb {

template <int W>
struct alignas(16) Item {
  BallotKey<W> prefix;  // candidates chosen so far
  uint64_t node_key;    // RNG identity of the node (hash chain over the prefix)
  uint32_t count;       // ballots to distribute below this node
  int32_t node;         // CSR node id, -1 = uninitialised (prior-only) subtree
  uint32_t used;        // bit c set <=> candidate c is in the prefix
  uint32_t pad_;
};

enum StatSlot {
  kStatEntries = 0, kStatItems, kStatGammas, kStatBinomials, kStatRereads, kStatSlots, kStatUrnItems, kStatCount
};

struct SimArgs {
  DeviceParams P;
  DeviceTree T;
  // observed ballots (replace == false): shared by all elections
  const void *obs_key;        // BallotKey<W>[n_obs]
  const uint32_t *obs_cnt;
  const uint8_t *obs_first;   // first preference, 0xFF for a blank ballot
  const uint32_t *obs_tally0; // [nc] first-preference tallies of the observed ballots
  uint32_t n_obs;
  // job
  uint64_t seed;
  uint64_t first_election;
  uint32_t n_elections;
  uint32_t n_sample;          // ballots to draw per election (N, or N - n_observed)
  uint32_t n_winners;
  uint32_t urn_max;
  int run_irv;                // 0: expansion only (predictive / posterior_set dumps)
  int sc_function;            // social choice function run on each election: 0 IRV, 1 plurality (lane_kernel only)
  int use_obs;                // observed ballots are part of the election (replace == false)
  // per-block scratch
  uint8_t *scratch;
  uint64_t scratch_stride;    // bytes per block
  uint32_t cap_items;         // frontier capacity (items) per buffer
  uint32_t cap_small;         // simulate_kernel: capacity of an election's small-item queue
  uint32_t cap_entries;
  uint32_t cap_queue, cap_big; // deferred kernel: circular work queue / big-count lists (items)
  // deferred kernel: a few worst-case arenas for the elections that outgrow their warp's scratch
  uint8_t *arena_scratch;
  uint64_t arena_stride;
  uint32_t n_arenas, arena_cap_items, arena_cap_queue, arena_cap_big;
  unsigned int *arena_locks;   // [n_arenas], 0 = free
  unsigned int *arena_runs;    // number of elections re-run in an arena
  uint32_t *usage;             // [0..2] high-water marks: parked pool, work queue, big list; [4] lane_kernel elections handed to the replay
  uint32_t max_warps;          // warps with a larger grid-wide index stay idle (pilot launches)
  // replaying a list of elections (lane_kernel's overflow list): indices relative to first_election
  const uint32_t *election_list;   // NULL: elections 0..n_elections-1
  const uint32_t *n_elections_ptr; // NULL, or device count that caps n_elections (known only on the device)
  uint32_t *retry_list;            // lane_kernel: elections whose arena overflowed
  unsigned int *retry_count;
  uint32_t retry_cap;
  uint32_t tau_quarters;           // deferred_kernel: share (in quarters) of the missing margin that may keep waiting (1..3)
  uint32_t lane_sync;              // lane_kernel: 0 warps run free; 1 the block claims together and meets after the root split; 2 and at every round; 3 and at every split step
  uint32_t lane_width;             // lane_kernel: elections a warp claims at a time (1..32 lanes of it work)
  uint32_t lane_max_splits;        // lane_kernel: an election that needs more node splits goes to the retry list (0: no limit)
  // outputs
  unsigned long long *win_counts;  // [nc], added to
  uint8_t *elim_orders;            // NULL or [n_elections * nc]
  uint32_t *entry_count_out;       // dump mode: number of entries of the (single) election
  uint32_t *tie_rounds_out;        // NULL or [n_elections]
  unsigned int *work_counter;      // next election index to hand out
  int *error_flag;                 // set to 1 on scratch overflow
  unsigned long long *stats;       // [kStatCount]
};

// Chunk geometry. An item's d outcomes are padded to dp = 2^log_dp leaves of a complete binary
// tree stored heap-style in 2*dp words (+1 word of padding against bank conflicts). dp is chosen
// per depth (d shrinks as the ballot grows), and a warp takes as many items per chunk as fit its
// tile, at most kMaxItems.
template <int W>
struct Chunk {
  static constexpr int kTileWords = 32 * 33;
  static constexpr int kMaxItems = W == 1 ? 32 : 16;
};

template <int W>
struct WarpTile {
  uint32_t val[Chunk<W>::kTileWords];     // float sums or uint32 counts, heap-indexed per item
  Item<W> item[Chunk<W>::kMaxItems];
  uint32_t nbase[Chunk<W>::kMaxItems];    // first CSR slot of the item's node, UINT32_MAX for a prior-only item
};

// Adds `amount` to s_tally[cand] for every lane with valid; lanes naming the same candidate are
// combined first so that a warp issues one shared-memory atomic per distinct candidate.
__device__ __forceinline__ void warp_tally_add(uint32_t *s_tally, bool valid, uint32_t cand, uint32_t amount) {
  uint32_t pending = __ballot_sync(0xffffffffu, valid);
  while (pending) {
    int leader = __ffs(pending) - 1;
    uint32_t c0 = __shfl_sync(0xffffffffu, cand, leader);
    uint32_t same = __ballot_sync(0xffffffffu, valid && cand == c0);
    uint32_t sum = __reduce_add_sync(0xffffffffu, (valid && cand == c0) ? amount : 0u);
    if ((int)(threadIdx.x & 31) == leader) atomicAdd(&s_tally[c0], sum);
    pending &= ~same;
  }
}

// i-th (0-based) candidate not in `used`, ascending.
__device__ __forceinline__ uint32_t nth_unused(uint32_t used, uint32_t nc, uint32_t i) {
  uint32_t v = ~used & (nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u));
  // branch-free select of the i-th set bit by halving (the __fns intrinsic is a slow software loop)
  uint32_t r = 0, c;
  c = (uint32_t)__popc(v & 0xFFFFu); if (i >= c) { r += 16u; i -= c; v >>= 16; }
  c = (uint32_t)__popc(v & 0xFFu);   if (i >= c) { r += 8u;  i -= c; v >>= 8; }
  c = (uint32_t)__popc(v & 0xFu);    if (i >= c) { r += 4u;  i -= c; v >>= 4; }
  c = (uint32_t)__popc(v & 0x3u);    if (i >= c) { r += 2u;  i -= c; v >>= 2; }
  c = v & 1u;                        if (i >= c) { r += 1u; }
  return r;
}

// First standing preference of a ballot: scans from the front, skipping eliminated candidates.
template <int W>
__device__ __forceinline__ uint32_t first_standing(const BallotKey<W> &k, uint32_t len, uint32_t eliminated) {
  for (uint32_t j = 0; j < len; ++j) {
    uint32_t c = key_get<W>(k, (int)j);
    if (!(eliminated >> c & 1u)) return c;
  }
  return 0xFFu;
}

struct LevelInfo {
  uint32_t depth, K, d;
  bool T, leaf_children;
  float a_prior;
};

__device__ __forceinline__ LevelInfo level_info(const DeviceParams &P, uint32_t depth) {
  LevelInfo lv;
  lv.depth = depth;
  lv.K = P.nc - depth;
  lv.T = depth >= P.min_depth;
  lv.d = lv.K + (lv.T ? 1u : 0u);
  lv.a_prior = P.a_prior[depth];
  // irv_node.cpp:136 (depth == maxDepth-1) and :46 (child depth == nc-1 or == maxDepth)
  lv.leaf_children = (depth == P.leaf_depth) || (depth + 2 == P.nc);
  return lv;
}

// Warp-wide reservation of consecutive slots behind a shared cursor. Returns the slot of this lane
// (meaningful for valid lanes only); all 32 lanes must call.
__device__ __forceinline__ uint32_t warp_reserve(uint32_t *cursor, bool valid, int lane) {
  uint32_t m = __ballot_sync(0xffffffffu, valid);
  if (m == 0u) return 0u;
  uint32_t base = 0;
  const int leader = __ffs(m) - 1;
  if (lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}

// Same for lanes that arrive divergently (the walkers): whoever is here right now forms a group.
__device__ __forceinline__ uint32_t coalesced_reserve(uint32_t *cursor) {
  const uint32_t m = __activemask();
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}

// Polya urn at one node for `count` <= 8 ballots: ball b copies a uniformly chosen earlier ball with
// probability b / (A + b), else draws a fresh outcome with probability alpha_i / A. Exactly
// Dirichlet-multinomial(count, alpha); no Gamma or binomial variates. Returns the outcome of
// ball b in bits [8b, 8b+8).
static __device__ __noinline__ uint64_t urn_picks(const SimArgs &a, const LevelInfo &lv, int32_t node, uint32_t nbase,
                                              uint32_t count, const RngKey key, uint64_t node_key) {
  // A = (observed ballots below the node) + d a_prior. A fresh outcome is drawn as a mixture: with probability
  // tot / A one of the observed ballots is copied (exact integer walk over the slot counts), else the prior picks
  // one of the d outcomes uniformly. Sums are formed in double and the uniform has 32 bits, so a prior-only
  // outcome next to slots holding millions of ballots keeps its probability a_prior / A.
  const uint32_t d = lv.d;
  uint32_t tot = 0;
  if (nbase != 0xFFFFFFFFu) {
    tot = a.T.node_total[node];
    if (lv.T) tot += a.T.slot_count[nbase + lv.K];
  }
  const double prior = (double)lv.a_prior, seen = (double)tot;
  const double A = __fma_rn((double)d, prior, seen);
  uint64_t picks = 0;
  U4 r;
#pragma unroll 1
  for (uint32_t b = 0; b < count; ++b) {
    if ((b & 3u) == 0u) r = node_random(key, node_key, kStreamUrn, b >> 2);
    const uint32_t bits = (b & 3u) == 0u ? r.x : (b & 3u) == 1u ? r.y : (b & 3u) == 2u ? r.z : r.w;
    uint32_t pick;
    if (!(A > 0.0)) {
      // all parameters 0 (distributions.cpp:69-77): the first ball is uniform, the rest follow it
      pick = b == 0u ? uniform_below(bits, d) : (uint32_t)(picks & 0xFFu);
    } else {
      const double u = __dmul_rn(u01d(bits), __dadd_rn(A, (double)b));
      if (u >= A) {  // copies an earlier ball
        uint32_t j = (uint32_t)__dadd_rn(u, -A);
        if (j >= b) j = b - 1u;
        pick = (uint32_t)(picks >> (8u * j)) & 0xFFu;
      } else if (u < seen) {  // copies an observed ballot: the slot holding the floor(u)-th of them
        const uint32_t target = (uint32_t)u;
        uint32_t acc = 0;
        pick = d - 1u;
#pragma unroll 1
        for (uint32_t i = 0; i < d; ++i) {
          acc += a.T.slot_count[nbase + i];
          if (target < acc) {
            pick = i;
            break;
          }
        }
      } else {  // the prior: uniform over the d outcomes
        pick = (uint32_t)__ddiv_rn(__dadd_rn(u, -seen), prior);
        if (pick >= d) pick = d - 1u;
      }
    }
    picks |= (uint64_t)pick << (8u * b);
  }
  return picks;
}


__device__ __forceinline__ int tree_log2(uint32_t d) { return d <= 2u ? 1 : 32 - __clz(d - 1u); }

// Dirichlet-multinomial split of the m items staged in wt.item / wt.nbase (depth of item i in
// wt.item[i].pad_ & 0xFF). On return the number of ballots that outcome `slot` of item `it` received
// is wt.val[it * stride + dp_it + slot], with stride = 2 * 2^log_dp_max + 1 and dp_it = 2^tree_log2(d_it).
//
//  count >  urn_max: Gamma variates, one lane per (item, outcome), into the leaves of the item's
//                    complete binary tree; partial sums bottom-up; then the count is pushed down with one
//                    binomial per internal node, one lane per (item, tree node). This is rMultinomial
//                    (distributions.cpp:21-53) with the conditional binomials arranged as a balanced tree
//                    instead of a chain: the same multinomial law, log2(d) dependent steps instead of d-1,
//                    and no 1-x cancellation because both child sums are at hand.
//  count <= urn_max: Polya urn, one lane per item.
template <int W, bool LOG>
__device__ void split_items(const SimArgs &a, WarpTile<W> &wt, uint32_t m, int log_dp_max, const RngKey key,
                            uint32_t *stat) {
  const int lane = threadIdx.x & 31;
  const uint32_t dp_max = 1u << log_dp_max, stride = 2u * dp_max + 1u;

  // ---- (a) Gamma variates into the leaves ------------------------------------------------------
  for (uint32_t p = lane; p < (m << log_dp_max); p += 32) {
    const uint32_t it = p >> log_dp_max, slot = p & (dp_max - 1u);
    const Item<W> &x = wt.item[it];
    const LevelInfo lv = level_info(a.P, x.pad_ & 0xFFu);
    const uint32_t dp = 1u << tree_log2(lv.d);
    if (slot >= dp) continue;
    float g = LOG ? -INFINITY : 0.0f;
    if (slot < lv.d && x.count > a.urn_max) {
      const uint32_t nb = wt.nbase[it];
      const uint32_t seen = nb != 0xFFFFFFFFu ? a.T.slot_count[nb + slot] : 0u;
      g = outcome_gamma<LOG>(seen, lv.a_prior, key, x.node_key, slot, stat[kStatGammas]);  // Gamma(0) is the point mass at 0
    }
    wt.val[it * stride + dp + slot] = __float_as_uint(g);
  }
  __syncwarp();
  if (LOG) {
    // Variates are logarithms here: subtract the item's maximum and exponentiate. If every
    // parameter is 0 the Dirichlet draw is a uniform one-hot (distributions.cpp:69-77).
    for (uint32_t it = lane; it < m; it += 32) {
      const Item<W> &x = wt.item[it];
      if (x.count <= a.urn_max) continue;
      const uint32_t d = level_info(a.P, x.pad_ & 0xFFu).d, dp = 1u << tree_log2(d);
      uint32_t *row = wt.val + it * stride + dp;
      float mx = -INFINITY;
      for (uint32_t i = 0; i < d; ++i) mx = fmaxf(mx, __uint_as_float(row[i]));
      if (mx == -INFINITY) {
        U4 r = node_random(key, x.node_key, kStreamOneHot, 0u);
        uint32_t pick = uniform_below(r.x, d);
        for (uint32_t i = 0; i < dp; ++i) row[i] = __float_as_uint(i == pick ? 1.0f : 0.0f);
      } else {
        for (uint32_t i = 0; i < dp; ++i)
          row[i] = __float_as_uint(i < d ? __expf(__fadd_rn(__uint_as_float(row[i]), -mx)) : 0.0f);
      }
    }
    __syncwarp();
  }
  // ---- partial sums, bottom-up (an item with a narrower tree joins at its own top level) ---------
#pragma unroll 1
  for (int lg = log_dp_max - 1; lg >= 0; --lg) {
    const uint32_t w = 1u << lg;
    for (uint32_t p = lane; p < (m << lg); p += 32) {
      const uint32_t it = p >> lg;
      if (lg >= tree_log2(level_info(a.P, wt.item[it].pad_ & 0xFFu).d)) continue;
      uint32_t *row = wt.val + it * stride;
      const uint32_t k = w + (p & (w - 1u));
      row[k] = __float_as_uint(__fadd_rn(__uint_as_float(row[2 * k]), __uint_as_float(row[2 * k + 1])));
    }
    __syncwarp();
  }
  for (uint32_t it = lane; it < m; it += 32)
    wt.val[it * stride + 1] = wt.item[it].count > a.urn_max ? wt.item[it].count : 0u;
  __syncwarp();
  // ---- (b) counts, top-down: node k hands Binomial(n_k, S_left / (S_left + S_right)) to its left child
#pragma unroll 1
  for (int lg = 0; lg < log_dp_max; ++lg) {
    const uint32_t w = 1u << lg;
    for (uint32_t p = lane; p < (m << lg); p += 32) {
      const uint32_t it = p >> lg;
      if (lg >= tree_log2(level_info(a.P, wt.item[it].pad_ & 0xFFu).d)) continue;
      uint32_t *row = wt.val + it * stride;
      const uint32_t k = w + (p & (w - 1u));
      const uint32_t n = row[k];
      const float sl = __uint_as_float(row[2 * k]), sr = __uint_as_float(row[2 * k + 1]);
      uint32_t left = 0;
      if (n != 0u) {
        if (!(sr > 0.0f)) {
          left = n;
        } else if (sl > 0.0f) {
          // draw on the smaller side so that neither tail loses precision
          const float tot = __fadd_rn(sl, sr);
          const bool flip = sl > sr;
          uint32_t att = 0;
          const uint32_t kk = binomial_small_p(n, __fdividef(flip ? sr : sl, tot), key, wt.item[it].node_key, k, att, a.P.binv_max);
          left = flip ? n - kk : kk;
          stat[kStatBinomials] += att;
        }
      }
      row[2 * k] = left;
      row[2 * k + 1] = n - left;
    }
    __syncwarp();
  }
  // ---- small counts: Polya urn, one lane per item ---------------------------------------------------
  for (uint32_t it = lane; it < m; it += 32) {
    const Item<W> &x = wt.item[it];
    if (x.count == 0u || x.count > a.urn_max) continue;
    const LevelInfo lv = level_info(a.P, x.pad_ & 0xFFu);
    uint32_t *row = wt.val + it * stride + (1u << tree_log2(lv.d));
    const uint64_t picks = urn_picks(a, lv, x.node, wt.nbase[it], x.count, key, x.node_key);
    for (uint32_t b = 0; b < x.count; ++b) row[(uint32_t)(picks >> (8u * b)) & 0xFFu] += 1u;
    stat[kStatUrnItems] += 1;
  }
  __syncwarp();
}

}  // namespace dtb
