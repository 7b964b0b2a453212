// deferred.cuh — sample_posterior without materialising the election.
//
// IRV only ever asks one question of a ballot: who is its first standing preference. All the
// ballots below a tree node share that answer until every candidate of the node's prefix has been
// eliminated, so a node's Dirichlet-multinomial split is needed only at that moment. This kernel
// draws it then and not before (principle of deferred decisions): the election starts as the root's
// split; every child whose candidate still stands is parked under that candidate with its ballots
// counted in the candidate's tally; when IRV eliminates a candidate, only the nodes parked under it
// are split further, cascading through children whose candidate is already out.
//
// Because every split is a pure function of (seed, election, node key, node parameters, count)
// (split.cuh), the nodes expanded here receive exactly the draws simulate_kernel would give them,
// the tallies seen by every elimination round are the same integers, and so are the tie-breaks:
// elimination orders and win counts are bit-identical to the materialising kernel's
// (tests/test_gpu_parity.py::test_deferred_equals_materialised). The nodes never asked about —
// typically everything below the two finalists' first preferences, i.e. most of the tree — are never
// drawn, which removes most of the work of IRVNode::sample (irv_node.cpp:107-174) and all of the
// list traffic of socialChoiceIRV (irv_ballot.cpp:42-138).
//
// Observed ballots (replace == false) are not copied per election (dirichlet_tree.h:229-232): a node's
// observed counts ARE its slot counts, so they ride along with the sampled counts of each child.
//
// One warp per election; a block is a bundle of independent warps (no block barrier in the loop).
#pragma once
#include <cstdint>

#include "split.cuh"

namespace dtb {

typedef Item<1> DItem;  // prefix unused; pad_ = depth | holder << 8

struct DeferredWarp {
  uint32_t tally[kMaxCandidates];
  uint32_t n_pending, n_out, overflow, pad;
};

// Splits the nodes in `in` and routes their children; repeats on the children whose candidate is
// already eliminated until none are left. Returns with every live node parked in `pend`.
template <bool LOG>
__device__ void deferred_cascade(const SimArgs &a, WarpTile<1> &wt, DeferredWarp &ws, DItem *in, DItem *out,
                                 uint32_t n_in, DItem *pend, const uint32_t eliminated, const int log_dp_max,
                                 const RngKey key, uint32_t *stat) {
  const int lane = threadIdx.x & 31;
  const uint32_t nc = a.P.nc, cap = a.cap_items;
  const uint32_t dp_max = 1u << log_dp_max, stride = 2u * dp_max + 1u;
  const uint32_t ci = min((uint32_t)Chunk<1>::kMaxItems, (uint32_t)Chunk<1>::kTileWords / stride);
  while (n_in > 0) {
    if (lane == 0) ws.n_out = 0;
    __syncwarp();
    for (uint32_t start = 0; start < n_in; start += ci) {
      const uint32_t m = min(ci, n_in - start);
      for (uint32_t it = lane; it < m; it += 32) {
        DItem x = in[start + it];
        wt.item[it] = x;
        wt.nbase[it] = x.node >= 0 ? a.T.node_base[x.node] : 0xFFFFFFFFu;
        stat[kStatItems] += 1;
      }
      __syncwarp();
      split_items<1, LOG>(a, wt, m, log_dp_max, key, stat);
      for (uint32_t base = 0; base < (m << log_dp_max); base += 32) {
        const uint32_t p = base + lane;
        const uint32_t it = p >> log_dp_max, slot = p & (dp_max - 1u);
        bool live = false, standing = false, cont = false;
        uint32_t tot = 0, cand = 0, cs = 0, depth = 0;
        int32_t child = -1;
        if (it < m) {
          const DItem &x = wt.item[it];
          depth = x.pad_ & 0xFFu;
          const LevelInfo lv = level_info(a.P, depth);
          if (slot < lv.K) {  // the stop outcome (slot K) exhausts its ballots: nothing to route
            const uint32_t nb = wt.nbase[it];
            const uint32_t c = x.count ? wt.val[it * stride + (1u << tree_log2(lv.d)) + slot] : 0u;
            uint32_t obs = 0;
            if (nb != 0xFFFFFFFFu) {
              if (a.use_obs) obs = a.T.slot_count[nb + slot];
              cand = a.T.slot_cand[nb + slot];
              child = a.T.slot_child[nb + slot];
            } else {
              cand = nth_unused(x.used, nc, slot);
            }
            tot = c + obs;
            cs = lv.leaf_children ? 0u : c;  // sampled ballots end with this preference at the last level
            cont = cs > 0u || (a.use_obs && child >= 0 && obs > 0u);
            live = tot > 0u;
            standing = !(eliminated >> cand & 1u);
          }
        }
        warp_tally_add(ws.tally, live && standing, cand, tot);
        const uint32_t po = warp_reserve(&ws.n_pending, live && standing && cont, lane);
        const uint32_t fo = warp_reserve(&ws.n_out, live && !standing && cont, lane);
        if (!(live && cont)) continue;
        const DItem &x = wt.item[it];
        DItem ch;
        ch.prefix.w[0] = 0;
        ch.node_key = child_node_key(x.node_key, cand);
        ch.count = cs;
        ch.node = child;
        ch.used = x.used | (1u << cand);
        ch.pad_ = (depth + 1u) | (cand << 8);
        const uint32_t pos = standing ? po : fo;
        if (pos < cap) {
          (standing ? pend : out)[pos] = ch;
        } else {
          ws.overflow = 1;
        }
      }
      __syncwarp();
    }
    __syncwarp();
    n_in = min(ws.n_out, cap);
    DItem *t = in;
    in = out;
    out = t;
  }
}

template <int BLOCK, bool LOG>
__global__ void __launch_bounds__(BLOCK) deferred_kernel(const __grid_constant__ SimArgs a) {
  __shared__ WarpTile<1> tiles[BLOCK / 32];
  __shared__ DeferredWarp wstate[BLOCK / 32];
  __shared__ uint32_t s_wins[kMaxCandidates];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t nc = a.P.nc, cap = a.cap_items;
  const uint32_t cand_mask = nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u);
  WarpTile<1> &wt = tiles[warp];
  DeferredWarp &ws = wstate[warp];
  DItem *pend = (DItem *)(a.scratch + ((uint64_t)blockIdx.x * (BLOCK / 32) + warp) * a.scratch_stride);
  DItem *fa = pend + cap, *fb = fa + cap;
  const int log_dp_max = tree_log2(level_info(a.P, 0).d);
  uint32_t stat[kStatCount];
#pragma unroll
  for (int i = 0; i < kStatCount; ++i) stat[i] = 0;
  if (tid < (int)kMaxCandidates) s_wins[tid] = 0;
  __syncthreads();
  const uint32_t n_rounds = a.elim_orders ? nc - 1 : nc - a.n_winners;

  for (;;) {
    uint32_t e = 0;
    if (lane == 0) e = atomicAdd(a.work_counter, 1u);
    e = __shfl_sync(0xffffffffu, e, 0);
    if (e >= a.n_elections) break;
    const RngKey key = election_key(a.seed, a.first_election + e);

    ws.tally[lane] = 0;
    if (lane == 0) {
      ws.n_pending = 0;
      ws.overflow = 0;
      DItem root;
      root.prefix.w[0] = 0;
      root.node_key = kRootNodeKey;
      root.count = a.n_sample;
      root.node = 0;
      root.used = 0;
      root.pad_ = 0;
      fa[0] = root;
    }
    __syncwarp();
    uint32_t eliminated = 0, tie_rounds = 0, my_order = 0xFFu;
    deferred_cascade<LOG>(a, wt, ws, fa, fb, 1u, pend, eliminated, log_dp_max, key, stat);

    for (uint32_t round = 0; round < n_rounds; ++round) {
      // candidates with the minimal tally among those standing (irv_ballot.cpp:88-98), one lane each
      const bool standing = lane < (int)nc && !(eliminated >> lane & 1u);
      const uint32_t t = standing ? ws.tally[lane] : 0xFFFFFFFFu;
      const uint32_t mn = __reduce_min_sync(0xffffffffu, t);
      uint32_t tied = __ballot_sync(0xffffffffu, standing && t == mn);
      const uint32_t ntied = (uint32_t)__popc(tied);
      if (ntied > 1u) {  // uniform among the tied, ascending index (irv_ballot.cpp:100-102)
        U4 r = node_random(key, 0ull, kStreamTie, round);
        const uint32_t pick = uniform_below(r.x, ntied);
        for (uint32_t k = 0; k < pick; ++k) tied &= tied - 1u;
        tie_rounds += 1;
      }
      const uint32_t chosen = (uint32_t)__ffs(tied) - 1u;
      if (lane == (int)round) my_order = chosen;
      eliminated |= 1u << chosen;
      // pull the nodes parked under the eliminated candidate; keep the rest, compacted in place
      const uint32_t np = min(ws.n_pending, cap);
      uint32_t kept = 0, n_in = 0;
      for (uint32_t base = 0; base < np; base += 32) {
        const uint32_t i = base + lane;
        DItem x;
        bool mine = false, keep = false;
        if (i < np) {
          x = pend[i];
          mine = ((x.pad_ >> 8) & 0xFFu) == chosen;
          keep = !mine;
        }
        __syncwarp();  // every lane has read its slot before any lane overwrites one
        const uint32_t mm = __ballot_sync(0xffffffffu, mine), mk = __ballot_sync(0xffffffffu, keep);
        const uint32_t lt = (1u << lane) - 1u;
        if (mine) fa[n_in + __popc(mm & lt)] = x;
        if (keep) pend[kept + __popc(mk & lt)] = x;
        n_in += __popc(mm);
        kept += __popc(mk);
      }
      if (lane == 0) ws.n_pending = kept;
      __syncwarp();
      if (n_in) deferred_cascade<LOG>(a, wt, ws, fa, fb, n_in, pend, eliminated, log_dp_max, key, stat);
    }

    if (ws.overflow) {
      if (lane == 0) *a.error_flag = 2;
    }
    const uint32_t standing_mask = ~eliminated & cand_mask;
    if (a.elim_orders) {
      uint8_t *o = a.elim_orders + (uint64_t)e * nc;
      if (lane + 1 < (int)nc) o[lane] = (uint8_t)my_order;
      if (lane == 0) o[nc - 1] = (uint8_t)(__ffs(standing_mask) - 1);
      // winners = last n_winners of the order
      const uint32_t w = lane + 1 < (int)nc ? my_order : (uint32_t)(__ffs(standing_mask) - 1);
      if (lane < (int)nc && lane >= (int)(nc - a.n_winners)) atomicAdd(&s_wins[w], 1u);
    } else {
      if (lane < (int)nc && (standing_mask >> lane & 1u)) atomicAdd(&s_wins[lane], 1u);
    }
    if (a.tie_rounds_out && lane == 0) a.tie_rounds_out[e] = tie_rounds;
    __syncwarp();
  }
  __syncthreads();
  if (tid < (int)nc && s_wins[tid]) atomicAdd(&a.win_counts[tid], (unsigned long long)s_wins[tid]);
  if (a.stats) {
#pragma unroll
    for (int i = 0; i < kStatCount; ++i) {
      uint32_t s = __reduce_add_sync(0xffffffffu, stat[i]);
      if (lane == 0 && s) atomicAdd(&a.stats[i], (unsigned long long)s);
    }
  }
}

}  // namespace dtb
