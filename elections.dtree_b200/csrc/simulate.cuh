// simulate.cuh — the posterior-simulation kernel. One thread block simulates one election at a
// time (persistent blocks pull election ids from a counter):
//
//   (a)+(b) level-synchronous expansion of the Dirichlet-tree (IRVNode::sample, irv_node.cpp:107-174,
//           and lazyIRVBallots, :27-84). The frontier of one depth lives in a two-ended queue in
//           global memory: items whose count is small enough for the Polya-urn shortcut grow from
//           the front, the others from the back. Each WARP repeatedly claims a chunk of items and
//           processes it on its own (no block barrier inside a level):
//             - Gamma variates, one lane per (item, outcome), into a per-warp shared-memory tile;
//             - a binary tree of partial sums over each item's outcomes (bottom-up), then the
//               item's count is pushed down that tree with one binomial per internal node
//               (top-down), again one lane per (item, tree node). This is rMultinomial
//               (distributions.cpp:21-53) with the conditional binomials arranged as a balanced
//               tree instead of a chain: the same multinomial law, log2(d) dependent steps instead
//               of d-1, and no 1-x cancellation because both child sums are at hand;
//             - children / finished ballots are compacted with warp ballots and appended to the
//               next frontier / the entry list with one shared-memory atomic per warp and queue;
//   (c)     IRV elimination over the election's (ballot, count) entries plus the shared observed
//           entries (socialChoiceIRV, irv_ballot.cpp:42-138);
//   (d)     win counts accumulated per block and flushed with one atomic per candidate.
#pragma once
#include <cstdint>

#include "split.cuh"

namespace dtb {

// Where the pieces of one block's scratch live.
template <int W>
struct Scratch {
  Item<W> *frontier[2];
  BallotKey<W> *ent_key;
  uint32_t *ent_cnt;
  uint8_t *ent_len;
  uint8_t *ent_holder;
  uint8_t *obs_holder;
  __host__ __device__ static uint64_t align16(uint64_t x) { return (x + 15) & ~15ull; }
  __host__ __device__ static uint64_t bytes(uint32_t cap_items, uint32_t cap_entries, uint32_t n_obs) {
    return 2 * align16((uint64_t)cap_items * sizeof(Item<W>)) + align16((uint64_t)cap_entries * sizeof(BallotKey<W>)) +
           align16((uint64_t)cap_entries * 4) + 2 * align16(cap_entries) + align16(n_obs);
  }
  __host__ __device__ void bind(uint8_t *p, uint32_t cap_items, uint32_t cap_entries) {
    frontier[0] = (Item<W> *)p; p += align16((uint64_t)cap_items * sizeof(Item<W>));
    frontier[1] = (Item<W> *)p; p += align16((uint64_t)cap_items * sizeof(Item<W>));
    ent_key = (BallotKey<W> *)p; p += align16((uint64_t)cap_entries * sizeof(BallotKey<W>));
    ent_cnt = (uint32_t *)p; p += align16((uint64_t)cap_entries * 4);
    ent_len = p; p += align16(cap_entries);
    ent_holder = p; p += align16(cap_entries);
    obs_holder = p;
  }
};

template <int BLOCK>
struct SharedFixed {
  // frontier bookkeeping of the level being processed / produced
  uint32_t n_small, n_big;        // items at the front (small counts) / back (Gamma path) of the input queue
  uint32_t next_small, next_big;  // chunk cursors
  uint32_t out_small, out_big;    // items appended to the output queue
  uint32_t n_entries;
  uint32_t overflow;
  // IRV
  uint32_t tally[kMaxCandidates];
  uint32_t wins[kMaxCandidates];
  uint32_t eliminated, chosen, tie_rounds, decided;
  uint32_t election;
  uint8_t order[kMaxCandidates];
};

// Small items (count <= urn_max <= 8): each lane finishes whole subtrees on its own, depth-first with
// an explicit stack of at most 8 frames (every frame holds >= 1 of the <= 8 ballots), applying the urn at
// each node and emitting finished ballots straight into the entry list. A lane whose stack runs empty
// claims the next item of the queue, so the warp keeps all lanes on the same loop body until the queue is
// drained. The draws at a node depend only on (election, node), so this gives the same ballots as pushing
// the item through the level-synchronous queue would.
template <int W, int BLOCK>
__device__ void walk_small_items(const SimArgs &a, Scratch<W> &sc, SharedFixed<BLOCK> &sm, const Item<W> *in,
                                 uint32_t n_small, uint32_t root_depth, const RngKey key, uint32_t *stat) {
  const int lane = threadIdx.x & 31;
  const uint32_t nc = a.P.nc;
  Item<W> stack[8];
  int top = 0;
  bool drained = false;
  for (;;) {
    const bool need = top == 0 && !drained;
    const uint32_t mask = __ballot_sync(0xffffffffu, need);
    if (mask) {
      const int leader = __ffs(mask) - 1;
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(&sm.next_small, (uint32_t)__popc(mask));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (need) {
        const uint32_t idx = base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
        if (idx < n_small) {
          stack[0] = in[idx];
          stack[0].pad_ = root_depth;
          top = 1;
        } else {
          drained = true;
        }
      }
    }
    if (!__any_sync(0xffffffffu, top > 0)) break;
    if (top == 0) continue;
    const Item<W> f = stack[--top];
    const LevelInfo lv = level_info(a.P, f.pad_);
    const uint32_t nbase = f.node >= 0 ? a.T.node_base[f.node] : 0xFFFFFFFFu;
    stat[kStatItems] += 1;
    stat[kStatUrnItems] += 1;
    if (f.node >= 0) stat[kStatSlots] += lv.d;
    const uint64_t picks = urn_picks(a, lv, f.node, nbase, f.count, key, f.node_key);
    // balls per outcome, 4 bits each (count <= 8, at most 33 outcomes); an outcome is emitted when its first ball
    // comes by and its counter is cleared, so every outcome is handled once without comparing balls pairwise
    uint64_t c4[3] = {0ull, 0ull, 0ull};
    for (uint32_t b = 0; b < f.count; ++b) {
      const uint32_t s = (uint32_t)(picks >> (8u * b)) & 0xFFu;
      c4[s >> 4] += 1ull << (4u * (s & 15u));
    }
    for (uint32_t b = 0; b < f.count; ++b) {
      const uint32_t slot = (uint32_t)(picks >> (8u * b)) & 0xFFu;
      const uint32_t c = (uint32_t)(c4[slot >> 4] >> (4u * (slot & 15u))) & 15u;
      if (c == 0u) continue;
      c4[slot >> 4] &= ~(15ull << (4u * (slot & 15u)));
      BallotKey<W> pk = f.prefix;
      uint32_t cand = 0;
      const bool stop = slot == lv.K;  // only possible when lv.T
      if (!stop) {
        cand = nbase != 0xFFFFFFFFu ? (uint32_t)a.T.slot_cand[nbase + slot] : nth_unused(f.used, nc, slot);
        key_set<W>(pk, (int)lv.depth, cand);
      }
      if (stop || lv.leaf_children) {
        const uint32_t eo = coalesced_reserve(&sm.n_entries);
        if (eo < a.cap_entries) {
          sc.ent_key[eo] = pk;
          sc.ent_cnt[eo] = c;
          sc.ent_len[eo] = (uint8_t)(stop ? lv.depth : lv.depth + 1);
        } else {
          sm.overflow = 1;
        }
      } else {
        Item<W> ch;
        ch.prefix = pk;
        ch.node_key = child_node_key(f.node_key, cand);
        ch.count = c;
        ch.node = nbase != 0xFFFFFFFFu ? a.T.slot_child[nbase + slot] : -1;
        ch.used = f.used | (1u << cand);
        ch.pad_ = lv.depth + 1;
        stack[top++] = ch;
      }
    }
  }
}

// One warp expands one chunk of m items of the same depth whose counts exceed urn_max.
template <int W, int BLOCK, bool LOG>
__device__ void expand_chunk(const SimArgs &a, Scratch<W> &sc, SharedFixed<BLOCK> &sm, WarpTile<W> &wt,
                             const Item<W> *in, uint32_t m, const LevelInfo &lv, int log_dp, Item<W> *out,
                             const RngKey key, uint32_t *stat) {
  const int lane = threadIdx.x & 31;
  const uint32_t nc = a.P.nc;
  const uint32_t d = lv.d, K = lv.K;
  const uint32_t dp = 1u << log_dp, stride = 2u * dp + 1u;

  for (uint32_t it = lane; it < m; it += 32) {
    Item<W> x = in[it];
    x.pad_ = lv.depth;
    wt.item[it] = x;
    wt.nbase[it] = x.node >= 0 ? a.T.node_base[x.node] : 0xFFFFFFFFu;
    stat[kStatItems] += 1;
    if (x.node >= 0) stat[kStatSlots] += d;
  }
  __syncwarp();
  split_items<W, LOG>(a, wt, m, log_dp, key, stat);

  // ---- emission: one lane per (item, outcome); ballots for warp-level compaction -----------------
  for (uint32_t base = 0; base < (m << log_dp); base += 32) {
    const uint32_t p = base + lane;
    const uint32_t it = p >> log_dp, slot = p & (dp - 1u);
    uint32_t c = 0;
    if (p < (m << log_dp) && slot < d) c = wt.val[it * stride + dp + slot];
    // destination: 0 none, 1 entry, 2 small-count queue, 3 big queue
    int dest = 0;
    if (c) {
      if (slot == K || lv.leaf_children) dest = 1;  // slot K: the ballot stops here (only when T)
      else dest = c <= a.urn_max ? 2 : 3;
    }
    const uint32_t eo = warp_reserve(&sm.n_entries, dest == 1, lane);
    const uint32_t so = warp_reserve(&sm.out_small, dest == 2, lane);
    const uint32_t bo = warp_reserve(&sm.out_big, dest == 3, lane);
    if (dest == 0) continue;
    const Item<W> &par = wt.item[it];
    BallotKey<W> pk = par.prefix;
    uint32_t cand = 0;
    const uint32_t nb = wt.nbase[it];
    if (slot != K) {
      cand = nb != 0xFFFFFFFFu ? (uint32_t)a.T.slot_cand[nb + slot] : nth_unused(par.used, nc, slot);
      key_set<W>(pk, (int)lv.depth, cand);
    }
    if (dest == 1) {
      if (eo < a.cap_entries) {
        sc.ent_key[eo] = pk;
        sc.ent_cnt[eo] = c;
        sc.ent_len[eo] = (uint8_t)(slot == K ? lv.depth : lv.depth + 1);
      } else {
        sm.overflow = 1;
      }
    } else {
      Item<W> ch;
      ch.prefix = pk;
      ch.node_key = child_node_key(par.node_key, cand);
      ch.count = c;
      ch.node = nb != 0xFFFFFFFFu ? a.T.slot_child[nb + slot] : -1;
      ch.used = par.used | (1u << cand);
      ch.pad_ = lv.depth + 1;
      // two-ended queue: small items from the front, big items from the back; they must not meet.
      // Both cursors already include this lane's reservation, so the test is conservative.
      if (sm.out_small + sm.out_big > a.cap_items) {
        sm.overflow = 1;
      } else if (dest == 2) {
        out[so] = ch;
      } else {
        out[a.cap_items - 1u - bo] = ch;
      }
    }
  }
  __syncwarp();
}

// One election's expansion. Leaves the entries in sc.ent_* and their number in sm.n_entries.
template <int W, int BLOCK, bool LOG>
__device__ void expand_election(const SimArgs &a, Scratch<W> &sc, SharedFixed<BLOCK> &sm, WarpTile<W> *tiles,
                                const RngKey key, uint32_t *stat) {
  typedef Chunk<W> C;
  const uint32_t nc = a.P.nc;
  const int tid = threadIdx.x, lane = tid & 31;
  WarpTile<W> &wt = tiles[tid >> 5];

  if (tid == 0) {
    sm.n_entries = 0;
    sm.overflow = 0;
    sm.n_small = 0;
    sm.n_big = 0;
    if (a.n_sample > 0) {
      Item<W> root;
#pragma unroll
      for (int w = 0; w < W; ++w) root.prefix.w[w] = 0;
      root.node_key = kRootNodeKey;
      root.count = a.n_sample;
      root.node = 0;
      root.used = 0;
      root.pad_ = 0;
      if (a.n_sample <= a.urn_max) {
        sc.frontier[0][0] = root;
        sm.n_small = 1;
      } else {
        sc.frontier[0][a.cap_items - 1] = root;
        sm.n_big = 1;
      }
    }
  }
  __syncthreads();

  int cur = 0;
  for (uint32_t depth = 0; depth + 1 < nc; ++depth) {
    const uint32_t n_small = sm.n_small, n_big = sm.n_big;
    if (n_small + n_big == 0) break;
    const LevelInfo lv = level_info(a.P, depth);
    const Item<W> *in = sc.frontier[cur];
    Item<W> *out = sc.frontier[cur ^ 1];
    if (tid == 0) {
      sm.next_small = 0;
      sm.next_big = 0;
      sm.out_small = 0;
      sm.out_big = 0;
    }
    __syncthreads();
    // Small items of this depth are walked to completion (they never re-enter a queue).
    if (n_small) walk_small_items<W, BLOCK>(a, sc, sm, in, n_small, depth, key, stat);
    // Big items: warps claim chunks and run the Gamma / binomial-tree path.
    int log_dp = 1;
    while ((1u << log_dp) < lv.d) ++log_dp;
    const uint32_t ci = min((uint32_t)C::kMaxItems, (uint32_t)C::kTileWords / (2u * (1u << log_dp) + 1u));
    for (;;) {
      uint32_t start = 0;
      if (lane == 0) start = atomicAdd(&sm.next_big, ci);
      start = __shfl_sync(0xffffffffu, start, 0);
      if (start >= n_big) break;
      const uint32_t m = min(ci, n_big - start);
      // big items sit at in[cap-1], in[cap-2], ...: a chunk is the ascending range ending at cap-1-start
      expand_chunk<W, BLOCK, LOG>(a, sc, sm, wt, in + (a.cap_items - start - m), m, lv, log_dp, out, key, stat);
    }
    __syncthreads();
    if (tid == 0) {
      sm.n_small = lv.leaf_children ? 0u : sm.out_small;
      sm.n_big = lv.leaf_children ? 0u : sm.out_big;
      if (sm.overflow) *a.error_flag = 1;
    }
    cur ^= 1;
    __syncthreads();
    if (sm.overflow) break;
  }
  if (tid == 0 && sm.n_entries > a.cap_entries) sm.n_entries = a.cap_entries;
  __syncthreads();
}

// socialChoiceIRV (irv_ballot.cpp:42-138) for one election. Every entry carries its current
// holder (first standing preference); eliminating a candidate re-examines only its entries.
template <int W, int BLOCK>
__device__ void irv_election(const SimArgs &a, Scratch<W> &sc, SharedFixed<BLOCK> &sm, const RngKey key,
                             uint32_t n_rounds, uint32_t *stat) {
  const uint32_t nc = a.P.nc;
  const int tid = threadIdx.x;
  const uint32_t n_ent = sm.n_entries;
  const uint32_t n_obs = a.n_obs;
  const BallotKey<W> *obs_key = (const BallotKey<W> *)a.obs_key;

  if (tid < (int)kMaxCandidates) sm.tally[tid] = (tid < (int)nc && n_obs) ? a.obs_tally0[tid] : 0u;
  if (tid == 0) {
    sm.eliminated = 0;
    sm.tie_rounds = 0;
  }
  __syncthreads();
  // first preferences (irv_ballot.cpp:78-83); blank ballots are dropped (:55-56)
  for (uint32_t base = 0; base < n_ent; base += BLOCK) {
    uint32_t i = base + tid;
    bool valid = false;
    uint32_t cand = 0, cnt = 0;
    if (i < n_ent) {
      uint32_t len = sc.ent_len[i];
      cand = len ? key_get<W>(sc.ent_key[i], 0) : 0xFFu;
      sc.ent_holder[i] = (uint8_t)cand;
      cnt = sc.ent_cnt[i];
      valid = len != 0;
    }
    warp_tally_add(sm.tally, valid, cand, cnt);
  }
  for (uint32_t i = tid; i < n_obs; i += BLOCK) sc.obs_holder[i] = a.obs_first[i];
  __syncthreads();

  for (uint32_t round = 0; round < n_rounds; ++round) {
    if (tid == 0) {
      // candidates with the minimal tally among those standing (irv_ballot.cpp:88-98)
      uint32_t elim = sm.eliminated, mn = 0xFFFFFFFFu, tied = 0, ntied = 0;
      uint32_t total = 0, top = 0, top_c = 0;
      for (uint32_t c = 0; c < nc; ++c) {
        if (elim >> c & 1u) continue;
        uint32_t t = sm.tally[c];
        if (t < mn) { mn = t; tied = 1u << c; ntied = 1; }
        else if (t == mn) { tied |= 1u << c; ++ntied; }
        total += t;
        if (t > top) { top = t; top_c = c; }
      }
      // A candidate holding more than half of the ballots still in play can never have the minimal tally
      // again (tallies only grow, ballots in play only shrink), so it is the winner: stop here.
      sm.decided = 0;
      if (a.n_winners == 1u && a.elim_orders == nullptr && top > total - top) {
        sm.eliminated = ~(1u << top_c) & (nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u));
        sm.decided = 1;
      }
      uint32_t pick = 0;
      if (ntied > 1) {  // uniform among the tied, ascending index (irv_ballot.cpp:100-102)
        U4 r = node_random(key, 0ull, kStreamTie, round);
        pick = uniform_below(r.x, ntied);
        sm.tie_rounds += 1;
      }
      for (uint32_t k = 0; k < pick; ++k) tied &= tied - 1u;
      uint32_t chosen = (uint32_t)__ffs(tied) - 1u;
      sm.chosen = chosen;
      sm.order[round] = (uint8_t)chosen;
      if (!sm.decided) sm.eliminated = elim | (1u << chosen);
    }
    __syncthreads();
    if (sm.decided) break;
    const uint32_t chosen = sm.chosen, eliminated = sm.eliminated;
    const bool any = sm.tally[chosen] != 0u;
    __syncthreads();
    if (!any) continue;  // nobody holds ballots for this candidate: nothing to redistribute
    // After the last elimination nobody reads the tallies again, so the loser's ballots stay where they are
    // (the reference redistributes them, irv_ballot.cpp:109-133, without any effect on its output).
    if (round + 1u == n_rounds) continue;
    // redistribution (irv_ballot.cpp:109-133): scan the holder bytes 16 at a time; only the entries
    // held by the eliminated candidate are re-examined
    const uint32_t pattern = chosen * 0x01010101u;
    for (int part = 0; part < 2; ++part) {
      uint8_t *holder = part == 0 ? sc.ent_holder : sc.obs_holder;
      const uint32_t n_part = part == 0 ? n_ent : n_obs;
      const uint32_t n_vec = (n_part + 15u) / 16u;
      for (uint32_t vbase = 0; vbase < n_vec; vbase += BLOCK) {
        const uint32_t v = vbase + tid;
        uint32_t hw[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        if (v < n_vec) {
          const uint4 h = reinterpret_cast<const uint4 *>(holder)[v];
          hw[0] = h.x; hw[1] = h.y; hw[2] = h.z; hw[3] = h.w;
        }
        // a matching byte is rare; the warp goes through its hits together
        uint32_t hits = 0;  // bit j: byte j of this lane's 16 equals `chosen`
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t eq = __vcmpeq4(hw[q], pattern);  // 0xFF per equal byte
          hits |= ((eq & 1u) | ((eq >> 7) & 2u) | ((eq >> 14) & 4u) | ((eq >> 21) & 8u)) << (4 * q);
        }
        while (__any_sync(0xffffffffu, hits != 0u)) {
          bool valid = false;
          uint32_t cand = 0, cnt = 0;
          if (hits) {
            const uint32_t j = (uint32_t)__ffs(hits) - 1u;
            hits &= hits - 1u;
            const uint32_t i = v * 16u + j;
            if (i < n_part) {
              if (part == 0) {
                cand = first_standing<W>(sc.ent_key[i], sc.ent_len[i], eliminated);
                cnt = sc.ent_cnt[i];
              } else {
                // an observed key stores at most nc-1 preferences; its length is implied by repetition
                cand = first_standing<W>(obs_key[i], nc - 1, eliminated);
                cnt = a.obs_cnt[i];
              }
              holder[i] = (uint8_t)cand;
              valid = cand != 0xFFu;
              stat[kStatRereads] += 1;
            }
          }
          warp_tally_add(sm.tally, valid, cand, cnt);
        }
      }
    }
    __syncthreads();
  }
}

template <int W, int BLOCK, bool LOG>
__global__ void __launch_bounds__(BLOCK) simulate_kernel(const __grid_constant__ SimArgs a) {
  __shared__ SharedFixed<BLOCK> sm;
  __shared__ WarpTile<W> tiles[BLOCK / 32];
  const int tid = threadIdx.x;
  const uint32_t nc = a.P.nc;
  Scratch<W> sc;
  sc.bind(a.scratch + (uint64_t)blockIdx.x * a.scratch_stride, a.cap_items, a.cap_entries);
  uint32_t stat[kStatCount];
#pragma unroll
  for (int i = 0; i < kStatCount; ++i) stat[i] = 0;
  if (tid < (int)kMaxCandidates) sm.wins[tid] = 0;

  for (;;) {
    __syncthreads();
    if (tid == 0) sm.election = atomicAdd(a.work_counter, 1u);
    __syncthreads();
    const uint32_t e = sm.election;
    if (e >= a.n_elections) break;
    const RngKey key = election_key(a.seed, a.first_election + e);

    expand_election<W, BLOCK, LOG>(a, sc, sm, tiles, key, stat);
    if (tid == 0) {
      stat[kStatEntries] += sm.n_entries;
      if (a.entry_count_out) *a.entry_count_out = sm.n_entries;
    }
    if (!a.run_irv) continue;

    // With the full order requested run nc-1 rounds (the survivor is appended); otherwise stop
    // when n_winners candidates are left (R_tree.cpp:245-253 only reads the last n_winners).
    const uint32_t n_rounds = a.elim_orders ? nc - 1 : nc - a.n_winners;
    irv_election<W, BLOCK>(a, sc, sm, key, n_rounds, stat);
    if (tid == 0) {
      uint32_t standing = ~sm.eliminated & (nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u));
      if (a.elim_orders) {
        uint8_t *o = a.elim_orders + (uint64_t)e * nc;
        for (uint32_t r = 0; r + 1 < nc; ++r) o[r] = sm.order[r];
        o[nc - 1] = (uint8_t)(__ffs(standing) - 1);
        // winners = last n_winners of the order
        for (uint32_t r = nc - a.n_winners; r < nc; ++r) sm.wins[o[r]] += 1;
      } else {
        while (standing) {
          sm.wins[__ffs(standing) - 1] += 1;
          standing &= standing - 1u;
        }
      }
      if (a.tie_rounds_out) a.tie_rounds_out[e] = sm.tie_rounds;
    }
  }
  __syncthreads();
  // (d) flush: one atomic per candidate per block
  if (a.run_irv && tid < (int)nc && sm.wins[tid]) atomicAdd(&a.win_counts[tid], (unsigned long long)sm.wins[tid]);
  if (a.stats) {
#pragma unroll
    for (int i = 0; i < kStatCount; ++i) {
      uint32_t s = __reduce_add_sync(0xffffffffu, stat[i]);
      if ((tid & 31) == 0 && s) atomicAdd(&a.stats[i], (unsigned long long)s);
    }
  }
}

}  // namespace dtb
