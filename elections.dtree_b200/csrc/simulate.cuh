// simulate.cuh — the materialising path: every sampled ballot of an election is created, then IRV runs over them.
// Per election, three stages; simulate_kernel runs them in one persistent block per election, sim_big_kernel /
// sim_walk_kernel / sim_irv_kernel run the same stages as three launches over a wave of elections (each at its own
// register budget and occupancy), handing the counts over through a header in the election's scratch:
//
//   (a) level loop (IRVNode::sample, irv_node.cpp:107-174): items with more than urn_max ballots, level by level.
//       Each WARP repeatedly claims a chunk of items and processes it on its own (no block barrier inside a level):
//         - Gamma variates, one lane per (item, outcome), into a per-warp shared-memory tile;
//         - a binary tree of partial sums over each item's outcomes (bottom-up), then the item's count is pushed down
//           that tree with one binomial per internal node (top-down), again one lane per (item, tree node). This is
//           rMultinomial (distributions.cpp:21-53) with the conditional binomials arranged as a balanced tree instead
//           of a chain: the same multinomial law, log2(d) dependent steps instead of d-1, and no 1-x cancellation
//           because both child sums are at hand;
//         - children are compacted with warp ballots into the next frontier, finished ballots into the entry list,
//           children with <= urn_max ballots into the election's small queue;
//   (b) walkers (and lazyIRVBallots, :27-84): the small queue, every item an independent subtree of <= 8 ballots
//       walked to completion with the Polya urn, from per-warp stacks in shared memory (walk_small_items);
//   (c) IRV elimination over the election's (ballot, count) entries plus the shared observed entries
//       (socialChoiceIRV, irv_ballot.cpp:42-138);
//   (d) win counts accumulated per block and flushed with one atomic per candidate.
#pragma once
#include <cstdint>

#include "split.cuh"

namespace dtb {

// Profiling build (-DDTB_PHASE_CLOCKS): lane 0 of every warp accumulates the cycles (in units of 64) it spends in the
// small-item walk, the big-item chunks, the barrier at the end of a level and the IRV stage; simulate_kernel reports
// them in place of the gammas / binomial_attempts / slots / irv_rereads counters. Never defined in the shipped library.
#ifdef DTB_PHASE_CLOCKS
#define DTB_CLK_BEGIN(t) const long long t = clock64()
#define DTB_CLK_END(t, i) do { if ((threadIdx.x & 31) == 0) stat[kStatCount + (i)] += (uint32_t)((clock64() - (t)) >> 6); } while (0)
constexpr int kClkSlots = 8;
#else
#define DTB_CLK_BEGIN(t)
#define DTB_CLK_END(t, i)
constexpr int kClkSlots = 0;
#endif

// One sampled (ballot, count) entry: a single 16-byte record for keys of one word (24 / 32 bytes for two / three), so that
// writing an entry is one store and re-examining it in the IRV stage touches one 32-byte sector instead of three.
template <int W>
struct alignas(W == 2 ? 8 : 16) Entry {
  BallotKey<W> key;
  uint32_t cnt;
  uint32_t len;
};

static_assert(sizeof(Entry<1>) == 16 && sizeof(Entry<2>) == 24 && sizeof(Entry<3>) == 32, "capi.cu packs these records");

// Where the pieces of one election's scratch live.
template <int W>
struct Scratch {
  uint32_t *hdr;       // kHdr* words: what one phase kernel hands to the next (phased launches)
  Item<W> *small;      // items with count <= urn_max, all depths: walked to completion after the level loop
  Item<W> *frontier[2];
  Entry<W> *ent;
  uint8_t *ent_holder;
  uint8_t *obs_holder;
  __host__ __device__ static uint64_t align16(uint64_t x) { return (x + 15) & ~15ull; }
  __host__ __device__ static uint64_t bytes(uint32_t cap_items, uint32_t cap_entries, uint32_t n_obs,
                                            uint32_t cap_small = 0) {
    return 16 + align16((uint64_t)cap_small * sizeof(Item<W>)) + 2 * align16((uint64_t)cap_items * sizeof(Item<W>)) +
           align16((uint64_t)cap_entries * sizeof(Entry<W>)) + align16(cap_entries) + align16(n_obs);
  }
  __host__ __device__ void bind(uint8_t *p, uint32_t cap_items, uint32_t cap_entries, uint32_t cap_small = 0) {
    hdr = (uint32_t *)p; p += 16;
    small = (Item<W> *)p; p += align16((uint64_t)cap_small * sizeof(Item<W>));
    frontier[0] = (Item<W> *)p; p += align16((uint64_t)cap_items * sizeof(Item<W>));
    frontier[1] = (Item<W> *)p; p += align16((uint64_t)cap_items * sizeof(Item<W>));
    ent = (Entry<W> *)p; p += align16((uint64_t)cap_entries * sizeof(Entry<W>));
    ent_holder = p; p += align16(cap_entries);
    obs_holder = p;
  }
};

enum { kHdrSmall = 0, kHdrEntries, kHdrOverflow, kHdrWords = 4 };

template <int BLOCK>
struct SharedFixed {
  // frontier bookkeeping of the level being processed / produced
  uint32_t n_big;                 // items of the input frontier (all with count > urn_max: the Gamma path)
  uint32_t next_small, next_big;  // chunk cursors
  uint32_t out_small, out_big;    // items appended to the small queue (never reset) / the output frontier
  uint32_t n_entries;
  uint32_t overflow;
  // IRV
  uint32_t tally[kMaxCandidates];
  uint32_t wins[kMaxCandidates];
  uint32_t eliminated, chosen, tie_rounds, decided;
  uint32_t election;
  uint8_t order[kMaxCandidates];
};

// Small items (count <= urn_max <= 8) never re-enter the level queues: a warp finishes their whole subtrees. Frames
// (an item = an independent subtree) wait in three LIFO stacks per warp, by ballots held: 1 / 2-3 / 4-8. A pass pops up
// to 32 frames of ONE stack, one per lane, applies the urn at each node and pushes the children (finished ballots go
// to the entry list, one atomic per warp and emission step). Which lane takes which frame is irrelevant, so all lanes
// stay busy until queue and stacks are empty; and because the frames of a pass hold about the same number of ballots,
// the per-ball and per-outcome loops have the same trip count in every lane (with mixed frames the warp ran 8 rounds
// for a mean of 1.9 balls: 8.7 of 32 lanes active). Single-ballot frames - more than half of them - take one round.
//
// Memory: the 1-ballot stack is this warp's slice of dynamic shared memory, the other two grow towards each other in
// the warp's tile (the big-item path is not using it meanwhile). Neither can overflow: a frame holds >= 1 ballot (>= 2,
// >= 4 in the other stacks), children never hold more ballots than their parent, and frames are drawn from the
// level's queue only while the ballots held stay <= kBudget = frames per tile. The draws at a node depend only on
// (election, node), so this gives the same ballots as pushing the item through the level-synchronous queue would.
template <int W, uint32_t B>
struct WalkStacks {
  Item<W> *light, *tile;  // B frames (1 ballot each), B / 2 frames (2-3 ballots from the front, 4-8 from the back)
  uint32_t n[3];          // warp-uniform
  __device__ __forceinline__ Item<W> *slot(int bucket, uint32_t i) const {
    return bucket == 0 ? light + i : bucket == 1 ? tile + i : tile + (B / 2u - 1u - i);
  }
};
// Ballot budget of a warp's stacks. In simulate_kernel the 2-3 / 4-8 stacks live in the warp's tile and the single-ballot
// stack in dynamic shared memory; the phase kernel sim_walk_kernel has both in dynamic shared memory and a smaller budget,
// so that four blocks fit an SM.
template <int W>
struct WalkBudget {
  static constexpr uint32_t kFused = (uint32_t)(sizeof(WarpTile<W>) / sizeof(Item<W>)) & ~1u;
  static constexpr uint32_t kPhased = W == 1 ? 128 : 80;  // 32 warps x 1.5 x kPhased frames within 227 KB
};

__device__ __forceinline__ int walk_bucket(uint32_t count) { return count <= 1u ? 0 : count <= 3u ? 1 : 2; }

template <int W, int BLOCK, uint32_t B>
__device__ void walk_small_items(const SimArgs &a, Scratch<W> &sc, SharedFixed<BLOCK> &sm, Item<W> *tile_stack,
                                 Item<W> *light_stack, const Item<W> *in, uint32_t n_small, const RngKey key,
                                 uint32_t *stat) {
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t nc = a.P.nc;
  constexpr uint32_t kBudget = B;
  WalkStacks<W, B> st;
  st.light = light_stack;
  st.tile = tile_stack;
  st.n[0] = st.n[1] = st.n[2] = 0;
  uint32_t ballots = 0;             // held by the frames in the stacks (warp-uniform)
  uint32_t pend_lo = 0, pend_hi = 0;  // queue range claimed by this warp, not yet taken in
  bool drained = false;             // nothing left to claim
  for (;;) {
    int bucket = st.n[2] >= 32u ? 2 : st.n[1] >= 32u ? 1 : st.n[0] >= 32u ? 0 : -1;
    if (bucket < 0) {
      // ---- take frames in from the level's queue, as many as the budget allows --------------------
      if (pend_lo == pend_hi && !drained) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&sm.next_small, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base < n_small) {
          pend_lo = base;
          pend_hi = min(base + 32u, n_small);
        }
        drained = base + 32u >= n_small;
      }
      uint32_t took = 0;
      if (pend_lo < pend_hi) {
        const uint32_t idx = pend_lo + (uint32_t)lane;
        Item<W> f;
        f.count = 0;
        if (idx < pend_hi) f = in[idx];
        // usually all of them fit; otherwise the longest prefix that does
        const uint32_t all = __reduce_add_sync(0xffffffffu, f.count);
        bool fits = idx < pend_hi;
        uint32_t taken = all;
        if (ballots + all > kBudget) {
          uint32_t incl = f.count;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
          }
          fits = fits && ballots + incl <= kBudget;
          taken = __reduce_add_sync(0xffffffffu, fits ? f.count : 0u);
        }
        const uint32_t mfit = __ballot_sync(0xffffffffu, fits);
        took = (uint32_t)__popc(mfit);
        if (took) {
          const int bk = walk_bucket(f.count);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const uint32_t m = __ballot_sync(0xffffffffu, fits && bk == k);
            if (fits && bk == k) *st.slot(k, st.n[k] + (uint32_t)__popc(m & lt_mask)) = f;
            st.n[k] += (uint32_t)__popc(m);
          }
          ballots += taken;
          pend_lo += took;
          __syncwarp();
          continue;
        }
      }
      // nothing could be taken in: work on what there is, heaviest first
      bucket = st.n[2] ? 2 : st.n[1] ? 1 : st.n[0] ? 0 : -1;
      if (bucket < 0) {
        if (pend_lo == pend_hi && drained) break;
        continue;  // unreachable: an empty pool always has room for a frame
      }
    }

    // ---- one pass over up to 32 frames of one stack ------------------------------------------------
    const uint32_t n_before = bucket == 0 ? st.n[0] : bucket == 1 ? st.n[1] : st.n[2], take = min(n_before, 32u);
    const bool have = (uint32_t)lane < take;
    Item<W> f;
    f.count = 0;
    f.pad_ = 0;
    if (have) f = *st.slot(bucket, n_before - 1u - (uint32_t)lane);
    if (bucket == 0) st.n[0] -= take; else if (bucket == 1) st.n[1] -= take; else st.n[2] -= take;
    ballots -= __reduce_add_sync(0xffffffffu, f.count);
    __syncwarp();

    // balls per outcome, 4 bits each (count <= 8, at most 33 outcomes)
    uint64_t c0 = 0ull, c1 = 0ull, c2 = 0ull;
    const LevelInfo lv = level_info(a.P, f.pad_);
    uint32_t nbase = 0xFFFFFFFFu;
    if (have) {
      if (f.node >= 0) nbase = a.T.node_base[f.node];
      stat[kStatItems] += 1;
      stat[kStatUrnItems] += 1;
      if (f.node >= 0) stat[kStatSlots] += lv.d;
#ifdef DTB_PHASE_CLOCKS
      stat[kStatCount + (f.count == 1u ? 4 : f.count <= 3u ? 5 : 6)] += 1;
      if (lane == 0) stat[kStatCount + 7] += 1;
#endif
      const uint64_t picks = urn_picks(a, lv, f.node, nbase, f.count, key, f.node_key);
      if (nc < 16u) {  // at most 16 outcomes anywhere in the tree: one word of nibbles
        for (uint32_t b = 0; b < f.count; ++b) c0 += 1ull << (4u * ((uint32_t)(picks >> (8u * b)) & 15u));
      } else {
        for (uint32_t b = 0; b < f.count; ++b) {
          const uint32_t s = (uint32_t)(picks >> (8u * b)) & 0xFFu;
          const uint64_t inc = 1ull << (4u * (s & 15u));
          c0 += (s >> 4) == 0u ? inc : 0ull;
          c1 += (s >> 4) == 1u ? inc : 0ull;
          c2 += (s >> 4) == 2u ? inc : 0ull;
        }
      }
    }
    // emission, one outcome per lane and step: ballots stop (entry list) or go on as a child frame
    uint32_t pushed = 0;
    for (;;) {
      const bool more = (c0 | c1 | c2) != 0ull;
      if (!__any_sync(0xffffffffu, more)) break;
      int dest = -1;  // 0..2: child frame into that stack, 3: entry
      uint32_t slot = 0, c = 0;
      if (more) {
        const int w = c0 ? 0 : c1 ? 1 : 2;
        const uint64_t cw = w == 0 ? c0 : w == 1 ? c1 : c2;
        const uint32_t pos = ((uint32_t)__ffsll((long long)cw) - 1u) & ~3u;
        c = (uint32_t)(cw >> pos) & 15u;
        const uint64_t cleared = cw & ~(15ull << pos);
        if (w == 0) c0 = cleared; else if (w == 1) c1 = cleared; else c2 = cleared;
        slot = 16u * (uint32_t)w + (pos >> 2);
        dest = (slot == lv.K || lv.leaf_children) ? 3 : walk_bucket(c);  // slot K: the ballot stops here (only when lv.T)
      }
      const uint32_t me = __ballot_sync(0xffffffffu, dest == 3);
      uint32_t ebase = 0;
      if (me) {
        if (lane == 0) ebase = atomicAdd(&sm.n_entries, (uint32_t)__popc(me));
        ebase = __shfl_sync(0xffffffffu, ebase, 0);
      }
      uint32_t at = ebase + (uint32_t)__popc(me & lt_mask);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const uint32_t m = __ballot_sync(0xffffffffu, dest == k);
        if (dest == k) at = st.n[k] + (uint32_t)__popc(m & lt_mask);
        st.n[k] += (uint32_t)__popc(m);
      }
      if (dest < 0) continue;
      BallotKey<W> pk = f.prefix;
      uint32_t cand = 0;
      const bool stop = slot == lv.K;
      if (!stop) {
        cand = nbase != 0xFFFFFFFFu ? (uint32_t)a.T.slot_cand[nbase + slot] : nth_unused(f.used, nc, slot);
        key_set<W>(pk, (int)lv.depth, cand);
      }
      if (dest == 3) {
        if (at < a.cap_entries) {
          Entry<W> en;
          en.key = pk;
          en.cnt = c;
          en.len = stop ? lv.depth : lv.depth + 1;
          sc.ent[at] = en;
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
        *st.slot(dest, at) = ch;
        pushed += c;
      }
    }
    ballots += __reduce_add_sync(0xffffffffu, pushed);
    __syncwarp();
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
        Entry<W> en;
        en.key = pk;
        en.cnt = c;
        en.len = slot == K ? lv.depth : lv.depth + 1;
        sc.ent[eo] = en;
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
      // small counts wait in the election's small queue (walked after the level loop), the others form the next frontier
      if (dest == 2) {
        if (so < a.cap_small) sc.small[so] = ch;
        else sm.overflow = 1;
      } else {
        if (bo < a.cap_items) out[bo] = ch;
        else sm.overflow = 1;
      }
    }
  }
  __syncwarp();
}

// The level loop of one election: items with count > urn_max, level-synchronously. Leaves their finished ballots in
// sc.ent_* (sm.n_entries) and every child with count <= urn_max in sc.small (sm.out_small).
template <int W, int BLOCK, bool LOG>
__device__ void expand_big_levels(const SimArgs &a, Scratch<W> &sc, SharedFixed<BLOCK> &sm, WarpTile<W> *tiles,
                                  const RngKey key, uint32_t *stat) {
  typedef Chunk<W> C;
  const uint32_t nc = a.P.nc;
  const int tid = threadIdx.x, lane = tid & 31;
  WarpTile<W> &wt = tiles[tid >> 5];

  if (tid == 0) {
    sm.n_entries = 0;
    sm.overflow = 0;
    sm.out_small = 0;
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
        sc.small[0] = root;
        sm.out_small = 1;
      } else {
        sc.frontier[0][0] = root;
        sm.n_big = 1;
      }
    }
  }
  __syncthreads();

  int cur = 0;
  for (uint32_t depth = 0; depth + 1 < nc; ++depth) {
    const uint32_t n_big = sm.n_big;
    if (n_big == 0) break;
    const LevelInfo lv = level_info(a.P, depth);
    const Item<W> *in = sc.frontier[cur];
    Item<W> *out = sc.frontier[cur ^ 1];
    if (tid == 0) {
      sm.next_big = 0;
      sm.out_big = 0;
    }
    __syncthreads();
    // warps claim chunks and run the Gamma / binomial-tree path
    int log_dp = 1;
    while ((1u << log_dp) < lv.d) ++log_dp;
    const uint32_t ci = min((uint32_t)C::kMaxItems, (uint32_t)C::kTileWords / (2u * (1u << log_dp) + 1u));
    DTB_CLK_BEGIN(t_big);
    for (;;) {
      uint32_t start = 0;
      if (lane == 0) start = atomicAdd(&sm.next_big, ci);
      start = __shfl_sync(0xffffffffu, start, 0);
      if (start >= n_big) break;
      const uint32_t m = min(ci, n_big - start);
      expand_chunk<W, BLOCK, LOG>(a, sc, sm, wt, in + start, m, lv, log_dp, out, key, stat);
    }
    DTB_CLK_END(t_big, 1);
    DTB_CLK_BEGIN(t_wait);
    __syncthreads();
    DTB_CLK_END(t_wait, 2);
    if (tid == 0) sm.n_big = lv.leaf_children ? 0u : min(sm.out_big, a.cap_items);
    cur ^= 1;
    __syncthreads();
    if (sm.overflow) break;
  }
  __syncthreads();
}

// The small queue of one election, walked to completion by all warps of the block. tile_stacks / light_stacks: B / 2
// and B frames of shared memory per warp.
template <int W, int BLOCK, uint32_t B>
__device__ void walk_election(const SimArgs &a, Scratch<W> &sc, SharedFixed<BLOCK> &sm, Item<W> *tile_stack,
                              Item<W> *light_stack, const RngKey key, uint32_t *stat) {
  const int tid = threadIdx.x;
  if (tid == 0) sm.next_small = 0;
  __syncthreads();
  const uint32_t n_small = min(sm.out_small, a.cap_small);
  DTB_CLK_BEGIN(t_small);
  if (n_small) walk_small_items<W, BLOCK, B>(a, sc, sm, tile_stack, light_stack, sc.small, n_small, key, stat);
  DTB_CLK_END(t_small, 0);
  __syncthreads();
  if (tid == 0) {
    if (sm.n_entries > a.cap_entries) sm.n_entries = a.cap_entries;
    if (sm.overflow) *a.error_flag = 1;
  }
  __syncthreads();
}

// Per-lane tallies: row c of a warp's 32 x 32 words holds what each of its lanes has added to candidate c, so adding is
// a plain shared-memory read-modify-write without atomics or bank conflicts. flush_private_tallies sums the rows into
// the block's tallies (lane c takes row c, starting at a different column per lane) and clears them.
constexpr int kPrivTallyWords = 32 * kMaxCandidates;
static_assert(kPrivTallyWords <= Chunk<1>::kTileWords, "a warp's tile doubles as its private tallies");

__device__ __forceinline__ void flush_private_tallies(uint32_t *priv, uint32_t *s_tally, uint32_t nc) {
  const uint32_t lane = threadIdx.x & 31;
  __syncwarp();
  if (lane < nc) {
    uint32_t sum = 0;
#pragma unroll 8
    for (uint32_t j = 0; j < 32; ++j) {
      const uint32_t col = (j + lane) & 31u;
      sum += priv[lane * 32u + col];
      priv[lane * 32u + col] = 0u;
    }
    if (sum) atomicAdd(&s_tally[lane], sum);
  }
  __syncwarp();
}

// socialChoiceIRV (irv_ballot.cpp:42-138) for one election. Every entry carries its current
// holder (first standing preference); eliminating a candidate re-examines only its entries.
// `priv` is this warp's kPrivTallyWords words of shared memory (cleared here).
template <int W, int BLOCK>
__device__ void irv_election(const SimArgs &a, Scratch<W> &sc, SharedFixed<BLOCK> &sm, uint32_t *priv, const RngKey key,
                             uint32_t n_rounds, uint32_t *stat) {
  const uint32_t nc = a.P.nc;
  const int tid = threadIdx.x;
  const uint32_t lane = tid & 31;
  const uint32_t n_ent = sm.n_entries;
  const uint32_t n_obs = a.n_obs;
  const BallotKey<W> *obs_key = (const BallotKey<W> *)a.obs_key;

  if (tid < (int)kMaxCandidates) sm.tally[tid] = (tid < (int)nc && n_obs) ? a.obs_tally0[tid] : 0u;
  if (tid == 0) {
    sm.eliminated = 0;
    sm.tie_rounds = 0;
  }
#pragma unroll 8
  for (uint32_t j = 0; j < (uint32_t)kMaxCandidates; ++j) priv[j * 32u + lane] = 0u;
  __syncthreads();
  // first preferences (irv_ballot.cpp:78-83); blank ballots are dropped (:55-56). Four entries per thread and step,
  // their loads issued together.
  for (uint32_t base = 0; base < n_ent; base += 4u * BLOCK) {
    uint32_t len[4], cnt[4], first[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t i = base + (uint32_t)u * BLOCK + (uint32_t)tid;
      len[u] = 0;
      cnt[u] = 0;
      first[u] = 0;
      if (i < n_ent) {
        const Entry<W> en = sc.ent[i];
        len[u] = en.len;
        first[u] = (uint32_t)en.key.w[0] & 31u;
        cnt[u] = en.cnt;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t i = base + (uint32_t)u * BLOCK + (uint32_t)tid;
      if (i < n_ent) {
        sc.ent_holder[i] = (uint8_t)(len[u] ? first[u] : 0xFFu);
        if (len[u]) priv[first[u] * 32u + lane] += cnt[u];
      }
    }
  }
  {
    // the observed ballots start from their first preferences, 16 holder bytes per copy
    const uint32_t n_vec = n_obs / 16u;
    const uint4 *src = reinterpret_cast<const uint4 *>(a.obs_first);
    uint4 *dst = reinterpret_cast<uint4 *>(sc.obs_holder);
    for (uint32_t v = tid; v < n_vec; v += BLOCK) dst[v] = src[v];
    for (uint32_t i = n_vec * 16u + (uint32_t)tid; i < n_obs; i += BLOCK) sc.obs_holder[i] = a.obs_first[i];
  }
  flush_private_tallies(priv, sm.tally, nc);
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
    // redistribution (irv_ballot.cpp:109-133): scan the holder bytes, 2 x 16 per thread and step; only the entries
    // held by the eliminated candidate are re-examined, each lane going through its own
    const uint32_t pattern = chosen * 0x01010101u;
    for (int part = 0; part < 2; ++part) {
      uint8_t *holder = part == 0 ? sc.ent_holder : sc.obs_holder;
      const uint32_t n_part = part == 0 ? n_ent : n_obs;
      const uint32_t n_vec = (n_part + 15u) / 16u;
      for (uint32_t vbase = 0; vbase < n_vec; vbase += 2u * BLOCK) {
        uint4 h[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint32_t v = vbase + (uint32_t)u * BLOCK + (uint32_t)tid;
          h[u] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
          if (v < n_vec) h[u] = reinterpret_cast<const uint4 *>(holder)[v];
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint32_t v = vbase + (uint32_t)u * BLOCK + (uint32_t)tid;
          const uint32_t hw[4] = {h[u].x, h[u].y, h[u].z, h[u].w};
          uint32_t hits = 0;  // bit j: byte j of this lane's 16 equals `chosen`
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t eq = __vcmpeq4(hw[q], pattern);  // 0xFF per equal byte
            hits |= ((eq & 1u) | ((eq >> 7) & 2u) | ((eq >> 14) & 4u) | ((eq >> 21) & 8u)) << (4 * q);
          }
          while (hits) {
            const uint32_t j = (uint32_t)__ffs(hits) - 1u;
            hits &= hits - 1u;
            const uint32_t i = v * 16u + j;
            if (i >= n_part) break;
            uint32_t cand, cnt;
            if (part == 0) {
              const Entry<W> en = sc.ent[i];
              cand = first_standing<W>(en.key, en.len, eliminated);
              cnt = en.cnt;
            } else {
              // an observed key stores at most nc-1 preferences; its length is implied by repetition
              cand = first_standing<W>(obs_key[i], nc - 1, eliminated);
              cnt = a.obs_cnt[i];
            }
            holder[i] = (uint8_t)cand;
            if (cand != 0xFFu) priv[cand * 32u + lane] += cnt;
            stat[kStatRereads] += 1;
          }
        }
      }
    }
    flush_private_tallies(priv, sm.tally, nc);
    __syncthreads();
  }
}

// Thread 0, after irv_election: the election's order / winners (R_tree.cpp:245-253 reads the last n_winners).
template <int BLOCK>
__device__ __forceinline__ void record_outcome(const SimArgs &a, SharedFixed<BLOCK> &sm, uint32_t e) {
  const uint32_t nc = a.P.nc;
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

// Per-thread counters into the call's statistics, one atomic per warp and counter.
__device__ __forceinline__ void flush_stats(const SimArgs &a, uint32_t *stat) {
#ifdef DTB_PHASE_CLOCKS
  stat[kStatGammas] = stat[kStatCount + 0];
  stat[kStatBinomials] = stat[kStatCount + 1];
  stat[kStatSlots] = stat[kStatCount + 2];
  stat[kStatRereads] = stat[kStatCount + 3];
  stat[kStatEntries] = stat[kStatCount + 4];   // frames with one ballot
  stat[kStatItems] = stat[kStatCount + 5];     // frames with 2-3 ballots
  stat[kStatUrnItems] = stat[kStatCount + 7];  // walker passes
#endif
  if (a.stats) {
#pragma unroll
    for (int i = 0; i < kStatCount; ++i) {
      uint32_t s = __reduce_add_sync(0xffffffffu, stat[i]);
      if ((threadIdx.x & 31) == 0 && s) atomicAdd(&a.stats[i], (unsigned long long)s);
    }
  }
}

#ifndef DTB_SIM_THREADS_PER_SM
#define DTB_SIM_THREADS_PER_SM 512  // 128 registers per thread
#endif
template <int W, int BLOCK, bool LOG>
__global__ void __launch_bounds__(BLOCK, DTB_SIM_THREADS_PER_SM / BLOCK) simulate_kernel(const __grid_constant__ SimArgs a) {
  __shared__ SharedFixed<BLOCK> sm;
  __shared__ WarpTile<W> tiles[BLOCK / 32];
  extern __shared__ __align__(16) uint8_t dyn_smem[];  // the walkers' single-ballot stacks (simulate_smem_bytes)
  const int tid = threadIdx.x;
  const uint32_t nc = a.P.nc;
  Scratch<W> sc;
  sc.bind(a.scratch + (uint64_t)blockIdx.x * a.scratch_stride, a.cap_items, a.cap_entries, a.cap_small);
  uint32_t stat[kStatCount + kClkSlots];
#pragma unroll
  for (int i = 0; i < kStatCount + kClkSlots; ++i) stat[i] = 0;
  if (tid < (int)kMaxCandidates) sm.wins[tid] = 0;

  for (;;) {
    __syncthreads();
    if (tid == 0) sm.election = atomicAdd(a.work_counter, 1u);
    __syncthreads();
    const uint32_t e = sm.election;
    if (e >= a.n_elections) break;
    const RngKey key = election_key(a.seed, a.first_election + e);

    expand_big_levels<W, BLOCK, LOG>(a, sc, sm, tiles, key, stat);
    walk_election<W, BLOCK, WalkBudget<W>::kFused>(
        a, sc, sm, reinterpret_cast<Item<W> *>(&tiles[tid >> 5]),
        reinterpret_cast<Item<W> *>(dyn_smem) + (uint32_t)(tid >> 5) * WalkBudget<W>::kFused, key, stat);
    if (tid == 0) {
      stat[kStatEntries] += sm.n_entries;
      if (a.entry_count_out) *a.entry_count_out = sm.n_entries;
    }
    if (!a.run_irv) continue;

    // With the full order requested run nc-1 rounds (the survivor is appended); otherwise stop
    // when n_winners candidates are left (R_tree.cpp:245-253 only reads the last n_winners).
    const uint32_t n_rounds = a.elim_orders ? nc - 1 : nc - a.n_winners;
    DTB_CLK_BEGIN(t_irv);
    irv_election<W, BLOCK>(a, sc, sm, tiles[tid >> 5].val, key, n_rounds, stat);
    DTB_CLK_END(t_irv, 3);
    if (tid == 0) record_outcome<BLOCK>(a, sm, e);
  }
  __syncthreads();
  // (d) flush: one atomic per candidate per block
  if (a.run_irv && tid < (int)nc && sm.wins[tid]) atomicAdd(&a.win_counts[tid], (unsigned long long)sm.wins[tid]);
  flush_stats(a, stat);
}

// ---- the same election in three launches ---------------------------------------------------------------------------
// The level loop needs 128 registers per thread (Marsaglia-Tsang, BTRS), the walkers and the IRV stage far fewer, and
// all three spend most of their time waiting on dependent loads and dependent arithmetic: halving the resident warps
// more than halves the rate. As separate kernels over a wave of elections - block b works on election b of the wave in
// scratch slot b, the header words of the slot carrying the counts from one launch to the next - the walkers and the IRV
// stage run with two to three times the warps per SM. Same device functions, same draws, same outcomes.
#ifndef DTB_BIG_THREADS_PER_SM
#define DTB_BIG_THREADS_PER_SM 1024  // 64 registers per thread: some spills, but twice the warps of the 128-register build (26.5 -> 20.6 ms per wave)
#endif
template <int W, int BLOCK, bool LOG>
__global__ void __launch_bounds__(BLOCK, DTB_BIG_THREADS_PER_SM / BLOCK) sim_big_kernel(const __grid_constant__ SimArgs a) {
  __shared__ SharedFixed<BLOCK> sm;
  __shared__ WarpTile<W> tiles[BLOCK / 32];
  const int tid = threadIdx.x;
  const uint32_t e = blockIdx.x;
  Scratch<W> sc;
  sc.bind(a.scratch + (uint64_t)e * a.scratch_stride, a.cap_items, a.cap_entries, a.cap_small);
  uint32_t stat[kStatCount + kClkSlots];
#pragma unroll
  for (int i = 0; i < kStatCount + kClkSlots; ++i) stat[i] = 0;
  expand_big_levels<W, BLOCK, LOG>(a, sc, sm, tiles, election_key(a.seed, a.first_election + e), stat);
  if (tid == 0) {
    sc.hdr[kHdrSmall] = sm.out_small;
    sc.hdr[kHdrEntries] = sm.n_entries;
    sc.hdr[kHdrOverflow] = sm.overflow;
    if (sm.overflow) *a.error_flag = 1;
  }
  flush_stats(a, stat);
}

#ifndef DTB_WALK_THREADS_PER_SM
#define DTB_WALK_THREADS_PER_SM 1024
#endif
template <int W, int BLOCK>
__global__ void __launch_bounds__(BLOCK, DTB_WALK_THREADS_PER_SM / BLOCK) sim_walk_kernel(const __grid_constant__ SimArgs a) {
  __shared__ SharedFixed<BLOCK> sm;
  extern __shared__ __align__(16) uint8_t dyn_smem[];  // per warp: kPhased + kPhased / 2 frames
  constexpr uint32_t B = WalkBudget<W>::kPhased;
  const int tid = threadIdx.x;
  const uint32_t e = blockIdx.x;
  Scratch<W> sc;
  sc.bind(a.scratch + (uint64_t)e * a.scratch_stride, a.cap_items, a.cap_entries, a.cap_small);
  uint32_t stat[kStatCount + kClkSlots];
#pragma unroll
  for (int i = 0; i < kStatCount + kClkSlots; ++i) stat[i] = 0;
  if (tid == 0) {
    sm.out_small = sc.hdr[kHdrSmall];
    sm.n_entries = sc.hdr[kHdrEntries];
    sm.overflow = sc.hdr[kHdrOverflow];
  }
  __syncthreads();
  Item<W> *mine = reinterpret_cast<Item<W> *>(dyn_smem) + (uint32_t)(tid >> 5) * (B + B / 2u);
  walk_election<W, BLOCK, B>(a, sc, sm, mine + B, mine, election_key(a.seed, a.first_election + e), stat);
  if (tid == 0) {
    sc.hdr[kHdrEntries] = sm.n_entries;
    stat[kStatEntries] += sm.n_entries;
  }
  flush_stats(a, stat);
}

#ifndef DTB_IRV_THREADS_PER_SM
#define DTB_IRV_THREADS_PER_SM 1536
#endif
template <int W, int BLOCK>
__global__ void __launch_bounds__(BLOCK, DTB_IRV_THREADS_PER_SM / BLOCK) sim_irv_kernel(const __grid_constant__ SimArgs a) {
  __shared__ SharedFixed<BLOCK> sm;
  extern __shared__ __align__(16) uint8_t dyn_smem[];  // kPrivTallyWords words per warp
  uint32_t *priv = reinterpret_cast<uint32_t *>(dyn_smem) + (uint32_t)(threadIdx.x >> 5) * kPrivTallyWords;
  const int tid = threadIdx.x;
  const uint32_t nc = a.P.nc;
  const uint32_t e = blockIdx.x;
  Scratch<W> sc;
  sc.bind(a.scratch + (uint64_t)e * a.scratch_stride, a.cap_items, a.cap_entries, a.cap_small);
  uint32_t stat[kStatCount + kClkSlots];
#pragma unroll
  for (int i = 0; i < kStatCount + kClkSlots; ++i) stat[i] = 0;
  if (tid < (int)kMaxCandidates) sm.wins[tid] = 0;
  if (tid == 0) sm.n_entries = sc.hdr[kHdrEntries];
  __syncthreads();
  const uint32_t n_rounds = a.elim_orders ? nc - 1 : nc - a.n_winners;
  irv_election<W, BLOCK>(a, sc, sm, priv, election_key(a.seed, a.first_election + e), n_rounds, stat);
  if (tid == 0) record_outcome<BLOCK>(a, sm, e);
  __syncthreads();
  if (tid < (int)nc && sm.wins[tid]) atomicAdd(&a.win_counts[tid], (unsigned long long)sm.wins[tid]);
  flush_stats(a, stat);
}

}  // namespace dtb
