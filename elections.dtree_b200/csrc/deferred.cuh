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
// typically everything below the finalists' first preferences, i.e. most of the tree — are never
// drawn, which removes most of the work of IRVNode::sample (irv_node.cpp:107-174) and all of the
// list traffic of socialChoiceIRV (irv_ballot.cpp:42-138).
//
// Observed ballots (replace == false) are not copied per election (dirichlet_tree.h:229-232): a node's
// observed counts ARE its slot counts, so they ride along with the sampled counts of each child.
//
// One warp per election; a block is a bundle of independent warps (no block barrier in the loop).
// Nodes with a small sampled count are split one lane per node; big counts go through the
// shared-memory tile path of split.cuh.
#pragma once
#include <cstdint>

#include "split.cuh"

#ifndef DTB_DEFERRED_THREADS_PER_SM
#define DTB_DEFERRED_THREADS_PER_SM 512  // resident threads per SM the register allocation aims at (128 registers each)
#endif

namespace dtb {

typedef Item<1> DItem;  // prefix unused; pad_ = depth

// Scratch capacities of the arena an election runs in (items).
struct DeferredCaps {
  uint32_t pool, queue, big;
};

struct DeferredWarp {
  uint32_t tally[kMaxCandidates];
  uint32_t q_head;    // mirror of the drain loop's head, for the queue-occupancy check
  uint32_t n_pool;    // high-water mark of the parked-node pool
  int32_t n_free;     // free-list depth (may dip below 0 while lanes race for the last free slots)
  uint32_t q_tail;    // small-count work queue (circular, monotone counters; q_head is a register)
  uint32_t q1_head, q1_tail;  // the same for the queue of single-ballot nodes
  uint32_t big_tail;  // big-count list of waiting nodes (big[0])
  uint32_t overflow;
  uint32_t unresolved;      // ballots (sampled + observed) inside the waiting nodes
  uint32_t small_mass_max;  // largest number of ballots any node in the small-count queue has had
  uint32_t rec_w;           // node records written by route_child (traffic accounting)
  // Ballots inside the waiting nodes by weight class: hist[b] = sum of the masses of the waiting nodes with
  // 2^b <= mass < 2^(b+1); big_cnt[b] = how many of them sit in the big-count list. They let a resolve pass pick
  // the highest threshold that still leaves too little waiting to matter, instead of a worst-case one.
  uint32_t hist[32];
  uint32_t big_cnt[32];
};

__device__ __forceinline__ uint32_t mass_class(uint32_t m) { return 31u - (uint32_t)__clz(m | 1u); }

// A waiting or parked node carries the number of ballots inside it in the (otherwise unused) prefix word.
__device__ __forceinline__ uint32_t node_mass(const DItem &x) { return (uint32_t)x.prefix.w[0]; }

// Per-warp scratch. Parked nodes live in `pool` (slot reuse through `free_list`), with the candidate they
// are parked under in `holder` (0xFF = free slot). Nodes being split in the current round are transient:
// small counts go through the circular `queue`, big counts through the ping-pong `big` lists.
// Capacities are exact bounds: a sampled ballot is in at most one live node, and a node carrying only
// observed ballots is one of the tree's nodes.
struct DeferredScratch {
  DItem *pool;
  uint8_t *holder;
  uint32_t *free_list;
  DItem *queue;   // small sampled counts, more than one ballot inside
  DItem *queue1;  // nodes holding exactly one ballot: the bulk of what waits deep in a large tree, and wanted only
                  // when a round cannot be settled otherwise, so they are kept out of the other passes' way
  DItem *big[2];
  __host__ __device__ static uint64_t a16(uint64_t x) { return (x + 15) & ~15ull; }
  __host__ __device__ static uint64_t bytes(uint32_t cap, uint32_t capq, uint32_t capb) {
    return (uint64_t)cap * sizeof(DItem) + a16(cap) + a16((uint64_t)cap * 4) + 2 * (uint64_t)capq * sizeof(DItem) +
           2 * (uint64_t)capb * sizeof(DItem);
  }
  __device__ void bind(uint8_t *p, uint32_t cap, uint32_t capq, uint32_t capb) {
    pool = (DItem *)p; p += (uint64_t)cap * sizeof(DItem);
    holder = p; p += a16(cap);
    free_list = (uint32_t *)p; p += a16((uint64_t)cap * 4);
    queue = (DItem *)p; p += (uint64_t)capq * sizeof(DItem);
    queue1 = (DItem *)p; p += (uint64_t)capq * sizeof(DItem);
    big[0] = (DItem *)p; p += (uint64_t)capb * sizeof(DItem);
    big[1] = (DItem *)p;
  }
};

// Puts a node whose candidate is out on the waiting list for its kind of split (big[0] or the circular queue).
__device__ __forceinline__ void enqueue_node(const SimArgs &a, DeferredWarp &ws, const DeferredScratch &sc,
                                             const DeferredCaps &cp, const DItem &ch) {
  const uint32_t m = node_mass(ch);
  atomicAdd(&ws.unresolved, m);
  atomicAdd(&ws.hist[mass_class(m)], m);
  if (ch.count > a.urn_max) {
    atomicAdd(&ws.big_cnt[mass_class(m)], 1u);
    const uint32_t q = coalesced_reserve(&ws.big_tail);
    if (q < cp.big) sc.big[0][q] = ch;
    else ws.overflow = 1;
  } else if (m == 1u) {
    const uint32_t q = coalesced_reserve(&ws.q1_tail);
    if (q - ws.q1_head < cp.queue) sc.queue1[q % cp.queue] = ch;
    else ws.overflow = 1;
  } else {
    atomicMax(&ws.small_mass_max, m);
    const uint32_t q = coalesced_reserve(&ws.q_tail);
    if (q - ws.q_head < cp.queue) sc.queue[q % cp.queue] = ch;  // q_head only grows: the test is conservative
    else ws.overflow = 1;
  }
}

// Routes the ballots of one child of a node that has just been split. May be called by any subset of
// lanes (divergent code). `tot` ballots now have `cand` as their next preference; `cs` of them are
// sampled ballots that continue below the child, `child` is the child's CSR node (-1: none).
static __device__ __noinline__ void route_child(const SimArgs &a, DeferredWarp &ws, const DeferredScratch &sc,
                                            const DeferredCaps &cp, const DItem &x, uint32_t depth, uint32_t eliminated, uint32_t cand,
                                            uint32_t tot, uint32_t cs, int32_t child, uint32_t obs) {
  const bool standing = !(eliminated >> cand & 1u);
  if (standing) atomicAdd(&ws.tally[cand], tot);
  const bool cont = cs > 0u || (obs > 0u && child >= 0);
  if (!cont) return;  // these ballots end with `cand`: they exhaust when it is eliminated
  atomicAdd(&ws.rec_w, 1u);
  DItem ch;
  ch.prefix.w[0] = cs + (child >= 0 ? obs : 0u);  // sampled ballots that go on + observed ones below
  ch.node_key = child_node_key(x.node_key, cand);
  ch.count = cs;
  ch.node = child;
  ch.used = x.used | (1u << cand);
  ch.pad_ = depth + 1u;
  if (!standing) {  // its candidate is already out: it waits to be split
    enqueue_node(a, ws, sc, cp, ch);
    return;
  }
  // park it under `cand`: reuse a freed pool slot if there is one, else grow the pool
  const uint32_t m = __activemask();
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  const int k = __popc(m), rank = __popc(m & ((1u << lane) - 1u));
  int old_free = 0;
  if (lane == leader) old_free = atomicSub(&ws.n_free, k);
  old_free = __shfl_sync(m, old_free, leader);
  const int from_free = max(0, min(k, old_free));  // lanes of rank < from_free get a recycled slot
  uint32_t grow_base = 0;
  if (lane == leader && k > from_free) grow_base = atomicAdd(&ws.n_pool, (uint32_t)(k - from_free));
  grow_base = __shfl_sync(m, grow_base, leader);
  uint32_t j;
  if (rank < from_free) j = sc.free_list[old_free - 1 - rank];
  else j = grow_base + (uint32_t)(rank - from_free);
  if (j >= cp.pool) {
    ws.overflow = 1;
    return;
  }
  sc.pool[j] = ch;
  sc.holder[j] = (uint8_t)cand;
}

// One lane splits one node with a small sampled count (0..urn_max) and routes its children.
__device__ __forceinline__ void split_small_node(const SimArgs &a, DeferredWarp &ws, const DeferredScratch &sc,
                                                 const DeferredCaps &cp, const DItem &x, uint32_t eliminated, const RngKey key,
                                                 uint32_t *stat) {
  const uint32_t nc = a.P.nc;
  const uint32_t depth = x.pad_ & 0xFFu;
  const LevelInfo lv = level_info(a.P, depth);
  const uint32_t nb = x.node >= 0 ? a.T.node_base[x.node] : 0xFFFFFFFFu;
  stat[kStatItems] += 1;
  stat[kStatUrnItems] += 1;
  const uint64_t picks = x.count ? urn_picks(a, lv, x.node, nb, x.count, key, x.node_key) : 0ull;
  if (nb != 0xFFFFFFFFu && a.use_obs) {
    // initialised node carrying observed ballots: every child slot may have something to route
    uint64_t c4[2] = {0ull, 0ull};  // 4-bit count per slot (count <= 8)
#pragma unroll 1
    for (uint32_t b = 0; b < x.count; ++b) {
      const uint32_t s = (uint32_t)(picks >> (8u * b)) & 0xFFu;
      if (s == lv.K) continue;  // the ballot stops here: exhausted
      c4[s >> 4] += 1ull << (4u * (s & 15u));
    }
    stat[kStatSlots] += lv.d;
#pragma unroll 1
    for (uint32_t slot = 0; slot < lv.K; ++slot) {
      const uint32_t c = (uint32_t)(c4[slot >> 4] >> (4u * (slot & 15u))) & 15u;
      const uint32_t obs = a.T.slot_count[nb + slot];
      if (c + obs == 0u) continue;
      route_child(a, ws, sc, cp, x, depth, eliminated, a.T.slot_cand[nb + slot], c + obs, lv.leaf_children ? 0u : c,
                  a.T.slot_child[nb + slot], obs);
    }
  } else {
    // prior-only subtree (or sampling with replacement): only the drawn outcomes matter
    uint64_t c4[3] = {0ull, 0ull, 0ull};  // balls per outcome, 4 bits each; cleared when the outcome is routed
#pragma unroll 1
    for (uint32_t b = 0; b < x.count; ++b) {
      const uint32_t s = (uint32_t)(picks >> (8u * b)) & 0xFFu;
      c4[s >> 4] += 1ull << (4u * (s & 15u));
    }
#pragma unroll 1
    for (uint32_t b = 0; b < x.count; ++b) {
      const uint32_t slot = (uint32_t)(picks >> (8u * b)) & 0xFFu;
      const uint32_t c = (uint32_t)(c4[slot >> 4] >> (4u * (slot & 15u))) & 15u;
      if (c == 0u) continue;
      c4[slot >> 4] &= ~(15ull << (4u * (slot & 15u)));
      if (slot == lv.K) continue;  // the ballots stop here: exhausted
      const uint32_t cand = nb != 0xFFFFFFFFu ? (uint32_t)a.T.slot_cand[nb + slot] : nth_unused(x.used, nc, slot);
      const int32_t child = nb != 0xFFFFFFFFu ? a.T.slot_child[nb + slot] : -1;
      route_child(a, ws, sc, cp, x, depth, eliminated, cand, c, lv.leaf_children ? 0u : c, child, 0u);
    }
  }
}

// The warp splits a list of nodes with big sampled counts through the shared-memory tile.
template <bool LOG>
__device__ void split_big_nodes(const SimArgs &a, WarpTile<1> &wt, DeferredWarp &ws, const DeferredScratch &sc,
                                const DeferredCaps &cp, const DItem *list, uint32_t n_list, uint32_t eliminated, int log_dp_max,
                                const RngKey key, uint32_t *stat) {
  const int lane = threadIdx.x & 31;
  const uint32_t nc = a.P.nc;
  const uint32_t dp_max = 1u << log_dp_max, stride = 2u * dp_max + 1u;
  const uint32_t ci = min((uint32_t)Chunk<1>::kMaxItems, (uint32_t)Chunk<1>::kTileWords / stride);
  for (uint32_t start = 0; start < n_list; start += ci) {
    const uint32_t m = min(ci, n_list - start);
    for (uint32_t it = lane; it < m; it += 32) {
      DItem x = list[start + it];
      wt.item[it] = x;
      wt.nbase[it] = x.node >= 0 ? a.T.node_base[x.node] : 0xFFFFFFFFu;
      stat[kStatItems] += 1;
      if (x.node >= 0) stat[kStatSlots] += level_info(a.P, x.pad_ & 0xFFu).d;
    }
    __syncwarp();
    split_items<1, LOG>(a, wt, m, log_dp_max, key, stat);
    for (uint32_t p = lane; p < (m << log_dp_max); p += 32) {
      const uint32_t it = p >> log_dp_max, slot = p & (dp_max - 1u);
      const DItem &x = wt.item[it];
      const uint32_t depth = x.pad_ & 0xFFu;
      const LevelInfo lv = level_info(a.P, depth);
      if (slot >= lv.K) continue;  // padding, or the stop outcome (its ballots are exhausted)
      const uint32_t nb = wt.nbase[it];
      const uint32_t c = wt.val[it * stride + (1u << tree_log2(lv.d)) + slot];
      uint32_t obs = 0, cand;
      int32_t child = -1;
      if (nb != 0xFFFFFFFFu) {
        if (a.use_obs) obs = a.T.slot_count[nb + slot];
        cand = a.T.slot_cand[nb + slot];
        child = a.T.slot_child[nb + slot];
      } else {
        cand = nth_unused(x.used, nc, slot);
      }
      if (c + obs == 0u) continue;
      route_child(a, ws, sc, cp, x, depth, eliminated, cand, c + obs, lv.leaf_children ? 0u : c, child, obs);
    }
    __syncwarp();
  }
}

// Splits every waiting node that holds at least `tau` ballots; lighter ones keep waiting. Children whose candidate
// is out join the waiting lists (and wait for the next call).
template <bool LOG>
__device__ void deferred_resolve(const SimArgs &a, WarpTile<1> &wt, DeferredWarp &ws, const DeferredScratch &sc,
                                 const DeferredCaps &cp, uint32_t &q_head, uint32_t &q1_head, uint32_t &hw_queue,
                                 uint32_t &hw_big, uint32_t eliminated, uint32_t tau, int log_dp_max, const RngKey key,
                                 uint32_t *stat) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  // single-ballot nodes: only when every waiting ballot counts (tau == 1); all of them are split, none goes back
  const uint32_t n_one = ws.q1_tail - q1_head;
  hw_queue = max(hw_queue, n_one);
  if (n_one && tau <= 1u) {
    for (uint32_t done = 0; done < n_one && !ws.overflow; done += 32u) {
      const uint32_t m = min(32u, n_one - done);
      DItem x;
      if ((uint32_t)lane < m) x = sc.queue1[(q1_head + lane) % cp.queue];
      q1_head += m;
      __syncwarp();  // the batch has been read: its queue slots may be reused from here on
      if (lane == 0) ws.q1_head = q1_head;
      __syncwarp();
      if ((uint32_t)lane < m) {
        stat[kStatRereads] += 1;
        atomicSub(&ws.unresolved, 1u);
        atomicSub(&ws.hist[0], 1u);
        // (walking a single ballot's chain on in this lane until a standing candidate takes it was measured again in
        // round 2: 53.6 k -> 89 k splits per C4 election, 36 k -> 30 k elections/s. One hop per waiting ballot usually
        // settles the round; the hops below it are never needed. Also measured and dropped: listing the children of the
        // batch's nodes in shared memory and routing them cooperatively, 32 at a time, with one reservation per
        // destination — bit-identical, but batches are too small to pay for the second phase: 36 k -> 24 k.)
        split_small_node(a, ws, sc, cp, x, eliminated, key, stat);
      }
      __syncwarp();
    }
  }
  __syncwarp();
  // small counts: one turn of the circular queue, 32 nodes at a time, one lane each
  const uint32_t n_small = ws.q_tail - q_head;
  hw_queue = max(hw_queue, n_small);
  if (n_small && ws.small_mass_max >= tau) {
    for (uint32_t done = 0; done < n_small && !ws.overflow; done += 32u) {
      const uint32_t m = min(32u, n_small - done);
      DItem x;
      if ((uint32_t)lane < m) x = sc.queue[(q_head + lane) % cp.queue];
      q_head += m;
      __syncwarp();  // the batch has been read: its queue slots may be reused from here on
      if (lane == 0) ws.q_head = q_head;
      __syncwarp();
      if ((uint32_t)lane < m) {
        const uint32_t mass = node_mass(x);
        stat[kStatRereads] += 1;
        if (mass >= tau) {
          atomicSub(&ws.unresolved, mass);
          atomicSub(&ws.hist[mass_class(mass)], mass);
          split_small_node(a, ws, sc, cp, x, eliminated, key, stat);
        } else {  // back to the end of the queue
          stat[kStatEntries] += 1;
          const uint32_t q = coalesced_reserve(&ws.q_tail);
          if (q - ws.q_head < cp.queue) sc.queue[q % cp.queue] = x;
          else ws.overflow = 1;
        }
      }
      __syncwarp();
    }
  }
  __syncwarp();
  // big counts: the heavy ones move to big[1] and go through the tile, the rest are compacted in big[0]
  const uint32_t n_big = min(ws.big_tail, cp.big);
  hw_big = max(hw_big, n_big);
  if (n_big == 0u || ws.overflow) return;
  {
    // nothing in the big-count list is heavy enough: leave it alone (tau is a power of two or 1)
    const uint32_t cls = mass_class(tau);
    const uint32_t heavy_big = __reduce_add_sync(0xffffffffu, (uint32_t)lane >= cls ? ws.big_cnt[lane] : 0u);
    if (heavy_big == 0u) return;
  }
  uint32_t n_keep = 0, n_work = 0, work_mass = 0;
  for (uint32_t start = 0; start < n_big; start += 32u) {
    const uint32_t i = start + lane;
    DItem x;
    bool heavy = false, keep = false;
    if (i < n_big) {
      x = sc.big[0][i];
      heavy = node_mass(x) >= tau;
      keep = !heavy;
      stat[kStatRereads] += 1;  // read here, written to its new place below
      stat[kStatEntries] += 1;
    }
    const uint32_t hm = __ballot_sync(0xffffffffu, heavy), km = __ballot_sync(0xffffffffu, keep);
    const uint32_t below = (1u << lane) - 1u;
    __syncwarp();  // every lane has read its node before any slot of big[0] is rewritten
    if (heavy) {
      sc.big[1][n_work + __popc(hm & below)] = x;
      work_mass += node_mass(x);
      atomicSub(&ws.hist[mass_class(node_mass(x))], node_mass(x));
      atomicSub(&ws.big_cnt[mass_class(node_mass(x))], 1u);
    }
    if (keep) sc.big[0][n_keep + __popc(km & below)] = x;
    n_work += __popc(hm);
    n_keep += __popc(km);
  }
  work_mass = __reduce_add_sync(0xffffffffu, work_mass);
  __syncwarp();
  if (lane == 0) {
    ws.big_tail = n_keep;
    ws.unresolved -= work_mass;
  }
  __syncwarp();
  if (n_work) split_big_nodes<LOG>(a, wt, ws, sc, cp, sc.big[1], n_work, eliminated, log_dp_max, key, stat);
  __syncwarp();
}

// One election in the arena (sc, cp). Returns false if the arena was too small (nothing is recorded then).
// On success *eliminated_out / *order_out / *ties_out describe the outcome (lane r holds round r's loser).
//
// Nodes released by an elimination are not split as they appear: they wait, their ballots counted in
// ws.unresolved, and a round is decided as soon as its outcome cannot depend on where those ballots go. They can
// only add to tallies, so the weakest candidate stays weakest if its tally plus all unresolved ballots is still
// below every other tally, and a leader keeps a majority if it has one with all of them counted against it. While
// neither holds, the waiting nodes heavy enough to matter are split (deferred_resolve) and the test is repeated.
// Every split is a pure function of the node, so which ones are carried out changes no outcome: elimination
// orders and winners are those of the eager evaluation (simulate_kernel).
template <bool LOG>
__device__ bool deferred_election(const SimArgs &a, WarpTile<1> &wt, DeferredWarp &ws, const DeferredScratch &sc,
                                  const DeferredCaps &cp, const RngKey key, uint32_t n_rounds, int log_dp_max,
                                  uint32_t *usage, uint32_t &eliminated_out, uint32_t &order_out, uint32_t &ties_out,
                                  uint32_t *stat) {
  const int lane = threadIdx.x & 31;
  const uint32_t nc = a.P.nc;
  ws.tally[lane] = 0;
  ws.hist[lane] = 0;
  ws.big_cnt[lane] = 0;
  __syncwarp();
  if (lane == 0) {
    DItem root;
    root.prefix.w[0] = max(a.n_sample, 1u);  // any positive number: nothing can be decided before the root is split
    root.node_key = kRootNodeKey;
    root.count = a.n_sample;
    root.node = 0;
    root.used = 0;
    root.pad_ = 0;
    ws.n_pool = 0;
    ws.n_free = 0;
    ws.overflow = 0;
    ws.q_tail = 0;
    ws.q_head = 0;
    ws.q1_tail = 0;
    ws.q1_head = 0;
    ws.big_tail = 0;
    ws.unresolved = (uint32_t)root.prefix.w[0];
    ws.small_mass_max = 0;
    ws.rec_w = 1;
    ws.hist[mass_class((uint32_t)root.prefix.w[0])] = (uint32_t)root.prefix.w[0];
    if (a.n_sample > a.urn_max) {
      sc.big[0][0] = root;
      ws.big_tail = 1;
      ws.big_cnt[mass_class((uint32_t)root.prefix.w[0])] = 1;
    } else {  // (the queue of nodes with a small count; its mass is "positive", not a ballot count, at the root)
      sc.queue[0] = root;
      ws.q_tail = 1;
      ws.small_mass_max = (uint32_t)root.prefix.w[0];
    }
  }
  __syncwarp();
  uint32_t eliminated = 0, tie_rounds = 0, my_order = 0xFFu, q_head = 0, q1_head = 0, hw_queue = 0, hw_big = 0;
  const bool early_stop = a.n_winners == 1u && a.elim_orders == nullptr;

  for (uint32_t round = 0; round < n_rounds && !ws.overflow; ++round) {
    // candidates with the minimal tally among those standing (irv_ballot.cpp:88-98), one lane each
    const bool standing = lane < (int)nc && !(eliminated >> lane & 1u);
    uint32_t mn, tied, ntied;
    bool won = false;
    for (;;) {
      __syncwarp();
      const uint32_t u = ws.unresolved;
      const uint32_t mine = standing ? ws.tally[lane] : 0u;
      const uint32_t total = __reduce_add_sync(0xffffffffu, mine), top = __reduce_max_sync(0xffffffffu, mine);
      const uint32_t t = standing ? mine : 0xFFFFFFFFu;
      mn = __reduce_min_sync(0xffffffffu, t);
      tied = __ballot_sync(0xffffffffu, standing && t == mn);
      ntied = (uint32_t)__popc(tied);
      const uint32_t second = ntied > 1u ? mn : __reduce_min_sync(0xffffffffu, (standing && t != mn) ? t : 0xFFFFFFFFu);
      // A candidate holding more than half of the ballots still in play can never have the minimal tally
      // again (tallies only grow, ballots in play only shrink), so it is the winner: stop here.
      const uint32_t lead = top > total - top ? top - (total - top) : 0u;  // its margin over everybody else together
      if (early_stop && lead > u) {
        const uint32_t who = __ballot_sync(0xffffffffu, standing && mine == top);
        eliminated = ~who & (nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u));
        won = true;
        break;
      }
      const uint32_t gap = second - mn;  // 0 when tied
      if (u == 0u || gap > u) break;     // the weakest candidate is known
      if (ws.overflow) break;
      // Split what can matter. The round is settled once fewer than `need` ballots wait; nodes lighter than 2^b hold
      // hist[0] + ... + hist[b-1] ballots together, so every node of weight >= 2^b is split for the largest b whose
      // lighter classes hold at most need * tau_quarters / 4 ballots (the rest of the allowance is for children of the
      // split nodes that have to wait in turn). With nothing to spare (need <= 1, e.g. a tie) everything is split.
      const uint32_t need = early_stop ? max(gap, lead) : gap;
      const uint32_t allow = (uint32_t)(((uint64_t)need * a.tau_quarters) >> 2);
      uint32_t below = ws.hist[lane];  // inclusive prefix sum over the weight classes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, below, o);
        if (lane >= o) below += y;
      }
      // class b qualifies as the threshold class if the classes below it (prefix of b-1) fit the allowance
      const uint32_t lighter = __shfl_up_sync(0xffffffffu, below, 1);
      const uint32_t fits = __ballot_sync(0xffffffffu, lane == 0 || lighter <= allow);
      const uint32_t cls = 31u - (uint32_t)__clz(fits);  // bit 0 is always set
      const uint32_t tau = cls == 0u ? 1u : (1u << cls);
      deferred_resolve<LOG>(a, wt, ws, sc, cp, q_head, q1_head, hw_queue, hw_big, eliminated, tau, log_dp_max, key, stat);
    }
    if (won || ws.overflow) break;
    if (ntied > 1u) {  // uniform among the tied, ascending index (irv_ballot.cpp:100-102)
      U4 r = node_random(key, 0ull, kStreamTie, round);
      const uint32_t pick = uniform_below(r.x, ntied);
      for (uint32_t k = 0; k < pick; ++k) tied &= tied - 1u;
      tie_rounds += 1;
    }
    const uint32_t chosen = (uint32_t)__ffs(tied) - 1u;
    if (lane == (int)round) my_order = chosen;
    eliminated |= 1u << chosen;
    if (mn == 0u) continue;  // nobody counts for this candidate: nothing is parked under it
    // After the last elimination nobody reads the tallies again (R_tree.cpp:245-253 only needs who is left), so
    // the loser's ballots are not redistributed. The reference does redistribute them (irv_ballot.cpp:109-133) -
    // typically the runner-up's whole subtree - without any effect on its output.
    if (round + 1u == n_rounds) continue;
    // release the nodes parked under the eliminated candidate: scan the holder bytes 16 per lane
    if (lane == 0 && ws.n_free < 0) ws.n_free = 0;
    __syncwarp();
    const uint32_t np = min(ws.n_pool, cp.pool);
    const uint32_t n_vec = (np + 15u) / 16u;
    const uint32_t pattern = chosen * 0x01010101u;
    for (uint32_t vbase = 0; vbase < n_vec; vbase += 32) {
      const uint32_t v = vbase + lane;
      uint32_t hits = 0;
      if (v < n_vec) {
        const uint4 h = reinterpret_cast<const uint4 *>(sc.holder)[v];
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t eq = __vcmpeq4(hw[q], pattern);
          hits |= ((eq & 1u) | ((eq >> 7) & 2u) | ((eq >> 14) & 4u) | ((eq >> 21) & 8u)) << (4 * q);
        }
      }
      while (hits) {
        const uint32_t j = v * 16u + (uint32_t)__ffs(hits) - 1u;
        hits &= hits - 1u;
        if (j >= np) break;
        sc.holder[j] = 0xFF;
        const DItem x = sc.pool[j];
        stat[kStatRereads] += 1;
        stat[kStatEntries] += 1;
        enqueue_node(a, ws, sc, cp, x);
        const uint32_t f = coalesced_reserve(reinterpret_cast<uint32_t *>(&ws.n_free));
        sc.free_list[f] = j;
      }
    }
    __syncwarp();
  }
  __syncwarp();
  if (lane == 0) stat[kStatEntries] += ws.rec_w;
  usage[0] = max(usage[0], min(ws.n_pool, 0x7FFFFFFFu));
  usage[1] = max(usage[1], max(hw_queue, max(ws.q_tail - q_head, ws.q1_tail - q1_head)));
  usage[2] = max(usage[2], max(hw_big, min(ws.big_tail, cp.big)));
  eliminated_out = eliminated;
  order_out = my_order;
  ties_out = tie_rounds;
  return ws.overflow == 0;
}

template <int BLOCK, bool LOG>
__global__ void __launch_bounds__(BLOCK, DTB_DEFERRED_THREADS_PER_SM / BLOCK) deferred_kernel(const __grid_constant__ SimArgs a) {
  __shared__ WarpTile<1> tiles[BLOCK / 32];
  __shared__ DeferredWarp wstate[BLOCK / 32];
  __shared__ uint32_t s_wins[kMaxCandidates];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t nc = a.P.nc;
  const uint32_t cand_mask = nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u);
  WarpTile<1> &wt = tiles[warp];
  DeferredWarp &ws = wstate[warp];
  const uint32_t gwarp = blockIdx.x * (BLOCK / 32) + warp;
  // Each warp owns a right-sized arena; elections that outgrow it are re-run in one of a few
  // worst-case arenas shared by the whole grid (try-lock; holders never wait, so no deadlock).
  DeferredCaps cp{a.cap_items, a.cap_queue, a.cap_big};
  DeferredScratch sc;
  sc.bind(a.scratch + (uint64_t)gwarp * a.scratch_stride, cp.pool, cp.queue, cp.big);
  const DeferredCaps cp_max{a.arena_cap_items, a.arena_cap_queue, a.arena_cap_big};
  const int log_dp_max = tree_log2(level_info(a.P, 0).d);
  uint32_t stat[kStatCount], usage[3] = {0u, 0u, 0u};
#pragma unroll
  for (int i = 0; i < kStatCount; ++i) stat[i] = 0;
  if (tid < (int)kMaxCandidates) s_wins[tid] = 0;
  __syncthreads();
  const uint32_t n_rounds = a.elim_orders ? nc - 1 : nc - a.n_winners;

  const uint32_t n_total = a.n_elections_ptr ? min(*a.n_elections_ptr, a.n_elections) : a.n_elections;
  for (;;) {
    if (gwarp >= a.max_warps) break;
    uint32_t e = 0;
    if (lane == 0) e = atomicAdd(a.work_counter, 1u);
    e = __shfl_sync(0xffffffffu, e, 0);
    if (e >= n_total) break;
    if (a.election_list) e = a.election_list[e];
    const RngKey key = election_key(a.seed, a.first_election + e);
    uint32_t eliminated = 0, my_order = 0xFFu, tie_rounds = 0;
    // First in the warp's own arena; if that proves too small, once more in a worst-case arena.
    bool ok = false;
    for (int attempt = 0; attempt < 2 && !ok; ++attempt) {
      DeferredScratch cur_sc = sc;
      DeferredCaps cur_cp = cp;
      uint32_t slot = 0;
      if (attempt == 1) {
        if (a.n_arenas == 0u) break;
        if (lane == 0) {
          slot = gwarp % a.n_arenas;
          while (atomicCAS(&a.arena_locks[slot], 0u, 1u) != 0u) {
            slot = (slot + 1u) % a.n_arenas;
            __nanosleep(200);
          }
          __threadfence();
          atomicAdd(a.arena_runs, 1u);
        }
        slot = __shfl_sync(0xffffffffu, slot, 0);
        cur_cp = cp_max;
        cur_sc.bind(a.arena_scratch + (uint64_t)slot * a.arena_stride, cp_max.pool, cp_max.queue, cp_max.big);
      }
      ok = deferred_election<LOG>(a, wt, ws, cur_sc, cur_cp, key, n_rounds, log_dp_max, usage, eliminated, my_order,
                                  tie_rounds, stat);
      __syncwarp();
      if (attempt == 1 && lane == 0) {
        __threadfence();
        atomicExch(&a.arena_locks[slot], 0u);
      }
    }
    if (!ok && lane == 0) *a.error_flag = 2;
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
  if (a.usage && lane == 0) {
    atomicMax(&a.usage[0], usage[0]);
    atomicMax(&a.usage[1], usage[1]);
    atomicMax(&a.usage[2], usage[2]);
  }
  if (a.stats) {
#pragma unroll
    for (int i = 0; i < kStatCount; ++i) {
      uint32_t s = __reduce_add_sync(0xffffffffu, stat[i]);
      if (lane == 0 && s) atomicAdd(&a.stats[i], (unsigned long long)s);
    }
  }
}

}  // namespace dtb
