// lane.cuh — the deferred algorithm of deferred.cuh with ONE THREAD PER ELECTION.
//
// deferred_kernel gives an election a whole warp; a light election (C3: 8 node splits) leaves most of those lanes
// idle and the warp-cooperative machinery (ballots, shared-memory tiles, reservations) is overhead. Here a lane owns
// an election outright: plain scalar code, its own arena, no atomics. The 32 elections of a warp start together and
// walk nearly the same nodes in nearly the same order (the root, then the nodes parked under the weakest
// candidates...), so the lanes stay largely in step where it matters — the big Gamma / binomial splits near the top
// of the tree.
//
// Every split is the same pure function of (seed, election, node key, parameters, count) as in split.cuh: the Gamma
// leaves, the sum tree and the binomial tree below are the scalar transcription of split_items, operation for
// operation, so outcomes are bit-identical to the other two kernels
// (tests/test_gpu_parity.py::test_deferred_equals_materialised runs all three).
//
// Memory layout. The heap-indexed tree of a node being split (sums, then counts) is a column of a shared-memory tile
// ([tree node][lane]: conflict-free; the round-1 version kept it in a thread-local array that thrashed L1). The
// arenas of a warp's 32 elections are interleaved record by record (record j of lane l sits at (32 j + l) * 32
// bytes), so lanes that park or release their j-th node together touch one contiguous kilobyte (DRAM traffic per C3
// election 11.8 KB -> 6.9 KB, profiles/r02).
//
// An election that outgrows its arena, or needs more node splits than a thread should do alone
// (SimArgs::lane_max_splits), is not recorded; its index goes to a retry list that a deferred_kernel launch
// enqueued right behind replays, a warp per election.
//
// Tried and dropped (profiles/r02f, r02g): spreading the variates of the 32 elections' pending splits over the warp
// as (election, outcome) pairs with shared retry lists, so that rejections are retried at full width. It removes
// the 5-of-32-lane slow paths but executes 49 % more warp-instructions (pair bookkeeping, set-up recomputed in the
// retry passes, under-filled upper tree levels) from a larger instruction footprint: 1.63 ms against 1.2 ms.
#pragma once
#include <cstdint>

#include "deferred.cuh"

namespace dtb {

#ifndef DTB_LANE_TOP
#define DTB_LANE_TOP 16  // 8: 0.717 ms per 100 000 C3 elections, 12: 0.704, 16: 0.688, 20: 0.702, 24: 0.694
#endif
constexpr uint32_t kLaneTop = DTB_LANE_TOP;      // heavy waiting nodes per election whose ordering never touches global memory

// Per-thread bookkeeping of one election (registers).
struct LaneState {
  uint32_t n_pool;            // pool slots handed out so far (fresh ones first)
  uint32_t hw_live;           // most nodes parked at once (n_pool - n_free): what the arena has to hold (usage report)
  uint32_t n_free;            // free-list depth
  uint32_t q_head, q_tail;    // circular small-count queue (monotone counters)
  uint32_t n_big;             // waiting heavy nodes in the heap (beyond the kLaneTop kept apart)
  uint32_t n_top;             // waiting heavy nodes whose ballot counts sit in shared memory
  uint32_t eliminated;
  uint32_t hw_queue, hw_big;
  uint32_t unresolved;        // ballots (sampled + observed) inside the waiting nodes
  uint32_t rec_w, rec_r;      // 32-byte node records written to / read back from the arena (traffic accounting)
  uint32_t n_splits;          // waiting nodes split so far (the root not counted)
  bool overflow;
};

// A queued or parked node carries the number of ballots inside it in the (otherwise unused) prefix word.
__device__ __forceinline__ uint32_t lane_mass(const DItem &x) { return (uint32_t)x.prefix.w[0]; }

// The arenas of the 32 elections of a warp, interleaved per record.
struct LaneScratch {
  DItem *pool, *queue, *big;
  uint8_t *holder;
  uint32_t *free_list;
  uint32_t lane;
  __host__ __device__ static uint64_t a16(uint64_t x) { return (x + 15) & ~15ull; }
  // bytes per election
  __host__ __device__ static uint64_t bytes(uint32_t cap, uint32_t capq, uint32_t capb) {
    return (uint64_t)cap * sizeof(DItem) + a16(cap) + a16((uint64_t)cap * 4) + ((uint64_t)capq + capb) * sizeof(DItem);
  }
  __device__ void bind(uint8_t *warp_base, uint32_t cap, uint32_t capq, uint32_t capb, uint32_t lane_) {
    uint8_t *p = warp_base;
    pool = (DItem *)p; p += 32ull * cap * sizeof(DItem);
    holder = p; p += 32ull * a16(cap);
    free_list = (uint32_t *)p; p += 32ull * a16((uint64_t)cap * 4);
    queue = (DItem *)p; p += 32ull * capq * sizeof(DItem);
    big = (DItem *)p;
    lane = lane_;
  }
  __device__ __forceinline__ DItem &pool_at(uint32_t j) const { return pool[(uint64_t)j * 32u + lane]; }
  __device__ __forceinline__ DItem &queue_at(uint32_t j) const { return queue[(uint64_t)j * 32u + lane]; }
  __device__ __forceinline__ DItem &big_at(uint32_t j) const { return big[(uint64_t)j * 32u + lane]; }
  __device__ __forceinline__ uint32_t &free_at(uint32_t j) const { return free_list[(uint64_t)j * 32u + lane]; }
  // holder bytes in groups of 16 per lane: one 16-byte load scans 16 pool slots
  __device__ __forceinline__ uint8_t &holder_at(uint32_t j) const {
    return holder[((uint64_t)(j >> 4) * 32u + lane) * 16u + (j & 15u)];
  }
  __device__ __forceinline__ uint4 holder_vec(uint32_t v) const {
    return reinterpret_cast<const uint4 *>(holder)[(uint64_t)v * 32u + lane];
  }
};

// Shared memory of one warp. Everything is indexed [row][lane] so that the 32 lanes hit 32 banks.
struct LaneShared {
  uint32_t *tally;     // [nc][32]
  uint32_t *top_mass;  // [kLaneTop][32] ballots in the lane's heavy waiting node i; 0 = slot free
  uint32_t *val;       // [2 dp][32] float sums or uint32 counts of the node being split, heap-indexed
  __host__ __device__ static uint32_t bytes(uint32_t nc, uint32_t dp) { return (nc + kLaneTop + 2u * dp) * 32u * 4u; }
  __device__ void bind(uint8_t *p, uint32_t nc) {
    tally = (uint32_t *)p; p += nc * 32 * 4;
    top_mass = (uint32_t *)p; p += kLaneTop * 32 * 4;
    val = (uint32_t *)p;
  }
};

// Waiting nodes: those holding more than urn_max ballots are kept so that the heaviest can be taken first (it is
// the one that matters most to an undecided round) — the first kLaneTop in fixed slots with their ballot counts in
// shared memory (a light election never has more), the rest in a max-heap behind them; nodes with at most urn_max
// ballots in a circular queue.
__device__ __forceinline__ void lane_enqueue(const SimArgs &a, const LaneShared &sh, const LaneScratch &sc,
                                             const DeferredCaps &cp, LaneState &st, const DItem &ch) {
  const uint32_t lane = sc.lane;
  const uint32_t m = lane_mass(ch);
  st.unresolved += m;
  ++st.rec_w;
  if (m > a.urn_max) {
    if (st.n_top < kLaneTop && kLaneTop <= cp.big) {
      uint32_t i = 0;
#pragma unroll
      for (uint32_t k = kLaneTop; k-- > 0u;)
        if (sh.top_mass[k * 32u + lane] == 0u) i = k;
      sh.top_mass[i * 32u + lane] = m;
      sc.big_at(i) = ch;
      ++st.n_top;
    } else if (kLaneTop + st.n_big < cp.big) {
      uint32_t i = st.n_big++;
      while (i > 0u) {
        const uint32_t parent = (i - 1u) >> 1;
        const DItem pa = sc.big_at(kLaneTop + parent);
        if (lane_mass(pa) >= m) break;
        sc.big_at(kLaneTop + i) = pa;
        i = parent;
      }
      sc.big_at(kLaneTop + i) = ch;
    } else {
      st.overflow = true;
    }
    st.hw_big = max(st.hw_big, st.n_top + st.n_big);
  } else {
    if (st.q_tail - st.q_head < cp.queue) sc.queue_at((st.q_tail++) % cp.queue) = ch;
    else st.overflow = true;
    st.hw_queue = max(st.hw_queue, st.q_tail - st.q_head);
  }
}

// Removes and returns the waiting node with the most ballots (there is one: n_top + n_big > 0).
__device__ __forceinline__ DItem lane_pop_heaviest(const LaneShared &sh, const LaneScratch &sc, LaneState &st) {
  const uint32_t lane = sc.lane;
  uint32_t best = 0, best_mass = 0;
#pragma unroll
  for (uint32_t k = 0; k < kLaneTop; ++k) {
    const uint32_t m = sh.top_mass[k * 32u + lane];
    if (m > best_mass) { best_mass = m; best = k; }
  }
  if (st.n_big == 0u || lane_mass(sc.big_at(kLaneTop)) <= best_mass) {
    if (best_mass != 0u) {
      sh.top_mass[best * 32u + lane] = 0u;
      --st.n_top;
      return sc.big_at(best);
    }
  }
  const DItem top = sc.big_at(kLaneTop);
  const uint32_t n = --st.n_big;
  if (n) {
    const DItem last = sc.big_at(kLaneTop + n);
    const uint32_t lm = lane_mass(last);
    uint32_t i = 0;
    for (;;) {
      uint32_t c = 2u * i + 1u;
      if (c >= n) break;
      DItem child = sc.big_at(kLaneTop + c);
      if (c + 1u < n) {
        const DItem right = sc.big_at(kLaneTop + c + 1u);
        if (lane_mass(right) > lane_mass(child)) { child = right; ++c; }
      }
      if (lane_mass(child) <= lm) break;
      sc.big_at(kLaneTop + i) = child;
      i = c;
    }
    sc.big_at(kLaneTop + i) = last;
  }
  return top;
}

// Same contract as route_child (deferred.cuh), for a lane that owns its election.
__device__ __forceinline__ void lane_route(const SimArgs &a, const LaneShared &sh, const LaneScratch &sc,
                                           const DeferredCaps &cp, LaneState &st, const DItem &x, uint32_t depth,
                                           uint32_t cand, uint32_t tot, uint32_t cs, int32_t child, uint32_t obs) {
  const bool standing = !(st.eliminated >> cand & 1u);
  if (standing) sh.tally[cand * 32u + sc.lane] += tot;
  const bool cont = cs > 0u || (obs > 0u && child >= 0);
  if (!cont) return;
  DItem ch;
  ch.prefix.w[0] = cs + (child >= 0 ? obs : 0u);  // sampled ballots that go on + observed ones below
  ch.node_key = child_node_key(x.node_key, cand);
  ch.count = cs;
  ch.node = child;
  ch.used = x.used | (1u << cand);
  ch.pad_ = depth + 1u;
  if (!standing) {
    lane_enqueue(a, sh, sc, cp, st, ch);
    return;
  }
  // park it: fresh pool slots while there are any (no load on the way), then recycled ones
  uint32_t j;
  if (st.n_pool < cp.pool) j = st.n_pool++;
  else if (st.n_free) j = sc.free_at(--st.n_free);
  else {
    st.overflow = true;
    return;
  }
  sc.pool_at(j) = ch;
  sc.holder_at(j) = (uint8_t)cand;
  ++st.rec_w;
  st.hw_live = max(st.hw_live, st.n_pool - st.n_free);
}

// Routes the children of a node that has just been split. The number of sampled ballots of child `slot` comes from
// the lane's column of the shared tile (heavy split, row != nullptr: row[slot * 32]) or from the packed urn counts c4
// (4 bits per slot, counts <= 8). One loop over the child slots for both, so that lane_route is instantiated once.
__device__ __forceinline__ void lane_route_children(const SimArgs &a, const LaneShared &sh, const LaneScratch &sc,
                                                    const DeferredCaps &cp, LaneState &st, const DItem &x, uint32_t nb,
                                                    const uint32_t *row, const uint64_t c4[3]) {
  const uint32_t nc = a.P.nc;
  const uint32_t depth = x.pad_ & 0xFFu;
  const LevelInfo lv = level_info(a.P, depth);
  const bool tree_node = nb != 0xFFFFFFFFu;
#pragma unroll 1
  for (uint32_t slot = 0; slot < lv.K; ++slot) {
    const uint32_t c = row ? row[slot * 32u] : (uint32_t)(c4[slot >> 4] >> (4u * (slot & 15u))) & 15u;
    const uint32_t obs = (tree_node && a.use_obs) ? a.T.slot_count[nb + slot] : 0u;
    if (c + obs == 0u) continue;
    const uint32_t cand = tree_node ? (uint32_t)a.T.slot_cand[nb + slot] : nth_unused(x.used, nc, slot);
    const int32_t child = tree_node ? a.T.slot_child[nb + slot] : -1;
    lane_route(a, sh, sc, cp, st, x, depth, cand, c + obs, lv.leaf_children ? 0u : c, child, obs);
  }
}

// Urn draws of a node with a small sampled count (0..urn_max), packed 4 bits per child slot (split_small_node of
// deferred.cuh). Ballots that stop at the node (outcome K) are exhausted and not counted.
__device__ __forceinline__ void lane_urn_counts(const SimArgs &a, const DItem &x, uint32_t nb, const RngKey key, uint64_t c4[3]) {
  const LevelInfo lv = level_info(a.P, x.pad_ & 0xFFu);
  c4[0] = c4[1] = c4[2] = 0ull;
  const uint64_t picks = x.count ? urn_picks(a, lv, x.node, nb, x.count, key, x.node_key) : 0ull;
#pragma unroll 1
  for (uint32_t b = 0; b < x.count; ++b) {
    const uint32_t s = (uint32_t)(picks >> (8u * b)) & 0xFFu;
    if (s == lv.K) continue;
    c4[s >> 4] += 1ull << (4u * (s & 15u));
  }
}

// Dirichlet-multinomial split of a node with a big sampled count: split_items (split.cuh) transcribed for one
// thread. The node's heap-indexed tree is the lane's column of sh.val; on return leaf dp + slot holds the number of
// ballots of outcome `slot`.
template <bool LOG>
__device__ void lane_split_big(const SimArgs &a, const LaneShared &sh, uint32_t lane, const DItem &x, uint32_t nb,
                               const RngKey key, uint32_t *stat) {
  const LevelInfo lv = level_info(a.P, x.pad_ & 0xFFu);
  const uint32_t d = lv.d;
  const uint32_t dp = 1u << tree_log2(d);
  uint32_t *col = sh.val + lane;  // node k of the tree: col[k * 32]
  // (a) Gamma leaves
#pragma unroll 1
  for (uint32_t slot = 0; slot < dp; ++slot) {
    float g = LOG ? -INFINITY : 0.0f;
    if (slot < d) {
      const uint32_t seen = nb != 0xFFFFFFFFu ? a.T.slot_count[nb + slot] : 0u;
      g = outcome_gamma<LOG>(seen, lv.a_prior, key, x.node_key, slot, stat[kStatGammas]);
    }
    col[(dp + slot) * 32u] = __float_as_uint(g);
  }
  if (LOG) {
    float mx = -INFINITY;
#pragma unroll 1
    for (uint32_t i = 0; i < d; ++i) mx = fmaxf(mx, __uint_as_float(col[(dp + i) * 32u]));
    if (mx == -INFINITY) {
      U4 r = node_random(key, x.node_key, kStreamOneHot, 0u);
      const uint32_t pick = uniform_below(r.x, d);
#pragma unroll 1
      for (uint32_t i = 0; i < dp; ++i) col[(dp + i) * 32u] = __float_as_uint(i == pick ? 1.0f : 0.0f);
    } else {
#pragma unroll 1
      for (uint32_t i = 0; i < dp; ++i)
        col[(dp + i) * 32u] = __float_as_uint(i < d ? __expf(__fadd_rn(__uint_as_float(col[(dp + i) * 32u]), -mx)) : 0.0f);
    }
  }
  // partial sums, bottom-up
#pragma unroll 1
  for (uint32_t k = dp - 1u; k >= 1u; --k)
    col[k * 32u] = __float_as_uint(__fadd_rn(__uint_as_float(col[2 * k * 32u]), __uint_as_float(col[(2 * k + 1) * 32u])));
  // (b) counts, top-down (heap order visits a node after its parent)
  col[32u] = x.count;
#pragma unroll 1
  for (uint32_t k = 1u; k < dp; ++k) {
    const uint32_t n = col[k * 32u];
    const float sl = __uint_as_float(col[2 * k * 32u]), sr = __uint_as_float(col[(2 * k + 1) * 32u]);
    uint32_t left = 0;
    if (n != 0u) {
      if (!(sr > 0.0f)) {
        left = n;
      } else if (sl > 0.0f) {
        const float tot = __fadd_rn(sl, sr);
        const bool flip = sl > sr;
        uint32_t att = 0;
        const uint32_t kk = binomial_small_p(n, __fdividef(flip ? sr : sl, tot), key, x.node_key, k, att, a.P.binv_max);
        left = flip ? n - kk : kk;
        stat[kStatBinomials] += att;
      }
    }
    col[2 * k * 32u] = left;
    col[(2 * k + 1) * 32u] = n - left;
  }
}

#ifndef DTB_LANE_THREADS_PER_SM
#define DTB_LANE_THREADS_PER_SM 768
#endif
template <int BLOCK, bool LOG>
__global__ void __launch_bounds__(BLOCK, DTB_LANE_THREADS_PER_SM / BLOCK > 0 ? DTB_LANE_THREADS_PER_SM / BLOCK : 1) lane_kernel(const __grid_constant__ SimArgs a) {
  extern __shared__ __align__(16) uint8_t lane_smem[];
  __shared__ uint32_t s_wins[kMaxCandidates];
  const int tid = threadIdx.x;
  const uint32_t lane = tid & 31u, warp = tid >> 5;
  const uint32_t nc = a.P.nc;
  const uint32_t cand_mask = nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u);
  const uint32_t dp_max = 1u << tree_log2(level_info(a.P, 0).d);  // the root is the widest node
  LaneShared sh;
  sh.bind(lane_smem + (size_t)warp * LaneShared::bytes(nc, dp_max), nc);
  const DeferredCaps cp{a.cap_items, a.cap_queue, a.cap_big};
  LaneScratch sc;
  sc.bind(a.scratch + ((uint64_t)blockIdx.x * (blockDim.x >> 5) + warp) * 32ull * a.scratch_stride, cp.pool, cp.queue, cp.big, lane);
  uint32_t stat[kStatCount], usage[3] = {0u, 0u, 0u};
#pragma unroll
  for (int i = 0; i < kStatCount; ++i) stat[i] = 0;
  if (tid < (int)kMaxCandidates) s_wins[tid] = 0;
  __syncthreads();
  const uint32_t n_rounds = a.elim_orders ? nc - 1 : nc - a.n_winners;
  const bool early_stop = a.n_winners == 1u && a.elim_orders == nullptr;

  for (;;) {
    // the elections of a warp are claimed, and finished, together: 32 of them, or fewer (a.lane_width) when the job
    // has too few elections to fill the machine that way - a warp's time grows with the number of lanes that diverge
    uint32_t base = 0;
    if (lane == 0u) base = atomicAdd(a.work_counter, a.lane_width);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (a.lane_sync) {
      // the warps of the block go through their batches together (a warp without one idles through the barriers)
      if (__syncthreads_and(base >= a.n_elections)) break;
    } else if (base >= a.n_elections) {
      break;
    }
    const uint32_t e = base + lane;
    const bool active = lane < a.lane_width && e < a.n_elections && base < a.n_elections;
    const RngKey key = election_key(a.seed, a.first_election + e);
    LaneState st;
    st.n_pool = 0; st.n_free = 0; st.q_head = 0; st.q_tail = 0; st.n_big = 0; st.n_top = 0; st.eliminated = 0;
    st.rec_w = 0; st.rec_r = 0; st.hw_live = 0; st.n_splits = 0;
    st.hw_queue = 0; st.hw_big = 0; st.unresolved = 0; st.overflow = false;
#pragma unroll
    for (uint32_t k = 0; k < kLaneTop; ++k) sh.top_mass[k * 32u + lane] = 0;
#pragma unroll 1
    for (uint32_t c = 0; c < nc; ++c) sh.tally[c * 32u + lane] = 0;
    bool decided = !active;

    // Splits the node x of every lane with `want` set and routes its children.
    auto split_node = [&](bool want, const DItem &x) {
      if (!want) return;
      const uint32_t nb = x.node >= 0 ? a.T.node_base[x.node] : 0xFFFFFFFFu;
      const bool heavy = x.count > a.urn_max;
      const uint32_t d = level_info(a.P, x.pad_ & 0xFFu).d;
      uint64_t c4[3];
      stat[kStatItems] += 1;
      if (nb != 0xFFFFFFFFu && (heavy || a.use_obs)) stat[kStatSlots] += d;
      if (heavy) {
        lane_split_big<LOG>(a, sh, lane, x, nb, key, stat);
      } else {
        stat[kStatUrnItems] += 1;
        lane_urn_counts(a, x, nb, key, c4);
      }
      lane_route_children(a, sh, sc, cp, st, x, nb, heavy ? sh.val + ((1u << tree_log2(d)) * 32u + lane) : nullptr, c4);
    };

    // Pass -1 splits the root; pass r >= 0 eliminates a candidate and releases the nodes parked under it.
    // Released nodes are NOT split as they appear: they wait (lane_enqueue), their ballots counted in
    // st.unresolved, and a round is decided as soon as its outcome cannot depend on where those ballots go. They
    // can only add to tallies, so the weakest candidate stays weakest if its tally plus all unresolved ballots is
    // still below every other tally (and a leader keeps a majority if it has one counting them all against it).
    // Only while that is not the case are waiting nodes split, the one with the most ballots first. Every split is
    // a pure function of the node, so which of them are carried out changes no outcome: elimination orders and
    // winners are those of the eager evaluation (deferred_kernel, simulate_kernel).
    {
      DItem root;
      root.prefix.w[0] = 0;
      root.node_key = kRootNodeKey;
      root.count = a.n_sample;
      root.node = 0;
      root.used = 0;
      root.pad_ = 0;
      split_node(active, root);  // the root holds every ballot: split at once
    }
    if (a.lane_sync) __syncthreads();
    if (a.sc_function == 1) {
      // Plurality (R/social_choice.R:58-83): the candidates with the most first preferences, all of them when tied.
      // First preferences are the root's split plus the observed first preferences that rode along: nothing below
      // the root is ever drawn.
      if (active) {
        uint32_t best = 0;
#pragma unroll 1
        for (uint32_t c = 0; c < nc; ++c) best = max(best, sh.tally[c * 32u + lane]);
#pragma unroll 1
        for (uint32_t c = 0; c < nc; ++c)
          if (best != 0u && sh.tally[c * 32u + lane] == best) atomicAdd(&s_wins[c], 1u);
      }
      __syncwarp();
      continue;
    }
    for (uint32_t round = 0; round < n_rounds; ++round) {
      if (a.lane_sync > 1u) __syncthreads();
      bool pending = !decided && !st.overflow;  // this lane still has to settle the round
      for (;;) {
        bool uncertain = false;
        if (pending) {
          uint32_t mn = 0xFFFFFFFFu, second = 0xFFFFFFFFu, tied = 0, ntied = 0, total = 0, top = 0, top_c = 0;
#pragma unroll 1
          for (uint32_t c = 0; c < nc; ++c) {
            if (st.eliminated >> c & 1u) continue;
            const uint32_t t = sh.tally[c * 32u + lane];
            if (t < mn) { second = mn; mn = t; tied = 1u << c; ntied = 1; }
            else if (t == mn) { tied |= 1u << c; ++ntied; }
            else if (t < second) second = t;
            total += t;
            if (t > top) { top = t; top_c = c; }
          }
          const uint32_t u = st.unresolved;
          if (early_stop && top - u > total - top && top > u) {
            st.eliminated = ~(1u << top_c) & cand_mask;
            decided = true;
            pending = false;
          } else if (u != 0u && (ntied > 1u || mn + u >= second)) {
            uncertain = true;
          } else {
            pending = false;
            if (ntied > 1u) {
              U4 r = node_random(key, 0ull, kStreamTie, round);
              const uint32_t pick = uniform_below(r.x, ntied);
              for (uint32_t k = 0; k < pick; ++k) tied &= tied - 1u;
            }
            const uint32_t chosen = (uint32_t)__ffs(tied) - 1u;
            if (a.elim_orders) a.elim_orders[(uint64_t)e * nc + round] = (uint8_t)chosen;
            st.eliminated |= 1u << chosen;
            if (mn != 0u && round + 1u != n_rounds) {
              // queue the nodes parked under the eliminated candidate
              const uint32_t np = min(st.n_pool, cp.pool);
              const uint32_t n_vec = (np + 15u) / 16u;
              const uint32_t pattern = chosen * 0x01010101u;
#pragma unroll 1
              for (uint32_t v = 0; v < n_vec; ++v) {
                const uint4 h = sc.holder_vec(v);
                const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
                uint32_t hits = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint32_t eq = __vcmpeq4(hw[q], pattern);
                  hits |= ((eq & 1u) | ((eq >> 7) & 2u) | ((eq >> 14) & 4u) | ((eq >> 21) & 8u)) << (4 * q);
                }
                while (hits) {
                  const uint32_t j = v * 16u + (uint32_t)__ffs(hits) - 1u;
                  hits &= hits - 1u;
                  if (j >= np) break;
                  sc.holder_at(j) = 0xFF;
                  const DItem x = sc.pool_at(j);
                  ++st.rec_r;
                  lane_enqueue(a, sh, sc, cp, st, x);
                  sc.free_at(st.n_free++) = j;
                }
              }
            }
          }
        }
        if (a.lane_sync > 2u ? !__syncthreads_or(uncertain) : !__any_sync(0xffffffffu, uncertain)) break;
        if (uncertain && !st.overflow && a.lane_max_splits != 0u && ++st.n_splits > a.lane_max_splits)
          st.overflow = true;  // a heavy election: a whole warp finishes it faster (replayed by deferred_kernel)
        const bool want = uncertain && !st.overflow;
        DItem x;
        x.count = 0;
        if (want) {
          x = (st.n_top + st.n_big) ? lane_pop_heaviest(sh, sc, st) : sc.queue_at((st.q_head++) % cp.queue);
          st.unresolved -= lane_mass(x);
          ++st.rec_r;
        }
        split_node(want, x);
        if (st.overflow) pending = false;
      }
    }
    if (active) {
      stat[kStatEntries] += st.rec_w;
      stat[kStatRereads] += st.rec_r;
      usage[0] = max(usage[0], st.hw_live);
      usage[1] = max(usage[1], st.hw_queue);
      usage[2] = max(usage[2], st.hw_big);
      if (st.overflow) {
        const uint32_t k = atomicAdd(a.retry_count, 1u);
        if (k < a.retry_cap) a.retry_list[k] = e;
        if (a.usage) atomicAdd(&a.usage[4], 1u);
      } else {
        const uint32_t standing_mask = ~st.eliminated & cand_mask;
        if (a.elim_orders) {
          uint8_t *o = a.elim_orders + (uint64_t)e * nc;
          o[nc - 1] = (uint8_t)(__ffs(standing_mask) - 1);
          for (uint32_t r = nc - a.n_winners; r < nc; ++r) atomicAdd(&s_wins[o[r]], 1u);
        } else {
          uint32_t m = standing_mask;
          while (m) {
            atomicAdd(&s_wins[__ffs(m) - 1], 1u);
            m &= m - 1u;
          }
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (tid < (int)nc && s_wins[tid]) atomicAdd(&a.win_counts[tid], (unsigned long long)s_wins[tid]);
  if (a.usage) {
    // the thread-per-election arenas report their high-water marks apart from deferred_kernel's ([5..7])
    const uint32_t u0 = __reduce_max_sync(0xffffffffu, usage[0]), u1 = __reduce_max_sync(0xffffffffu, usage[1]),
                   u2 = __reduce_max_sync(0xffffffffu, usage[2]);
    if (lane == 0u) {
      atomicMax(&a.usage[5], u0);
      atomicMax(&a.usage[6], u1);
      atomicMax(&a.usage[7], u2);
    }
  }
  if (a.stats) {
#pragma unroll
    for (int i = 0; i < kStatCount; ++i) {
      uint32_t s = __reduce_add_sync(0xffffffffu, stat[i]);
      if (lane == 0u && s) atomicAdd(&a.stats[i], (unsigned long long)s);
    }
  }
}

}  // namespace dtb
