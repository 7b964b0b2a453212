// lane.cuh — the deferred algorithm of deferred.cuh with ONE THREAD PER ELECTION.
//
// deferred_kernel gives an election a whole warp; with ~10^3 node splits per election most of those lanes
// idle (7 of 32 active in the profile) and the warp-cooperative machinery (ballots, shared-memory tiles,
// reservations) is overhead. Here a lane owns an election outright: plain scalar code, private queues in
// its own arena, no atomics. The 32 elections of a warp start together and walk nearly the same nodes in
// nearly the same order (the root, then the nodes parked under the weakest candidates...), so the lanes stay
// largely in step where it matters — the big Gamma / binomial splits near the top of the tree.
//
// Every split is the same pure function of (seed, election, node key, parameters, count) as in split.cuh:
// the Gamma leaves, the sum tree and the binomial tree below are the scalar transcription of split_items,
// operation for operation, so outcomes are bit-identical to the other two kernels
// (tests/test_gpu_parity.py::test_deferred_equals_materialised runs all three).
//
// An election that outgrows its arena is not recorded; its index goes to a retry list that the host replays
// with deferred_kernel (which has worst-case arenas).
#pragma once
#include <cstdint>

#include "deferred.cuh"

namespace dtb {

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

constexpr uint32_t kLaneTop = 32;  // heavy waiting nodes per election whose ordering never touches global memory

template <int BLOCK>
struct LaneShared {
  uint32_t tally[kMaxCandidates][BLOCK];  // [candidate][thread]: a warp reading one candidate hits 32 banks
  uint32_t top_mass[kLaneTop][BLOCK];     // ballots in big[0][i], i < kLaneTop; 0 = slot free
};

// Waiting nodes: those holding more than urn_max ballots are kept so that the heaviest can be taken first (it is
// the one that matters most to an undecided round) — the first kLaneTop in fixed slots of big[0] with their
// ballot counts in shared memory (a light election never has more), the rest in a max-heap behind them; nodes
// with at most urn_max ballots in a circular queue.
template <int BLOCK>
__device__ __forceinline__ void lane_enqueue(const SimArgs &a, LaneShared<BLOCK> &sh, const DeferredScratch &sc,
                                             const DeferredCaps &cp, LaneState &st, const DItem &ch) {
  const uint32_t m = lane_mass(ch);
  st.unresolved += m;
  ++st.rec_w;
  if (m > a.urn_max) {
    if (st.n_top < kLaneTop && kLaneTop <= cp.big) {
      uint32_t i = 0;
#pragma unroll
      for (uint32_t k = kLaneTop; k-- > 0u;)
        if (sh.top_mass[k][threadIdx.x] == 0u) i = k;
      sh.top_mass[i][threadIdx.x] = m;
      sc.big[0][i] = ch;
      ++st.n_top;
    } else if (kLaneTop + st.n_big < cp.big) {
      DItem *heap = sc.big[0] + kLaneTop;
      uint32_t i = st.n_big++;
      while (i > 0u) {
        const uint32_t parent = (i - 1u) >> 1;
        if (lane_mass(heap[parent]) >= m) break;
        heap[i] = heap[parent];
        i = parent;
      }
      heap[i] = ch;
    } else {
      st.overflow = true;
    }
    st.hw_big = max(st.hw_big, st.n_top + st.n_big);
  } else {
    if (st.q_tail - st.q_head < cp.queue) sc.queue[(st.q_tail++) % cp.queue] = ch;
    else st.overflow = true;
    st.hw_queue = max(st.hw_queue, st.q_tail - st.q_head);
  }
}

// Removes and returns the waiting node with the most ballots (there is one: n_top + n_big > 0).
template <int BLOCK>
__device__ __forceinline__ DItem lane_pop_heaviest(LaneShared<BLOCK> &sh, const DeferredScratch &sc, LaneState &st) {
  uint32_t best = 0, best_mass = 0;
#pragma unroll
  for (uint32_t k = 0; k < kLaneTop; ++k) {
    const uint32_t m = sh.top_mass[k][threadIdx.x];
    if (m > best_mass) { best_mass = m; best = k; }
  }
  DItem *heap = sc.big[0] + kLaneTop;
  if (st.n_big == 0u || lane_mass(heap[0]) <= best_mass) {
    if (best_mass != 0u) {
      sh.top_mass[best][threadIdx.x] = 0u;
      --st.n_top;
      return sc.big[0][best];
    }
  }
  const DItem top = heap[0];
  const uint32_t n = --st.n_big;
  if (n) {
    const DItem last = heap[n];
    const uint32_t lm = lane_mass(last);
    uint32_t i = 0;
    for (;;) {
      uint32_t c = 2u * i + 1u;
      if (c >= n) break;
      uint32_t cm = lane_mass(heap[c]);
      if (c + 1u < n) {
        const uint32_t rm = lane_mass(heap[c + 1u]);
        if (rm > cm) { cm = rm; ++c; }
      }
      if (cm <= lm) break;
      heap[i] = heap[c];
      i = c;
    }
    heap[i] = last;
  }
  return top;
}

// Same contract as route_child (deferred.cuh), for a lane that owns its election.
template <int BLOCK>
__device__ __forceinline__ void lane_route(const SimArgs &a, LaneShared<BLOCK> &sh, const DeferredScratch &sc,
                                           const DeferredCaps &cp, LaneState &st, const DItem &x, uint32_t depth,
                                           uint32_t cand, uint32_t tot, uint32_t cs, int32_t child, uint32_t obs) {
  const bool standing = !(st.eliminated >> cand & 1u);
  if (standing) sh.tally[cand][threadIdx.x] += tot;
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
    lane_enqueue<BLOCK>(a, sh, sc, cp, st, ch);
    return;
  }
  // park it: fresh pool slots while there are any (no load on the way), then recycled ones
  uint32_t j;
  if (st.n_pool < cp.pool) j = st.n_pool++;
  else if (st.n_free) j = sc.free_list[--st.n_free];
  else {
    st.overflow = true;
    return;
  }
  sc.pool[j] = ch;
  sc.holder[j] = (uint8_t)cand;
  ++st.rec_w;
  st.hw_live = max(st.hw_live, st.n_pool - st.n_free);
}

// Split of a node with a small sampled count (0..urn_max): split_small_node of deferred.cuh without the atomics.
template <int BLOCK>
__device__ __forceinline__ void lane_split_small(const SimArgs &a, LaneShared<BLOCK> &sh, const DeferredScratch &sc,
                                                 const DeferredCaps &cp, LaneState &st, const DItem &x, const RngKey key,
                                                 uint32_t *stat) {
  const uint32_t nc = a.P.nc;
  const uint32_t depth = x.pad_ & 0xFFu;
  const LevelInfo lv = level_info(a.P, depth);
  const uint32_t nb = x.node >= 0 ? a.T.node_base[x.node] : 0xFFFFFFFFu;
  stat[kStatItems] += 1;
  stat[kStatUrnItems] += 1;
  const uint64_t picks = x.count ? urn_picks(a, lv, x.node, nb, x.count, key, x.node_key) : 0ull;
  if (nb != 0xFFFFFFFFu && a.use_obs) {
    uint64_t c4[2] = {0ull, 0ull};
#pragma unroll 1
    for (uint32_t b = 0; b < x.count; ++b) {
      const uint32_t s = (uint32_t)(picks >> (8u * b)) & 0xFFu;
      if (s == lv.K) continue;
      c4[s >> 4] += 1ull << (4u * (s & 15u));
    }
    stat[kStatSlots] += lv.d;
#pragma unroll 1
    for (uint32_t slot = 0; slot < lv.K; ++slot) {
      const uint32_t c = (uint32_t)(c4[slot >> 4] >> (4u * (slot & 15u))) & 15u;
      const uint32_t obs = a.T.slot_count[nb + slot];
      if (c + obs == 0u) continue;
      lane_route<BLOCK>(a, sh, sc, cp, st, x, depth, a.T.slot_cand[nb + slot], c + obs, lv.leaf_children ? 0u : c,
                        a.T.slot_child[nb + slot], obs);
    }
  } else {
    uint32_t done = 0;
#pragma unroll 1
    for (uint32_t b = 0; b < x.count; ++b) {
      if (done >> b & 1u) continue;
      const uint32_t slot = (uint32_t)(picks >> (8u * b)) & 0xFFu;
      uint32_t c = 1;
#pragma unroll 1
      for (uint32_t b2 = b + 1; b2 < x.count; ++b2)
        if (((uint32_t)(picks >> (8u * b2)) & 0xFFu) == slot) {
          ++c;
          done |= 1u << b2;
        }
      if (slot == lv.K) continue;
      const uint32_t cand = nb != 0xFFFFFFFFu ? (uint32_t)a.T.slot_cand[nb + slot] : nth_unused(x.used, nc, slot);
      const int32_t child = nb != 0xFFFFFFFFu ? a.T.slot_child[nb + slot] : -1;
      lane_route<BLOCK>(a, sh, sc, cp, st, x, depth, cand, c, lv.leaf_children ? 0u : c, child, 0u);
    }
  }
}

// Split of a node with a big sampled count: split_items (split.cuh) transcribed for one thread. `row` is the
// heap-indexed tree of the node (2 * dp + 1 words of thread-local memory).
template <int BLOCK, bool LOG>
__device__ void lane_split_big(const SimArgs &a, LaneShared<BLOCK> &sh, const DeferredScratch &sc,
                               const DeferredCaps &cp, LaneState &st, const DItem &x, const RngKey key, uint32_t *row,
                               uint32_t *stat) {
  const uint32_t nc = a.P.nc;
  const uint32_t depth = x.pad_ & 0xFFu;
  const LevelInfo lv = level_info(a.P, depth);
  const uint32_t d = lv.d;
  const int log_dp = tree_log2(d);
  const uint32_t dp = 1u << log_dp;
  const uint32_t nb = x.node >= 0 ? a.T.node_base[x.node] : 0xFFFFFFFFu;
  stat[kStatItems] += 1;
  if (nb != 0xFFFFFFFFu) stat[kStatSlots] += d;
  // (a) Gamma leaves
#pragma unroll 1
  for (uint32_t slot = 0; slot < dp; ++slot) {
    float g = LOG ? -INFINITY : 0.0f;
    if (slot < d) {
      float alpha = lv.a_prior;
      if (nb != 0xFFFFFFFFu) alpha = __fadd_rn((float)a.T.slot_count[nb + slot], lv.a_prior);
      if (alpha > 0.0f) {
        uint32_t att = 0;
        g = gamma_variate<LOG>(alpha, key, x.node_key, slot, att);
        stat[kStatGammas] += 1;
      }
    }
    row[dp + slot] = __float_as_uint(g);
  }
  if (LOG) {
    float mx = -INFINITY;
#pragma unroll 1
    for (uint32_t i = 0; i < d; ++i) mx = fmaxf(mx, __uint_as_float(row[dp + i]));
    if (mx == -INFINITY) {
      U4 r = node_random(key, x.node_key, kStreamOneHot, 0u);
      const uint32_t pick = uniform_below(r.x, d);
#pragma unroll 1
      for (uint32_t i = 0; i < dp; ++i) row[dp + i] = __float_as_uint(i == pick ? 1.0f : 0.0f);
    } else {
#pragma unroll 1
      for (uint32_t i = 0; i < dp; ++i)
        row[dp + i] = __float_as_uint(i < d ? __expf(__fadd_rn(__uint_as_float(row[dp + i]), -mx)) : 0.0f);
    }
  }
  // partial sums, bottom-up
#pragma unroll 1
  for (uint32_t k = dp - 1u; k >= 1u; --k)
    row[k] = __float_as_uint(__fadd_rn(__uint_as_float(row[2 * k]), __uint_as_float(row[2 * k + 1])));
  // (b) counts, top-down (heap order visits a node after its parent)
  row[1] = x.count;
#pragma unroll 1
  for (uint32_t k = 1u; k < dp; ++k) {
    const uint32_t n = row[k];
    const float sl = __uint_as_float(row[2 * k]), sr = __uint_as_float(row[2 * k + 1]);
    uint32_t left = 0;
    if (n != 0u) {
      if (!(sr > 0.0f)) {
        left = n;
      } else if (sl > 0.0f) {
        const float tot = __fadd_rn(sl, sr);
        const bool flip = sl > sr;
        uint32_t att = 0;
        const uint32_t kk = binomial_small_p(n, __fdividef(flip ? sr : sl, tot), key, x.node_key, k, att);
        left = flip ? n - kk : kk;
        stat[kStatBinomials] += att;
      }
    }
    row[2 * k] = left;
    row[2 * k + 1] = n - left;
  }
  // routing (split_big_nodes of deferred.cuh)
#pragma unroll 1
  for (uint32_t slot = 0; slot < lv.K; ++slot) {
    const uint32_t c = row[dp + slot];
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
    lane_route<BLOCK>(a, sh, sc, cp, st, x, depth, cand, c + obs, lv.leaf_children ? 0u : c, child, obs);
  }
}

template <int BLOCK, bool LOG>
__global__ void __launch_bounds__(BLOCK, 768 / BLOCK) lane_kernel(const __grid_constant__ SimArgs a) {
  __shared__ LaneShared<BLOCK> sh;
  __shared__ uint32_t s_wins[kMaxCandidates];
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t nc = a.P.nc;
  const uint32_t cand_mask = nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u);
  const DeferredCaps cp{a.cap_items, a.cap_queue, a.cap_big};
  DeferredScratch sc;
  sc.bind(a.scratch + ((uint64_t)blockIdx.x * BLOCK + tid) * a.scratch_stride, cp.pool, cp.queue, cp.big);
  uint32_t row[2 * 64 + 1];
  uint32_t stat[kStatCount], usage[3] = {0u, 0u, 0u};
#pragma unroll
  for (int i = 0; i < kStatCount; ++i) stat[i] = 0;
  if (tid < (int)kMaxCandidates) s_wins[tid] = 0;
  __syncthreads();
  const uint32_t n_rounds = a.elim_orders ? nc - 1 : nc - a.n_winners;
  const bool early_stop = a.n_winners == 1u && a.elim_orders == nullptr;

  for (;;) {
    // the 32 elections of a warp are claimed, and finished, together
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(a.work_counter, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= a.n_elections) break;
    const uint32_t e = base + lane;
    const bool active = e < a.n_elections;
    const RngKey key = election_key(a.seed, a.first_election + e);
    LaneState st;
    st.n_pool = 0; st.n_free = 0; st.q_head = 0; st.q_tail = 0; st.n_big = 0; st.n_top = 0; st.eliminated = 0; st.rec_w = 0; st.rec_r = 0; st.hw_live = 0; st.n_splits = 0;
#pragma unroll
    for (uint32_t k = 0; k < kLaneTop; ++k) sh.top_mass[k][tid] = 0;
    st.hw_queue = 0; st.hw_big = 0; st.unresolved = 0; st.overflow = false;
#pragma unroll 1
    for (uint32_t c = 0; c < nc; ++c) sh.tally[c][tid] = 0;
    DItem root;
    root.prefix.w[0] = 0;
    root.node_key = kRootNodeKey;
    root.count = a.n_sample;
    root.node = 0;
    root.used = 0;
    root.pad_ = 0;
    bool decided = !active;
    // Pass -1 splits the root; pass r >= 0 eliminates a candidate and releases the nodes parked under it.
    // Released nodes are NOT split as they appear: they wait (lane_enqueue), their ballots counted in
    // st.unresolved, and a round is decided as soon as its outcome cannot depend on where those ballots go. They
    // can only add to tallies, so the weakest candidate stays weakest if its tally plus all unresolved ballots is
    // still below every other tally (and a leader keeps a majority if it has one counting them all against it).
    // Only while that is not the case are waiting nodes split, the one with the most ballots first. Every split is
    // a pure function of the node, so which of them are carried out changes no outcome: elimination orders and
    // winners are those of the eager evaluation (deferred_kernel, simulate_kernel).
    for (int pass = -1; pass < (int)n_rounds; ++pass) {
      if (pass < 0) {
        // the root holds every ballot: split at once
        if (active) {
          if (a.n_sample > a.urn_max) lane_split_big<BLOCK, LOG>(a, sh, sc, cp, st, root, key, row, stat);
          else lane_split_small<BLOCK>(a, sh, sc, cp, st, root, key, stat);
        }
      } else {
        const uint32_t round = (uint32_t)pass;
        bool pending = !decided && !st.overflow;  // this lane still has to settle the round
        for (;;) {
          bool uncertain = false;
          if (pending) {
            uint32_t mn = 0xFFFFFFFFu, second = 0xFFFFFFFFu, tied = 0, ntied = 0, total = 0, top = 0, top_c = 0;
#pragma unroll 1
            for (uint32_t c = 0; c < nc; ++c) {
              if (st.eliminated >> c & 1u) continue;
              const uint32_t t = sh.tally[c][tid];
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
                  const uint4 h = reinterpret_cast<const uint4 *>(sc.holder)[v];
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
                    sc.holder[j] = 0xFF;
                    const DItem x = sc.pool[j];
                    ++st.rec_r;
                    lane_enqueue<BLOCK>(a, sh, sc, cp, st, x);
                    sc.free_list[st.n_free++] = j;
                  }
                }
              }
            }
          }
          if (!__any_sync(0xffffffffu, uncertain)) break;
          if (uncertain && !st.overflow && a.lane_max_splits != 0u && ++st.n_splits > a.lane_max_splits)
            st.overflow = true;  // a heavy election: a whole warp finishes it faster (replayed by deferred_kernel)
          if (uncertain && !st.overflow) {
            const DItem x = (st.n_top + st.n_big) ? lane_pop_heaviest<BLOCK>(sh, sc, st) : sc.queue[(st.q_head++) % cp.queue];
            st.unresolved -= lane_mass(x);
            ++st.rec_r;
            if (x.count > a.urn_max) lane_split_big<BLOCK, LOG>(a, sh, sc, cp, st, x, key, row, stat);
            else lane_split_small<BLOCK>(a, sh, sc, cp, st, x, key, stat);
          }
          if (st.overflow) pending = false;
        }
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
    const uint32_t u0 = __reduce_max_sync(0xffffffffu, usage[0]), u1 = __reduce_max_sync(0xffffffffu, usage[1]),
                   u2 = __reduce_max_sync(0xffffffffu, usage[2]);
    if (lane == 0) {
      atomicMax(&a.usage[0], u0);
      atomicMax(&a.usage[1], u1);
      atomicMax(&a.usage[2], u2);
    }
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
