// simulate.cuh — the posterior-simulation kernel: one thread block simulates one election at a
// time (persistent blocks pulling election ids from a counter):
//
//   (a)+(b) level-synchronous expansion of the Dirichlet-tree (IRVNode::sample, irv_node.cpp:107-174,
//           and lazyIRVBallots, :27-84): the frontier of one depth is processed in chunks of BLOCK
//           items; Gamma variates are drawn one thread per (item, outcome) into shared memory, then
//           one thread per item normalises them and splits the item's count with sequential
//           binomials (rMultinomial, distributions.cpp:21-53); children and finished ballots are
//           compacted with a block-wide prefix sum into the next frontier / the entry list;
//   (c)     IRV elimination over the election's (ballot, count) entries plus the shared observed
//           entries (socialChoiceIRV, irv_ballot.cpp:42-138);
//   (d)     win counts accumulated per block and flushed with one atomic per candidate.
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
  // per-block scratch
  uint8_t *scratch;
  uint64_t scratch_stride;    // bytes per block
  uint32_t cap_items;         // frontier capacity (items) per buffer
  uint32_t cap_entries;
  // outputs
  unsigned long long *win_counts;  // [nc], added to
  uint8_t *elim_orders;            // NULL or [n_elections * nc]
  uint32_t *entry_count_out;       // dump mode: number of entries of the (single) election
  uint32_t *tie_rounds_out;        // NULL or [n_elections]
  unsigned int *work_counter;      // next election index to hand out
  int *error_flag;                 // set to 1 on scratch overflow
  unsigned long long *stats;       // [kStatCount]
};

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

// ---- block-wide helpers ------------------------------------------------------------------------

// Exclusive prefix sum of one uint32 per thread; returns this thread's offset, *total = block sum.
template <int BLOCK>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *s_warp /*[33]*/, uint32_t *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = (BLOCK + 31) / 32;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (NW == 1) {
    *total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
  }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < NW ? s_warp[lane] : 0u;
    uint32_t xs = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, xs, o);
      if (lane >= o) xs += y;
    }
    if (lane < NW) s_warp[lane] = xs - w;
    if (lane == NW - 1) s_warp[32] = xs;
  }
  __syncthreads();
  uint32_t off = s_warp[warp] + x - v;
  *total = s_warp[32];
  __syncthreads();  // s_warp may be reused immediately by the caller
  return off;
}

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
  uint32_t avail = ~used & (nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u));
  for (uint32_t k = 0; k < i; ++k) avail &= avail - 1u;
  return (uint32_t)__ffs(avail) - 1u;
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

// ---- shared-memory layout ------------------------------------------------------------------------
// Dynamic shared memory: BLOCK * dstride floats (the Gamma / count matrix, column = item).
template <int BLOCK>
struct SharedFixed {
  uint64_t node_key[BLOCK];
  uint32_t nbase[BLOCK];       // first CSR slot of the item's node, UINT32_MAX for a prior-only item
  uint32_t count[BLOCK];
  uint16_t gamma_list[BLOCK];  // chunk-local indices of the items that need Gamma variates
  uint32_t n_gamma;
  uint32_t warp_scan[33];
  uint32_t n_in, n_out, n_entries;
  uint32_t tally[kMaxCandidates];
  uint32_t wins[kMaxCandidates];
  uint32_t eliminated, chosen, tie_rounds;
  uint32_t election;
  uint8_t order[kMaxCandidates];
};

// One election's expansion. Leaves the entries in sc.ent_* and their number in sm.n_entries.
template <int W, int BLOCK, bool LOG>
__device__ void expand_election(const SimArgs &a, Scratch<W> &sc, SharedFixed<BLOCK> &sm, float *s_val,
                                const RngKey key, uint32_t *stat) {
  const uint32_t nc = a.P.nc;
  const int tid = threadIdx.x;
  const uint32_t dstride = (nc + 1u) | 1u;  // odd: column accesses are bank-conflict free

  if (tid == 0) {
    sm.n_entries = 0;
    sm.n_in = a.n_sample > 0 ? 1u : 0u;
    if (a.n_sample > 0) {
      Item<W> root;
#pragma unroll
      for (int w = 0; w < W; ++w) root.prefix.w[w] = 0;
      root.node_key = kRootNodeKey;
      root.count = a.n_sample;
      root.node = 0;
      root.used = 0;
      root.pad_ = 0;
      sc.frontier[0][0] = root;
    }
  }
  __syncthreads();

  int cur = 0;
  for (uint32_t depth = 0; depth + 1 < nc; ++depth) {
    const uint32_t n_in = sm.n_in;
    if (n_in == 0) break;
    const uint32_t K = nc - depth;
    const bool T = depth >= a.P.min_depth;
    const uint32_t d = K + (T ? 1u : 0u);
    const float a_prior = a.P.a_prior[depth];
    // irv_node.cpp:136 (depth == maxDepth-1) and :46 (child depth == nc-1 or == maxDepth)
    const bool leaf_children = (depth == a.P.leaf_depth) || (depth + 2 == nc);
    const Item<W> *in = sc.frontier[cur];
    Item<W> *out = sc.frontier[cur ^ 1];
    if (tid == 0) sm.n_out = 0;
    __syncthreads();

    for (uint32_t base = 0; base < n_in; base += BLOCK) {
      const uint32_t gi = base + tid;
      const bool active = gi < n_in;
      Item<W> it;
      bool use_gamma = false;
      if (tid == 0) sm.n_gamma = 0;
      __syncthreads();
      if (active) {
        it = in[gi];
        sm.node_key[tid] = it.node_key;
        sm.nbase[tid] = it.node >= 0 ? a.T.node_base[it.node] : 0xFFFFFFFFu;
        sm.count[tid] = it.count;
        use_gamma = it.count > a.urn_max;
      }
      {  // compact the Gamma-mode items of this chunk
        uint32_t m = __ballot_sync(0xffffffffu, use_gamma);
        uint32_t wbase = 0;
        if ((tid & 31) == 0 && m) wbase = atomicAdd(&sm.n_gamma, (uint32_t)__popc(m));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (use_gamma) sm.gamma_list[wbase + __popc(m & ((1u << (tid & 31)) - 1u))] = (uint16_t)tid;
      }
      __syncthreads();

      // ---- (a) Gamma variates: one thread per (item, outcome) --------------------------------
      {
        const uint32_t total = sm.n_gamma * d;
        for (uint32_t s = tid; s < total; s += BLOCK) {
          const uint32_t li = sm.gamma_list[s / d];
          const uint32_t slot = s % d;  // slot K is the stop outcome (present only when T)
          const uint32_t nb = sm.nbase[li];
          float alpha = a_prior;
          if (nb != 0xFFFFFFFFu) alpha = __fadd_rn((float)a.T.slot_count[nb + slot], a_prior);
          float g;
          if (alpha > 0.0f) {
            uint32_t att = 0;
            g = gamma_variate<LOG>(alpha, key, sm.node_key[li], slot, att);
            stat[kStatGammas] += 1;
          } else {
            g = LOG ? -INFINITY : 0.0f;  // Gamma(0) is the point mass at 0
          }
          s_val[li * dstride + slot] = g;
        }
      }
      __syncthreads();

      // ---- (b) split the item's count over its outcomes: one thread per item -------------------
      uint32_t n_child = 0, n_ent = 0;
      float *col = s_val + tid * dstride;
      if (active) {
        stat[kStatItems] += 1;
        if (it.node >= 0) stat[kStatSlots] += d;
        const uint32_t nbase = sm.nbase[tid];
        if (use_gamma) {
          bool one_hot = false;
          if (LOG) {
            float mx = -INFINITY;
            for (uint32_t i = 0; i < d; ++i) mx = fmaxf(mx, col[i]);
            if (mx == -INFINITY) {
              one_hot = true;  // every parameter is 0: distributions.cpp:69-77
            } else {
              for (uint32_t i = 0; i < d; ++i) col[i] = __expf(__fadd_rn(col[i], -mx));
            }
          } else {
            float sum = 0.0f;
            for (uint32_t i = 0; i < d; ++i) sum = __fadd_rn(sum, col[i]);
            one_hot = !(sum > 0.0f);
          }
          if (one_hot) {
            U4 r = node_random(key, it.node_key, kStreamOneHot, 0u);
            uint32_t pick = uniform_below(r.x, d);
            for (uint32_t i = 0; i < d; ++i) col[i] = __uint_as_float(i == pick ? it.count : 0u);
          } else {
            // Conditional probabilities from suffix sums: outcome i takes Binomial(remaining,
            // g_i / sum_{j>=i} g_j). The smaller of (p, 1-p) is stored, negated when it is 1-p, so
            // that neither tail loses precision.
            float suffix = 0.0f;
            for (int i = (int)d - 1; i >= 0; --i) {
              float g = col[i];
              float tot = __fadd_rn(suffix, g);
              float v = 0.0f;
              if (tot > 0.0f) v = (g <= suffix) ? __fdividef(g, tot) : -__fdividef(suffix, tot);
              col[i] = v;
              suffix = tot;
            }
            uint32_t remaining = it.count;
            for (uint32_t i = 0; i < d; ++i) {
              float v = col[i];
              const bool flip = (__float_as_uint(v) >> 31) != 0u;
              uint32_t att = 0;
              uint32_t k = binomial_small_p(remaining, fabsf(v), key, it.node_key, i, att);
              stat[kStatBinomials] += att;
              if (flip) k = remaining - k;
              remaining -= k;
              col[i] = __uint_as_float(k);
            }
          }
        } else {
          // Polya urn for small counts: ball b copies a uniformly chosen earlier ball with
          // probability b / (A + b), else draws a fresh outcome with probability alpha_i / A.
          // Exactly Dirichlet-multinomial, no Gamma or binomial variates needed.
          stat[kStatUrnItems] += 1;
          for (uint32_t i = 0; i < d; ++i) col[i] = __uint_as_float(0u);
          float A;
          if (it.node >= 0) {
            uint32_t tot = 0;
            for (uint32_t i = 0; i < d; ++i) tot += a.T.slot_count[nbase + i];
            A = __fmaf_rn((float)d, a_prior, (float)tot);
          } else {
            A = __fmul_rn((float)d, a_prior);
          }
          uint64_t picks = 0;  // 8 bits per ball
          U4 r;
          for (uint32_t b = 0; b < it.count; ++b) {
            if ((b & 3u) == 0u) r = node_random(key, it.node_key, kStreamUrn, b >> 2);
            uint32_t bits = (b & 3u) == 0u ? r.x : (b & 3u) == 1u ? r.y : (b & 3u) == 2u ? r.z : r.w;
            uint32_t pick;
            if (!(A > 0.0f)) {
              // all parameters 0: the first ball is uniform, the rest follow it
              pick = b == 0u ? uniform_below(bits, d) : (uint32_t)(picks & 0xFFu);
            } else {
              float u = __fmul_rn(u01f(bits), __fadd_rn(A, (float)b));
              if (u >= A) {
                uint32_t j = (uint32_t)__fadd_rn(u, -A);
                if (j >= b) j = b - 1u;
                pick = (uint32_t)(picks >> (8u * j)) & 0xFFu;
              } else if (it.node >= 0) {
                float acc = 0.0f;
                pick = 0u;  // on fall-through (rounding): the last outcome with a positive parameter
                for (uint32_t i = 0; i < d; ++i) {
                  float al = __fadd_rn((float)a.T.slot_count[nbase + i], a_prior);
                  acc = __fadd_rn(acc, al);
                  if (al > 0.0f) pick = i;
                  if (u < acc) break;
                }
              } else {
                pick = (uint32_t)__fdividef(u, a_prior);
                if (pick >= d) pick = d - 1u;
              }
            }
            picks |= (uint64_t)pick << (8u * b);
            col[pick] = __uint_as_float(__float_as_uint(col[pick]) + 1u);
          }
        }
        // count what this item emits
        for (uint32_t i = 0; i < K; ++i) {
          if (__float_as_uint(col[i]) == 0u) continue;
          if (leaf_children) ++n_ent; else ++n_child;
        }
        if (T && __float_as_uint(col[K]) != 0u) ++n_ent;
      }

      // ---- compaction: block-wide prefix sums give every thread its output ranges --------------
      uint32_t tot;
      uint32_t off = block_exclusive_scan<BLOCK>((n_ent << 16) | n_child, sm.warp_scan, &tot);
      const uint32_t out_base = sm.n_out, ent_base = sm.n_entries;
      const uint32_t tot_child = tot & 0xFFFFu, tot_ent = tot >> 16;
      const bool overflow = (out_base + tot_child > a.cap_items) || (ent_base + tot_ent > a.cap_entries);
      if (overflow) {
        if (tid == 0) *a.error_flag = 1;
      } else if (active) {
        uint32_t co = out_base + (off & 0xFFFFu), eo = ent_base + (off >> 16);
        const uint32_t nbase = sm.nbase[tid];
        if (T) {
          uint32_t c = __float_as_uint(col[K]);
          if (c) {  // irv_node.cpp:128-132 / :63-68: ballots that stop at this node
            sc.ent_key[eo] = it.prefix;
            sc.ent_cnt[eo] = c;
            sc.ent_len[eo] = (uint8_t)depth;
            ++eo;
          }
        }
        for (uint32_t i = 0; i < K; ++i) {
          uint32_t c = __float_as_uint(col[i]);
          if (c == 0u) continue;
          uint32_t cand = it.node >= 0 ? (uint32_t)a.T.slot_cand[nbase + i] : nth_unused(it.used, nc, i);
          BallotKey<W> pk = it.prefix;
          key_set<W>(pk, (int)depth, cand);
          if (leaf_children) {
            sc.ent_key[eo] = pk;
            sc.ent_cnt[eo] = c;
            sc.ent_len[eo] = (uint8_t)(depth + 1);
            ++eo;
          } else {
            Item<W> ch;
            ch.prefix = pk;
            ch.node_key = child_node_key(it.node_key, cand);
            ch.count = c;
            ch.node = it.node >= 0 ? a.T.slot_child[nbase + i] : -1;
            ch.used = it.used | (1u << cand);
            ch.pad_ = 0;
            out[co++] = ch;
          }
        }
      }
      __syncthreads();
      if (tid == 0 && !overflow) {
        sm.n_out = out_base + tot_child;
        sm.n_entries = ent_base + tot_ent;
      }
      __syncthreads();
    }
    if (tid == 0) sm.n_in = leaf_children ? 0u : sm.n_out;
    cur ^= 1;
    __syncthreads();
  }
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
      for (uint32_t c = 0; c < nc; ++c) {
        if (elim >> c & 1u) continue;
        uint32_t t = sm.tally[c];
        if (t < mn) { mn = t; tied = 1u << c; ntied = 1; }
        else if (t == mn) { tied |= 1u << c; ++ntied; }
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
      sm.eliminated = elim | (1u << chosen);
    }
    __syncthreads();
    const uint32_t chosen = sm.chosen, eliminated = sm.eliminated;
    const bool any = sm.tally[chosen] != 0u;
    __syncthreads();
    if (!any) continue;  // nobody holds ballots for this candidate: nothing to redistribute
    // redistribution (irv_ballot.cpp:109-133)
    const uint32_t n_tot = n_ent + n_obs;
    for (uint32_t base = 0; base < n_tot; base += BLOCK) {
      uint32_t i = base + tid;
      bool valid = false;
      uint32_t cand = 0, cnt = 0;
      if (i < n_ent) {
        if (sc.ent_holder[i] == chosen) {
          cand = first_standing<W>(sc.ent_key[i], sc.ent_len[i], eliminated);
          sc.ent_holder[i] = (uint8_t)cand;
          cnt = sc.ent_cnt[i];
          valid = cand != 0xFFu;
          stat[kStatRereads] += 1;
        }
      } else if (i < n_tot) {
        uint32_t j = i - n_ent;
        if (sc.obs_holder[j] == chosen) {
          // an observed key stores at most nc-1 preferences; its length is implied by repetition
          cand = first_standing<W>(obs_key[j], nc - 1, eliminated);
          sc.obs_holder[j] = (uint8_t)cand;
          cnt = a.obs_cnt[j];
          valid = cand != 0xFFu;
          stat[kStatRereads] += 1;
        }
      }
      warp_tally_add(sm.tally, valid, cand, cnt);
    }
    __syncthreads();
  }
}

template <int W, int BLOCK, bool LOG>
__global__ void __launch_bounds__(BLOCK) simulate_kernel(const __grid_constant__ SimArgs a) {
  extern __shared__ __align__(16) float s_val[];
  __shared__ SharedFixed<BLOCK> sm;
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

    expand_election<W, BLOCK, LOG>(a, sc, sm, s_val, key, stat);
    __syncthreads();
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
