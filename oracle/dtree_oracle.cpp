/* dtree_oracle.cpp — CPU restatement ("port") of the reference's posterior-simulation
 * hot path, exported through oracle_api.h.
 *
 * TEST INFRASTRUCTURE ONLY: the checker for the CUDA path. Never linked into or
 * called from the product library (elections.dtree_b200/csrc); see oracle_api.h.
 *
 * Parity status: PINNED. The variate generators are libstdc++ <random>
 * (GCC 13.3 in this image; the reference depends on them un-vendored:
 * distributions.cpp:45-46,63-64,72-73, irv_ballot.cpp:49,100-102). This file calls
 * the same std:: distributions on a std::mt19937 in the same order as the
 * reference, so every output below — node counts, sampled (ballot,count) lists,
 * IRV elimination orders, win frequencies — is bit-identical to the reference's
 * own code compiled as oracle/_ref/libdtree_ref.so (tests/test_oracle_pin.py),
 * and reproduces the reference's golden vector (Wakehurst 2023,
 * tests/testthat/test-socialchoice.R:11-23) and its a0=0 known answer
 * (tests/testthat/test-minmaxdepth.R:1-27).
 *
 * The restatement is deliberately flat: nodes live in one arena addressed by
 * index, ballots are byte vectors, IRV keeps a cursor per ballot instead of
 * popping list nodes. Each function cites the reference lines it follows.
 */
#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "oracle_api.h"

namespace {

typedef std::vector<uint8_t> Ballot;
struct BallotCount {
  Ballot b;
  uint32_t count;
};

/* IRVParameters, irv_node.h:22-154 */
struct Params {
  uint32_t nc, minDepth, maxDepth;
  double a0;
  bool vd;
  std::vector<double> depthFactors;

  /* irv_node.cpp:13-25: F[maxDepth-1] = 1, F[d-1] = F[d] * (#outcomes at depth d) */
  void calculateDepthFactors() {
    depthFactors.assign(maxDepth, 0.0);
    double f = 1.0;
    for (int depth = (int)maxDepth - 1; depth >= 0; --depth) {
      uint32_t outcomes = nc - depth + ((uint32_t)depth >= minDepth ? 1 : 0);
      depthFactors[depth] = f;
      f *= outcomes;
    }
  }
};

/* IRVNode, tree_node.h:36-58 + irv_node.cpp:86-95: K = nc - depth children, K+1 counts. */
struct Node {
  uint32_t depth, K;
  std::vector<double> as;     /* as[K] = "ballot stops here" */
  std::vector<int32_t> child; /* arena index or -1 */
};

/* distributions.cpp:55-84 */
void rDirichlet(const double *a, uint32_t d, double *p, std::mt19937 *e) {
  double sum = 0.0;
  for (uint32_t i = 0; i < d; ++i) {
    std::gamma_distribution<double> g(a[i]);
    p[i] = g(*e);
    sum += p[i];
  }
  if (sum == 0.0) { /* :69-77: all gammas underflowed / were zero: one-hot at a uniform index */
    std::uniform_int_distribution<unsigned> rint(0, d - 1);
    unsigned idx = rint(*e);
    for (uint32_t i = 0; i < d; ++i) p[i] = 0.0;
    p[idx] = 1.0;
    return;
  }
  for (uint32_t i = 0; i < d; ++i) p[i] = p[i] / sum;
}

/* distributions.cpp:21-53 */
void rMultinomial(uint32_t N, const double *p, uint32_t d, uint32_t *out, std::mt19937 *e) {
  double norm = 0.0;
  for (uint32_t i = 0; i < d; ++i) norm += p[i];
  double sum_ps = 0.0;
  uint32_t n = N;
  for (uint32_t i = 0; i < d; ++i) {
    if (norm - (sum_ps + p[i]) == 0.0) { /* :37-41 last positive p takes the remainder */
      out[i] = n;
      for (uint32_t j = i + 1; j < d; ++j) out[j] = 0;
      break;
    }
    double pnorm = p[i] / (norm - sum_ps);
    std::binomial_distribution<unsigned> b(n, pnorm);
    out[i] = b(*e);
    n -= out[i];
    sum_ps += p[i];
  }
}

/* distributions.cpp:12-19 */
void rDirichletMultinomial(uint32_t N, const double *a, uint32_t d, uint32_t *out,
                           std::mt19937 *e) {
  double p[64];
  rDirichlet(a, d, p, e);
  rMultinomial(N, p, d, out, e);
}

struct Counters {
  uint64_t slots = 0;     /* S: outcome slots read in visited initialised nodes */
  uint64_t dirmult = 0;   /* Dirichlet-multinomial draws */
  uint64_t gammas = 0;    /* V_gamma */
  uint64_t binomials = 0; /* upper bound d-1 per draw */
  uint64_t rereads = 0;   /* R: ballots re-examined on redistribution */
};

struct Tree {
  Params P;
  std::vector<Node> nodes; /* nodes[0] = root */
  std::map<Ballot, uint32_t> observed;
  uint32_t nObservedTree = 0; /* dirichlet_tree.h:39 */
  size_t nObserved = 0;       /* R_tree.h:44 */
  std::mt19937 engine;
  Counters *ctr = nullptr;

  int32_t newNode(uint32_t depth) {
    Node n;
    n.depth = depth;
    n.K = P.nc - depth;
    n.as.assign(n.K + 1, 0.0);
    n.child.assign(n.K, -1);
    nodes.push_back(std::move(n));
    return (int32_t)nodes.size() - 1;
  }

  /* dirichlet_tree.h:151-156 */
  void setSeed(const std::string &seed) {
    std::seed_seq ss(seed.begin(), seed.end());
    engine.seed(ss);
    for (unsigned i = 1000; i; --i) engine();
  }

  /* dirichlet_tree.h:172-180 */
  void reset() {
    nodes.clear();
    newNode(0);
    observed.clear();
    nObservedTree = 0;
  }

  /* dirichlet_tree.h:182-193 then irv_node.cpp:176-221, recursion unrolled into a loop.
   * `path` starts as 0..nc-1 and is permuted by the same swaps, so slot j of a node at
   * depth d is candidate path[d+j]. The recursive call in the reference drops `count`
   * (irv_node.cpp:220), i.e. below the root the increment is 1; R always passes 1. */
  void update(const Ballot &b, uint32_t count) {
    observed[b] += count;
    nObservedTree += count;
    std::vector<uint32_t> path(P.nc);
    for (uint32_t i = 0; i < P.nc; ++i) path[i] = i;
    int32_t cur = 0;
    uint32_t inc = count;
    for (;;) {
      uint32_t depth = nodes[cur].depth, K = nodes[cur].K;
      if (depth == b.size()) { /* :191-194 */
        nodes[cur].as[K] += inc;
        return;
      }
      uint32_t next = b[depth], i = depth;
      while (path[i] != next) ++i; /* :203-205 */
      uint32_t slot = i - depth;
      nodes[cur].as[slot] += inc;
      if (K == 2) return; /* :210 */
      if (nodes[cur].child[slot] < 0) {
        int32_t c = newNode(depth + 1); /* may reallocate `nodes` */
        nodes[cur].child[slot] = c;
      }
      std::swap(path[depth], path[i]);
      cur = nodes[cur].child[slot];
      inc = 1; /* :220 (default argument) */
    }
  }

  void emit(std::vector<BallotCount> &out, const std::vector<uint32_t> &path, uint32_t len,
            uint32_t count) {
    BallotCount bc;
    bc.b.assign(path.begin(), path.begin() + len);
    bc.count = count;
    out.push_back(std::move(bc));
  }

  /* irv_node.cpp:27-84 */
  void lazy(uint32_t count, std::vector<uint32_t> &path, uint32_t depth,
            std::vector<BallotCount> &out, std::mt19937 *e) {
    double a0 = P.a0;
    if (P.vd) a0 = a0 * P.depthFactors[depth];
    uint32_t K = P.nc - depth;
    uint32_t d = K + (depth >= P.minDepth ? 1 : 0);
    if (depth == P.nc - 1 || depth == P.maxDepth) { /* :46-52 */
      emit(out, path, depth, count);
      return;
    }
    double a[64];
    uint32_t c[64];
    for (uint32_t i = 0; i < d; ++i) a[i] = a0;
    rDirichletMultinomial(count, a, d, c, e);
    if (ctr) { ctr->dirmult++; ctr->gammas += d; ctr->binomials += d - 1; }
    if (depth >= P.minDepth && c[d - 1] > 0) emit(out, path, depth, c[d - 1]); /* :63-68 */
    for (uint32_t i = 0; i < K; ++i) {
      if (c[i] == 0) continue;
      std::swap(path[depth], path[depth + i]);
      lazy(c[i], path, depth + 1, out, e);
      std::swap(path[depth], path[depth + i]);
    }
  }

  /* irv_node.cpp:107-174 */
  void sampleNode(int32_t ni, uint32_t count, std::vector<uint32_t> &path,
                  std::vector<BallotCount> &out, std::mt19937 *e) {
    uint32_t depth = nodes[ni].depth, K = nodes[ni].K;
    double a0 = P.a0;
    if (P.vd) a0 = a0 * P.depthFactors[depth];
    uint32_t d = K + (depth >= P.minDepth ? 1 : 0);
    double a[64];
    uint32_t c[64];
    for (uint32_t i = 0; i < d; ++i) a[i] = nodes[ni].as[i] + a0;
    rDirichletMultinomial(count, a, d, c, e);
    if (ctr) { ctr->dirmult++; ctr->gammas += d; ctr->binomials += d - 1; ctr->slots += d; }
    if (depth >= P.minDepth && c[K] > 0) emit(out, path, depth, c[K]); /* :128-132 */
    if (depth == P.maxDepth - 1) {                                     /* :136-151 */
      for (uint32_t i = 0; i < K; ++i) {
        if (c[i] == 0) continue;
        std::swap(path[depth], path[depth + i]);
        emit(out, path, depth + 1, c[i]);
        std::swap(path[depth], path[depth + i]);
      }
      return;
    }
    for (uint32_t i = 0; i < K; ++i) { /* :156-171 */
      if (c[i] == 0) continue;
      std::swap(path[depth], path[depth + i]);
      if (nodes[ni].child[i] < 0)
        lazy(c[i], path, depth + 1, out, e);
      else
        sampleNode(nodes[ni].child[i], c[i], path, out, e);
      std::swap(path[depth], path[depth + i]);
    }
  }

  /* dirichlet_tree.h:195-209 */
  std::vector<BallotCount> sample(uint32_t n, std::mt19937 *e) {
    if (!e) e = &engine;
    std::vector<uint32_t> path(P.nc);
    for (uint32_t i = 0; i < P.nc; ++i) path[i] = i;
    std::vector<BallotCount> out;
    sampleNode(0, n, path, out, e);
    return out;
  }

  /* dirichlet_tree.h:216-235 */
  std::vector<BallotCount> posteriorSet(uint32_t N, bool replace, std::mt19937 *e) {
    if (replace) return sample(N, e);
    if (nObservedTree > N) return {};
    std::vector<BallotCount> out;
    out.reserve(observed.size());
    for (const auto &kv : observed) out.push_back(BallotCount{kv.first, kv.second});
    std::vector<BallotCount> s = sample(N - nObservedTree, e);
    out.insert(out.end(), std::make_move_iterator(s.begin()), std::make_move_iterator(s.end()));
    return out;
  }
};

/* irv_ballot.cpp:42-138. The reference pops eliminated leading preferences off a
 * std::list; here each ballot keeps a cursor `pos` instead. Per-candidate groups
 * hold ballot indices in arrival order like tally_groups (:73-83). */
std::vector<uint32_t> socialChoiceIRV(const std::vector<BallotCount> &ballots, uint32_t nc,
                                      std::mt19937 *e, Counters *ctr = nullptr) {
  std::vector<uint32_t> out;
  std::vector<uint32_t> pos(ballots.size(), 0);
  std::vector<char> eliminated(nc, 0);
  std::vector<std::vector<uint32_t>> groups(nc);
  std::vector<uint32_t> tallies(nc, 0);
  for (uint32_t i = 0; i < ballots.size(); ++i) {
    if (ballots[i].b.empty()) continue; /* :55-56 */
    uint32_t fp = ballots[i].b[0];
    groups[fp].push_back(i);
    tallies[fp] += ballots[i].count;
  }
  std::vector<uint32_t> tied;
  for (uint32_t round = 0; round < nc; ++round) {
    uint32_t min_tally = std::numeric_limits<uint32_t>::max();
    for (uint32_t i = 0; i < nc; ++i) { /* :88-98 */
      if (!eliminated[i] && tallies[i] <= min_tally) {
        if (tallies[i] == min_tally) {
          tied.push_back(i);
        } else {
          tied.assign(1, i);
          min_tally = tallies[i];
        }
      }
    }
    /* :100-102 — one engine draw per round, tie or not */
    std::uniform_int_distribution<> r(0, (int)tied.size() - 1);
    uint32_t elim = tied[r(*e)];
    eliminated[elim] = 1;
    out.push_back(elim);
    for (uint32_t bi : groups[elim]) { /* :109-133 */
      const Ballot &b = ballots[bi].b;
      uint32_t p = pos[bi];
      while (p < b.size() && eliminated[b[p]]) ++p;
      pos[bi] = p;
      if (ctr) ctr->rereads++;
      if (p == b.size()) continue; /* exhausted: dropped */
      groups[b[p]].push_back(bi);
      tallies[b[p]] += ballots[bi].count;
    }
    groups[elim].clear();
  }
  return out;
}

Ballot csrBallot(const uint8_t *prefs, const uint64_t *offsets, uint64_t i) {
  return Ballot(prefs + offsets[i], prefs + offsets[i + 1]);
}

} /* namespace */

struct orc_tree {
  Tree T;
};
struct orc_ballots {
  std::vector<BallotCount> v;
};

extern "C" {

const char *orc_kind(void) { return "port"; }

orc_tree *orc_create(uint32_t nc, uint32_t min_depth, uint32_t max_depth, double a0, int vd,
                     const char *seed, size_t seed_len) {
  orc_tree *t = new orc_tree;
  t->T.P.nc = nc;
  t->T.P.minDepth = min_depth;
  t->T.P.maxDepth = max_depth;
  t->T.P.a0 = a0;
  t->T.P.vd = vd != 0;
  t->T.P.calculateDepthFactors();
  t->T.newNode(0);
  t->T.setSeed(std::string(seed, seed_len));
  return t;
}

void orc_destroy(orc_tree *t) { delete t; }

void orc_set_min_depth(orc_tree *t, uint32_t v) { t->T.P.minDepth = v; } /* irv_node.h:127: factors left stale */
void orc_set_max_depth(orc_tree *t, uint32_t v) {                         /* irv_node.h:134-137 */
  t->T.P.maxDepth = v;
  t->T.P.calculateDepthFactors();
}
void orc_set_a0(orc_tree *t, double v) { t->T.P.a0 = v; }
void orc_set_vd(orc_tree *t, int v) { t->T.P.vd = v != 0; }

uint32_t orc_depth_factors(orc_tree *t, double *out, uint32_t cap) {
  uint32_t n = t->T.P.depthFactors.size();
  for (uint32_t i = 0; i < n && i < cap; ++i) out[i] = t->T.P.depthFactors[i];
  return n;
}

void orc_reset(orc_tree *t) {
  t->T.reset();
  t->T.nObserved = 0; /* R_tree.cpp:119-123 */
}

void orc_update(orc_tree *t, const uint8_t *prefs, const uint64_t *offsets, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) {
    t->T.nObserved += 1;
    t->T.update(csrBallot(prefs, offsets, i), 1);
  }
}

uint64_t orc_n_observed(orc_tree *t) { return t->T.nObserved; }

static void dumpNode(const Tree &T, int32_t ni, std::vector<uint32_t> &path,
                     std::vector<uint32_t> &out) {
  const Node &n = T.nodes[ni];
  out.push_back(n.depth);
  for (uint32_t i = 0; i < n.depth; ++i) out.push_back(path[i]);
  out.push_back(n.K);
  for (uint32_t j = 0; j < n.K; ++j) out.push_back(path[n.depth + j]);
  for (uint32_t j = 0; j <= n.K; ++j) out.push_back((uint32_t)n.as[j]);
  for (uint32_t j = 0; j < n.K; ++j) out.push_back(n.child[j] >= 0);
  for (uint32_t j = 0; j < n.K; ++j) {
    if (n.child[j] < 0) continue;
    std::swap(path[n.depth], path[n.depth + j]);
    dumpNode(T, n.child[j], path, out);
    std::swap(path[n.depth], path[n.depth + j]);
  }
}

uint64_t orc_dump_nodes(orc_tree *t, uint32_t *buf, uint64_t cap) {
  std::vector<uint32_t> out, path(t->T.P.nc);
  for (uint32_t i = 0; i < t->T.P.nc; ++i) path[i] = i;
  dumpNode(t->T, 0, path, out);
  if (buf && cap >= out.size()) std::copy(out.begin(), out.end(), buf);
  return out.size();
}

orc_ballots *orc_observed(orc_tree *t) {
  orc_ballots *b = new orc_ballots;
  for (const auto &kv : t->T.observed) b->v.push_back(BallotCount{kv.first, kv.second});
  return b;
}

orc_ballots *orc_posterior_set(orc_tree *t, uint32_t n_ballots, int replace, uint32_t engine_seed) {
  std::mt19937 e(engine_seed); /* R_tree.cpp:215-216 */
  e.discard(e.state_size * 100);
  orc_ballots *b = new orc_ballots;
  b->v = t->T.posteriorSet(n_ballots, replace != 0, &e);
  return b;
}

orc_ballots *orc_sample_predictive(orc_tree *t, uint32_t n, const char *seed, size_t seed_len) {
  t->T.setSeed(std::string(seed, seed_len)); /* R_tree.cpp:157 */
  orc_ballots *b = new orc_ballots;
  b->v = t->T.sample(n, nullptr); /* :162 */
  return b;
}

uint64_t orc_ballots_size(const orc_ballots *b) { return b->v.size(); }
uint64_t orc_ballots_total_prefs(const orc_ballots *b) {
  uint64_t s = 0;
  for (const auto &bc : b->v) s += bc.b.size();
  return s;
}
void orc_ballots_export(const orc_ballots *b, uint8_t *prefs, uint64_t *offsets, uint32_t *counts) {
  uint64_t o = 0;
  offsets[0] = 0;
  for (uint64_t i = 0; i < b->v.size(); ++i) {
    for (uint8_t c : b->v[i].b) prefs[o++] = c;
    counts[i] = b->v[i].count;
    offsets[i + 1] = o;
  }
}
void orc_ballots_free(orc_ballots *b) { delete b; }

void orc_irv(const uint8_t *prefs, const uint64_t *offsets, const uint32_t *counts, uint64_t n,
             uint32_t nc, const char *seed, size_t seed_len, uint32_t *order_out) {
  std::vector<BallotCount> in(n);
  for (uint64_t i = 0; i < n; ++i) in[i] = BallotCount{csrBallot(prefs, offsets, i), counts[i]};
  std::string s(seed, seed_len);
  std::seed_seq ss(s.begin(), s.end()); /* R_social_choice.cpp:72-74 */
  std::mt19937 e(ss);
  e.discard(e.state_size * 100);
  std::vector<uint32_t> order = socialChoiceIRV(in, nc, &e);
  for (uint32_t i = 0; i < nc; ++i) order_out[i] = order[i];
}

/* R_tree.cpp:177-257 */
int orc_sample_posterior(orc_tree *t, uint32_t nElections, uint32_t nBallots, uint32_t nWinners,
                         int replace, uint32_t nThreads, const char *seed, size_t seed_len,
                         double *out, double *pool_seconds) {
  Tree &T = t->T;
  if (nBallots < T.nObserved) return 1; /* :183-186 */
  T.setSeed(std::string(seed, seed_len)); /* :188 */
  uint32_t nc = T.P.nc;
  std::vector<uint32_t> seeds; /* :193-197: nThreads+1 draws from the tree engine */
  for (uint32_t i = 0; i <= nThreads; ++i) seeds.push_back(T.engine());
  uint32_t batch = nElections <= 1 ? 0 : nElections / nThreads; /* :200-207 */
  uint32_t rem = nElections <= 1 ? nElections : nElections % nThreads;
  std::vector<std::vector<uint32_t>> winsPerThread(nThreads, std::vector<uint32_t>(nc, 0));
  auto work = [&](uint32_t idx, uint32_t size) { /* :213-229 */
    std::mt19937 e(seeds[idx]);
    e.discard(e.state_size * 100);
    for (uint32_t j = 0; j < size; ++j) {
      std::vector<BallotCount> election = T.posteriorSet(nBallots, replace != 0, &e);
      std::vector<uint32_t> order = socialChoiceIRV(election, nc, &e);
      for (uint32_t k = nc - nWinners; k < nc; ++k) winsPerThread[idx][order[k]] += 1; /* :245-253 */
    }
  };
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool; /* :232-242: workers 0..nThreads-2 get the remainder, caller runs the last batch */
  for (uint32_t i = 0; i + 1 < nThreads; ++i) pool.emplace_back(work, i, batch + (i < rem ? 1 : 0));
  work(nThreads - 1, batch);
  for (auto &th : pool) th.join();
  auto t1 = std::chrono::steady_clock::now();
  if (pool_seconds) *pool_seconds = std::chrono::duration<double>(t1 - t0).count();
  for (uint32_t k = 0; k < nc; ++k) {
    double s = 0;
    for (uint32_t i = 0; i < nThreads; ++i) s += winsPerThread[i][k];
    out[k] = s / nElections; /* :255 */
  }
  return 0;
}

/* ---- port-only extras (not part of oracle_api.h; no reference counterpart) ---- */

/* Roofline denominators of SURVEY.md §8(d) for ONE election drawn with the
 * per-thread engine convention: out = {S, E, R, V_gamma, dirmult draws, binomials(max)}. */
void orc_election_stats(orc_tree *t, uint32_t n_ballots, int replace, uint32_t engine_seed,
                        uint64_t *out) {
  Counters c;
  t->T.ctr = &c;
  std::mt19937 e(engine_seed);
  e.discard(e.state_size * 100);
  std::vector<BallotCount> el = t->T.posteriorSet(n_ballots, replace != 0, &e);
  t->T.ctr = nullptr;
  socialChoiceIRV(el, t->T.P.nc, &e, &c);
  out[0] = c.slots;
  out[1] = el.size();
  out[2] = c.rereads;
  out[3] = c.gammas;
  out[4] = c.dirmult;
  out[5] = c.binomials;
}

/* Is `order` an elimination order the reference's rule could produce on these
 * ballots, i.e. is every eliminated candidate among the standing candidates with the
 * minimal tally at its round (irv_ballot.cpp:88-102)? Returns 1/0; *tie_rounds =
 * number of rounds (with >= 2 standing) where more than one candidate tied. */
int orc_irv_check_order(const uint8_t *prefs, const uint64_t *offsets, const uint32_t *counts,
                        uint64_t n, uint32_t nc, const uint32_t *order, uint32_t *tie_rounds) {
  std::vector<uint32_t> pos(n, 0);
  std::vector<char> eliminated(nc, 0);
  uint32_t ties = 0;
  for (uint32_t round = 0; round < nc; ++round) {
    std::vector<uint64_t> tallies(nc, 0);
    for (uint64_t i = 0; i < n; ++i) {
      uint64_t p = offsets[i] + pos[i];
      while (p < offsets[i + 1] && eliminated[prefs[p]]) ++p;
      pos[i] = p - offsets[i];
      if (p < offsets[i + 1]) tallies[prefs[p]] += counts[i];
    }
    uint64_t mn = UINT64_MAX;
    uint32_t nmin = 0;
    for (uint32_t c = 0; c < nc; ++c)
      if (!eliminated[c]) {
        if (tallies[c] < mn) { mn = tallies[c]; nmin = 1; }
        else if (tallies[c] == mn) ++nmin;
      }
    uint32_t el = order[round];
    if (el >= nc || eliminated[el] || tallies[el] != mn) {
      if (tie_rounds) *tie_rounds = ties;
      return 0;
    }
    if (nmin > 1) ++ties;
    eliminated[el] = 1;
  }
  if (tie_rounds) *tie_rounds = ties;
  return 1;
}

} /* extern "C" */
