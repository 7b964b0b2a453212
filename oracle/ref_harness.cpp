/* ref_harness.cpp — drives the UNMODIFIED reference core through oracle_api.h.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_api.h). This file contains no algorithm:
 * it #includes the reference headers from /root/reference/src (path given with
 * -I by oracle/Makefile), is linked with the reference's own
 * distributions.cpp / irv_ballot.cpp / irv_node.cpp compiled where they lie, and
 * is built with -fno-access-control so that the node dump can read
 * DirichletTree::root / ::observed and IRVNode::as / ::children.
 *
 * The only restated code is the Rcpp-free half of
 * RDirichletTree::samplePosterior (R_tree.cpp:192-255), because R_tree.cpp
 * itself needs Rcpp.h, which is not in this image.
 */
#include <algorithm>
#include <chrono>
#include <functional>
#include <thread>

#include "dirichlet_tree.h"
#include "irv_ballot.h"
#include "irv_node.h"
#include "oracle_api.h"

typedef DirichletTree<IRVNode, IRVBallot, IRVParameters> RefTree;

struct orc_tree {
  RefTree *tree;
  IRVParameters *params;
  size_t nObserved = 0; /* R_tree.h:44 */
};

struct orc_ballots {
  std::list<IRVBallotCount> l;
};

extern "C" {

const char *orc_kind(void) { return "reference"; }

orc_tree *orc_create(uint32_t nc, uint32_t min_depth, uint32_t max_depth, double a0, int vd,
                     const char *seed, size_t seed_len) {
  orc_tree *t = new orc_tree;
  t->params = new IRVParameters(nc, min_depth, max_depth, a0, vd != 0);
  t->tree = new RefTree(t->params, std::string(seed, seed_len));
  return t;
}

void orc_destroy(orc_tree *t) {
  delete t->tree->getParameters();
  delete t->tree;
  delete t;
}

void orc_set_min_depth(orc_tree *t, uint32_t v) { t->params->setMinDepth(v); }
void orc_set_max_depth(orc_tree *t, uint32_t v) { t->params->setMaxDepth(v); }
void orc_set_a0(orc_tree *t, double v) { t->params->setA0(v); }
void orc_set_vd(orc_tree *t, int v) { t->params->setVD(v != 0); }

uint32_t orc_depth_factors(orc_tree *t, double *out, uint32_t cap) {
  uint32_t n = t->params->depthFactors.size();
  for (uint32_t i = 0; i < n && i < cap; ++i) out[i] = t->params->depthFactors[i];
  return n;
}

void orc_reset(orc_tree *t) {
  t->tree->reset();
  t->nObserved = 0;
}

void orc_update(orc_tree *t, const uint8_t *prefs, const uint64_t *offsets, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) {
    std::list<unsigned> p;
    for (uint64_t j = offsets[i]; j < offsets[i + 1]; ++j) p.push_back(prefs[j]);
    IRVBallotCount bc(IRVBallot(std::move(p)), 1); /* R_tree.cpp:38 */
    t->nObserved += bc.second;                     /* R_tree.cpp:149 */
    t->tree->update(bc);                           /* R_tree.cpp:150 */
  }
}

uint64_t orc_n_observed(orc_tree *t) { return t->nObserved; }

static void dump_node(IRVNode *node, std::vector<unsigned> path, std::vector<uint32_t> &out) {
  unsigned depth = node->depth, K = node->nChildren;
  out.push_back(depth);
  for (unsigned i = 0; i < depth; ++i) out.push_back(path[i]);
  out.push_back(K);
  for (unsigned j = 0; j < K; ++j) out.push_back(path[depth + j]);
  for (unsigned j = 0; j <= K; ++j) out.push_back((uint32_t)node->as[j]);
  for (unsigned j = 0; j < K; ++j) out.push_back(node->children[j] != nullptr);
  for (unsigned j = 0; j < K; ++j) {
    if (node->children[j] == nullptr) continue;
    std::swap(path[depth], path[depth + j]);
    dump_node(node->children[j], path, out);
    std::swap(path[depth], path[depth + j]);
  }
}

uint64_t orc_dump_nodes(orc_tree *t, uint32_t *buf, uint64_t cap) {
  std::vector<uint32_t> out;
  dump_node(t->tree->root, t->params->defaultPath(), out);
  if (buf && cap >= out.size()) std::copy(out.begin(), out.end(), buf);
  return out.size();
}

orc_ballots *orc_observed(orc_tree *t) {
  orc_ballots *b = new orc_ballots;
  b->l = std::list<IRVBallotCount>(t->tree->observed.begin(), t->tree->observed.end());
  return b;
}

orc_ballots *orc_posterior_set(orc_tree *t, uint32_t n_ballots, int replace, uint32_t engine_seed) {
  std::mt19937 e(engine_seed);
  e.discard(e.state_size * 100);
  orc_ballots *b = new orc_ballots;
  b->l = t->tree->posteriorSet(n_ballots, replace != 0, &e);
  return b;
}

orc_ballots *orc_sample_predictive(orc_tree *t, uint32_t n, const char *seed, size_t seed_len) {
  t->tree->setSeed(std::string(seed, seed_len));
  orc_ballots *b = new orc_ballots;
  b->l = t->tree->sample(n);
  return b;
}

uint64_t orc_ballots_size(const orc_ballots *b) { return b->l.size(); }

uint64_t orc_ballots_total_prefs(const orc_ballots *b) {
  uint64_t s = 0;
  for (const auto &bc : b->l) s += bc.first.nPreferences();
  return s;
}

void orc_ballots_export(const orc_ballots *b, uint8_t *prefs, uint64_t *offsets, uint32_t *counts) {
  uint64_t i = 0, o = 0;
  offsets[0] = 0;
  for (const auto &bc : b->l) {
    for (unsigned c : bc.first.preferences) prefs[o++] = (uint8_t)c;
    counts[i] = bc.second;
    offsets[++i] = o;
  }
}

void orc_ballots_free(orc_ballots *b) { delete b; }

void orc_irv(const uint8_t *prefs, const uint64_t *offsets, const uint32_t *counts, uint64_t n,
             uint32_t nc, const char *seed, size_t seed_len, uint32_t *order_out) {
  std::list<IRVBallotCount> in;
  for (uint64_t i = 0; i < n; ++i) {
    std::list<unsigned> p;
    for (uint64_t j = offsets[i]; j < offsets[i + 1]; ++j) p.push_back(prefs[j]);
    in.emplace_back(IRVBallot(std::move(p)), counts[i]);
  }
  std::string s(seed, seed_len);
  std::seed_seq ss(s.begin(), s.end()); /* R_social_choice.cpp:72-74 */
  std::mt19937 e(ss);
  e.discard(e.state_size * 100);
  std::vector<unsigned> order = socialChoiceIRV(in, nc, &e);
  for (uint32_t i = 0; i < nc; ++i) order_out[i] = order[i];
}

/* R_tree.cpp:177-257, Rcpp and checkUserInterrupt removed, nothing else changed. */
int orc_sample_posterior(orc_tree *t, uint32_t nElections, uint32_t nBallots, uint32_t nWinners,
                         int replace_, uint32_t nThreads, const char *seed, size_t seed_len,
                         double *out, double *pool_seconds) {
  bool replace = replace_ != 0;
  RefTree *tree = t->tree;
  if (nBallots < t->nObserved) return 1;

  tree->setSeed(std::string(seed, seed_len));
  size_t nCandidates = t->params->getNCandidates();

  std::mt19937 *treeGen = tree->getEnginePtr();
  std::vector<unsigned> seeds{};
  for (unsigned i = 0; i <= nThreads; ++i) seeds.push_back((*treeGen)());

  unsigned batchSize, batchRemainder;
  if (nElections <= 1) {
    batchSize = 0;
    batchRemainder = nElections;
  } else {
    batchSize = nElections / nThreads;
    batchRemainder = nElections % nThreads;
  }

  std::vector<std::vector<std::vector<unsigned>>> results(nThreads);

  auto processBatch = [&](size_t thread_idx, size_t size) -> void {
    std::mt19937 e(seeds[thread_idx]);
    e.discard(e.state_size * 100);
    results[thread_idx].resize(size);
    for (unsigned j = 0; j < size; ++j) {
      std::list<IRVBallotCount> election = tree->posteriorSet(nBallots, replace, &e);
      results[thread_idx][j] = socialChoiceIRV(election, nCandidates, &e);
    }
  };

  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool(nThreads - 1);
  for (unsigned i = 0; i < nThreads - 1; ++i)
    pool[i] = std::thread(std::bind(processBatch, i, batchSize + (i < batchRemainder)));
  processBatch(nThreads - 1, batchSize);
  std::for_each(pool.begin(), pool.end(), [](std::thread &th) { th.join(); });
  auto t1 = std::chrono::steady_clock::now();
  if (pool_seconds) *pool_seconds = std::chrono::duration<double>(t1 - t0).count();

  for (size_t k = 0; k < nCandidates; ++k) out[k] = 0.0;
  for (unsigned i = 0; i < nThreads; ++i)
    for (unsigned j = 0; j < results[i].size(); ++j)
      for (unsigned k = nCandidates - nWinners; k < nCandidates; ++k)
        out[results[i][j][k]] = out[results[i][j][k]] + 1;
  for (size_t k = 0; k < nCandidates; ++k) out[k] = out[k] / nElections;
  return 0;
}

} /* extern "C" */
