// R_tree_b200.cpp — the Rcpp layer of elections.dtree re-pointed at libdtree_b200.so.
//
// Drop-in replacement for src/R_tree.{h,cpp} + src/R_social_choice.cpp + src/RcppModules.cpp of the
// reference: same module name, class name, constructor, properties, methods and exported function
// (src/RcppModules.cpp:15-34, src/RcppExports.cpp:16-27), so R/dtree.R and R/social_choice.R run
// unchanged. What changes is the body: candidate-name <-> index mapping, validation and warnings
// stay here (as in R_tree.cpp:13-153); everything below is one call into the C ABI of
// include/dtree_b200.h. Cannot be built in this repo's image (no R, no Rcpp.h);
// tests/test_abi_cpu.py syntax-checks it against tests/stubs/Rcpp.h. See INTEGRATION.md.
#include <Rcpp.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "dtree_b200.h"

namespace {

void check(int rc) {
  if (rc != DTREE_OK) Rcpp::stop(dtree_last_error());  // same route to an R error as R_tree.cpp:30-31
}

struct Csr {
  std::vector<uint8_t> prefs;
  std::vector<uint64_t> offsets{0};
};

int poll_interrupt(void *) {
  // RcppThread::checkUserInterrupt() of R_tree.cpp:222: called on R's main thread between waves.
  try {
    Rcpp::checkUserInterrupt();
  } catch (...) {
    return 1;
  }
  return 0;
}

}  // namespace

class RDirichletTree {
 private:
  dtree_t *tree = nullptr;
  Rcpp::CharacterVector candidateVector{};
  std::unordered_map<std::string, size_t> candidateMap{};

  // R_tree.cpp:13-42
  Csr parseBallotList(Rcpp::List bs) {
    Csr out;
    for (R_xlen_t i = 0; i < bs.size(); ++i) {
      Rcpp::CharacterVector namePrefs = bs[i];
      for (R_xlen_t j = 0; j < namePrefs.size(); ++j) {
        std::string cName = Rcpp::as<std::string>(namePrefs[j]);
        auto it = candidateMap.find(cName);
        if (it == candidateMap.end()) Rcpp::stop("Unknown candidate encountered in ballot!");
        out.prefs.push_back((uint8_t)it->second);
      }
      out.offsets.push_back(out.prefs.size());
    }
    return out;
  }

 public:
  // R_tree.cpp:44-60. `seed_` seeded the tree's own mt19937; every sampling call passes its own seed.
  RDirichletTree(Rcpp::CharacterVector candidates, unsigned minDepth_, unsigned maxDepth_, double a0_, bool vd_,
                 std::string seed_) {
    for (R_xlen_t i = 0; i < candidates.size(); ++i) {
      std::string cName = Rcpp::as<std::string>(candidates[i]);
      candidateVector.push_back(cName);
      candidateMap[cName] = (size_t)i;
    }
    check(dtree_create((uint32_t)candidates.size(), minDepth_, maxDepth_, a0_, vd_ ? 1 : 0, -1, &tree));
    dtree_set_abort_callback(tree, poll_interrupt, nullptr);
  }
  ~RDirichletTree() { dtree_destroy(tree); }  // R_tree.cpp:63-66

  // Getters, R_tree.cpp:69-84
  unsigned getNCandidates() { return dtree_n_candidates(tree); }
  unsigned getMinDepth() { return dtree_min_depth(tree); }
  unsigned getMaxDepth() { return dtree_max_depth(tree); }
  double getA0() { return dtree_a0(tree); }
  bool getVD() { return dtree_vd(tree) != 0; }
  Rcpp::CharacterVector getCandidates() { return candidateVector; }

  // Setters, R_tree.cpp:87-116
  void setMinDepth(unsigned minDepth_) {
    if (minDepth_ > dtree_max_depth(tree)) Rcpp::stop("Cannot set `minDepth` to a value larger than `maxDepth`.");
    check(dtree_set_min_depth(tree, minDepth_));
    uint64_t mask = dtree_observed_depth_mask(tree);
    for (unsigned d = 1; d < minDepth_ && d < 64; ++d)
      if (mask >> d & 1) {
        Rcpp::warning(
            "Ballots with fewer than `minDepth` preferences specified have been observed. Some sampling "
            "techniques could now exhibit undefined behaviour. A Dirichlet Posterior can no longer reduce to a "
            "tree of height 1. Consider setting `minDepth` to a value lower than the length of the smallest "
            "ballot.");
        break;
      }
  }
  void setMaxDepth(unsigned maxDepth_) {
    if (maxDepth_ < dtree_min_depth(tree)) Rcpp::stop("Cannot set `maxDepth` to a value less than `minDepth`.");
    check(dtree_set_max_depth(tree, maxDepth_));
  }
  void setA0(double a0_) { check(dtree_set_a0(tree, a0_)); }
  void setVD(bool vd_) { check(dtree_set_vd(tree, vd_ ? 1 : 0)); }

  void reset() { check(dtree_reset(tree)); }  // R_tree.cpp:119-123

  // R_tree.cpp:125-153
  void update(Rcpp::List ballots) {
    Csr b = parseBallotList(ballots);
    uint64_t n_short = 0;
    check(dtree_update(tree, b.prefs.data(), b.offsets.data(), b.offsets.size() - 1, &n_short));
    if (n_short)
      Rcpp::warning(
          "Updating a Dirichlet-tree distribution with a ballot specifying fewer than `minDepth` preferences. "
          "This introduces undefined behaviour to the sampling methods, and the resulting posterior can no "
          "longer reduce to a Dirichlet distribution when using the `vd` option. Consider setting `minDepth` to a "
          "value lower than the length of the smallest ballot.");
  }

  // R_tree.cpp:155-175. The reference pushes count copies of every ballot onto an Rcpp::List (O(n^2));
  // R then re-aggregates them (R/dtree.R:463-468). The list is expanded here only to keep the R code unchanged.
  Rcpp::List samplePredictive(unsigned nSamples, std::string seed) {
    dtree_ballots_t *b = nullptr;
    check(dtree_sample_predictive(tree, nSamples, seed.data(), seed.size(), &b));
    uint64_t n = dtree_ballots_size(b);
    std::vector<uint8_t> prefs(dtree_ballots_total_prefs(b) + 1);
    std::vector<uint64_t> offsets(n + 1);
    std::vector<uint32_t> counts(n + 1);
    dtree_ballots_export(b, prefs.data(), offsets.data(), counts.data());
    dtree_ballots_free(b);
    Rcpp::List out(nSamples);
    R_xlen_t k = 0;
    for (uint64_t i = 0; i < n; ++i) {
      Rcpp::CharacterVector rBallot(offsets[i + 1] - offsets[i]);
      for (uint64_t j = offsets[i]; j < offsets[i + 1]; ++j) rBallot[j - offsets[i]] = candidateVector[prefs[j]];
      for (uint32_t c = 0; c < counts[i]; ++c) out[k++] = rBallot;
    }
    return out;
  }

  // R_tree.cpp:177-257. nThreads is accepted and ignored: the reference's thread pool over elections
  // (:200-242) becomes a pool over the visible GPUs, each taking a contiguous range of elections.
  Rcpp::NumericVector samplePosterior(unsigned nElections, unsigned nBallots, unsigned nWinners, bool replace,
                                      unsigned nThreads, std::string seed) {
    (void)nThreads;
    size_t nc = getNCandidates();
    std::vector<uint64_t> wins(nc, 0);
    int nDevices = dtree_device_count();
    if (nDevices > 1)
      check(dtree_sample_posterior_multi(tree, nullptr, nDevices, 0, nElections, nBallots, nWinners, replace ? 1 : 0,
                                         seed.data(), seed.size(), wins.data()));
    else
      check(dtree_sample_posterior(tree, 0, nElections, nBallots, nWinners, replace ? 1 : 0, seed.data(), seed.size(),
                                   wins.data(), nullptr));
    Rcpp::NumericVector out(nc);
    out.names() = candidateVector;
    for (size_t i = 0; i < nc; ++i) out[i] = (double)wins[i] / nElections;  // R_tree.cpp:255
    return out;
  }
};

// src/R_social_choice.cpp:16-95
// [[Rcpp::export]]
Rcpp::List social_choice_irv(Rcpp::List bs, unsigned nWinners, Rcpp::CharacterVector candidates, std::string seed) {
  std::unordered_map<std::string, size_t> c2Index{};
  std::vector<std::string> cNames{};
  for (R_xlen_t i = 0; i < candidates.size(); ++i) {
    std::string cName = Rcpp::as<std::string>(candidates[i]);
    if (c2Index.count(cName) == 0) {
      c2Index[cName] = cNames.size();
      cNames.push_back(cName);
    }
  }
  // The reference de-duplicates ballots with an O(n^2) search (:52-62); the kernel does not need that.
  Csr in;
  std::vector<uint32_t> counts;
  for (R_xlen_t i = 0; i < bs.size(); ++i) {
    if (Rf_isNull(bs[i])) continue;  // :40-41
    Rcpp::CharacterVector bNames = bs[i];
    for (R_xlen_t j = 0; j < bNames.size(); ++j) {
      std::string cName = Rcpp::as<std::string>(bNames[j]);
      if (c2Index.count(cName) == 0) Rcpp::stop("Invalid candidate found during social-choice evaluation.");
      in.prefs.push_back((uint8_t)c2Index[cName]);
    }
    in.offsets.push_back(in.prefs.size());
    counts.push_back(1);
  }
  if (nWinners < 1 || nWinners >= cNames.size()) Rcpp::stop("`nWinners` must be >= 1 and <= the number of candidates.");
  if (counts.empty()) Rcpp::stop("No valid ballots for the IRV social choice function.");
  std::vector<uint32_t> order(cNames.size());
  check(dtree_irv(in.prefs.data(), in.offsets.data(), counts.data(), counts.size(), (uint32_t)cNames.size(),
                  seed.data(), seed.size(), -1, order.data(), nullptr));
  Rcpp::CharacterVector elimination_order{}, winners{};
  for (size_t i = 0; i < cNames.size() - nWinners; ++i) elimination_order.push_back(cNames[order[i]]);
  for (size_t i = cNames.size() - nWinners; i < cNames.size(); ++i) winners.push_back(cNames[order[i]]);
  Rcpp::List out;
  out["elimination_order"] = elimination_order;
  out["winners"] = winners;
  return out;
}

// src/RcppModules.cpp:15-34, unchanged
RCPP_MODULE(dirichlet_tree_module) {
  Rcpp::class_<RDirichletTree>("RDirichletTree")
      .constructor<Rcpp::CharacterVector, unsigned, unsigned, double, bool, std::string>()
      .property("n_candidates", &RDirichletTree::getNCandidates)
      .property("a0", &RDirichletTree::getA0, &RDirichletTree::setA0)
      .property("min_depth", &RDirichletTree::getMinDepth, &RDirichletTree::setMinDepth)
      .property("max_depth", &RDirichletTree::getMaxDepth, &RDirichletTree::setMaxDepth)
      .property("vd", &RDirichletTree::getVD, &RDirichletTree::setVD)
      .property("candidates", &RDirichletTree::getCandidates)
      .method("reset", &RDirichletTree::reset)
      .method("update", &RDirichletTree::update)
      .method("sample_predictive", &RDirichletTree::samplePredictive)
      .method("sample_posterior", &RDirichletTree::samplePosterior);
}
