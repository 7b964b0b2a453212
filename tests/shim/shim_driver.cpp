// shim_driver.cpp — runs the Rcpp shim itself (elections.dtree_b200/host/r/R_tree_b200.cpp, the file a maintainer
// would drop into the R package) against tests/stubs/Rcpp.h, replaying the scenarios of the reference's testthat
// suite that reach the Rcpp layer, with the arguments R/dtree.R would pass:
//   tests/testthat/test-posterior.R:1-93, test-update_and_reset.R:1-83, test-minmaxdepth.R:1-59,
//   test-determinism.R:5-42, test-socialchoice.R:11-23 (Wakehurst golden vector, file given as argv[1]).
// R-side checks (R/dtree.R argument validation, prefio conversions) are not part of the shim and not replayed.
// Exit status 0 and "shim ok: N checks" on success. Built and run by tests/test_gpu_shim.py on the GPU box.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <sstream>

#include "../../elections.dtree_b200/host/r/R_tree_b200.cpp"

static int n_checks = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    ++n_checks;                                                            \
    if (!(cond)) {                                                         \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                        \
    }                                                                      \
  } while (0)

static Rcpp::CharacterVector letters(int n) {
  Rcpp::CharacterVector v;
  for (int i = 0; i < n; ++i) v.push_back(std::string(1, (char)('A' + i)));
  return v;
}

static Rcpp::List ballots(const std::vector<std::vector<int>> &bs) {  // 1-based candidate numbers, like the R tests
  Rcpp::List out;
  for (const auto &b : bs) {
    Rcpp::CharacterVector v;
    for (int c : b) v.push_back(std::string(1, (char)('A' + c - 1)));
    out.push_back(v);
  }
  return out;
}

static bool throws(const std::function<void()> &f, std::string *msg = nullptr) {
  try {
    f();
  } catch (const Rcpp::exception &e) {
    if (msg) *msg = e.what();
    return true;
  }
  return false;
}

static std::vector<double> probs(Rcpp::NumericVector v) { return *v.v.num; }

int main(int argc, char **argv) {
  auto &warnings = Rcpp::stub::warnings();

  {  // test-posterior.R:1-22 — the posterior shifts after observing data, for vd = FALSE and TRUE
    RDirichletTree t(letters(4), 0, 4, 1.0, false, "ABCDEFGHIJ");
    auto prior = probs(t.samplePosterior(1000, 10, 1, false, 2, "SEEDAAAAAA"));
    for (int i = 0; i < 5; ++i) t.update(ballots({{1, 2, 3, 4}}));
    auto post_dt = probs(t.samplePosterior(1000, 10, 1, false, 2, "SEEDBBBBBB"));
    t.setVD(true);
    auto post_d = probs(t.samplePosterior(1000, 10, 1, false, 2, "SEEDCCCCCC"));
    CHECK(post_d[0] > prior[0] && post_dt[0] > prior[0]);
    for (int c = 1; c < 4; ++c) CHECK(post_d[c] < prior[c] && post_dt[c] < prior[c]);
    double s = 0;
    for (double p : post_dt) s += p;
    CHECK(std::fabs(s - 1.0) < 1e-12);  // R_tree.cpp:245-255: probabilities sum to n_winners
    auto v = t.samplePosterior(10, 10, 1, false, 2, "x");
    CHECK(v.names_vector() == std::vector<std::string>({"A", "B", "C", "D"}));  // names in constructor order, :245-246
  }
  {  // test-posterior.R:24-29 — a0 = 0 on an empty tree: win probabilities nearly uniform
    RDirichletTree t(letters(10), 0, 10, 0.0, false, "s");
    auto p = probs(t.samplePosterior(1000, 10, 1, false, 2, "UNIFORMAAA"));
    double m = 0, var = 0;
    for (double x : p) m += x / 10;
    for (double x : p) var += (x - m) * (x - m) / 9;
    CHECK(var < 1e-2);
  }
  {  // test-posterior.R:31-44 — n_ballots below the number observed is an error (R_tree.cpp:183-186)
    RDirichletTree t(letters(10), 0, 10, 1.0, false, "s");
    t.update(ballots({{1, 2, 3, 4, 5, 6, 7, 8, 9, 10}}));
    t.update(ballots({{10, 9, 8, 7, 6, 5, 4, 3, 2, 1}}));
    std::string msg;
    CHECK(throws([&] { t.samplePosterior(1, 1, 1, false, 2, "s"); }, &msg));
    CHECK(msg.find("nBallots") != std::string::npos);
    CHECK(throws([&] { t.samplePosterior(1, 1, 1, true, 2, "s"); }));  // the reference stops even with replace
  }
  {  // test-posterior.R:60-69 — any n_threads is accepted; one election on one thread still runs one election
     // (the reference's n_elections = 1, n_threads = 1 runs NONE, R_tree.cpp:201-207: deliberately not reproduced)
    RDirichletTree t(letters(3), 0, 3, 1.0, false, "s");
    for (unsigned threads : {1u, 2u, 1000000u}) {
      auto p = probs(t.samplePosterior(2, 2, 1, false, threads, "THREADSAAA"));
      CHECK(p[0] + p[1] + p[2] == 1.0);
    }
    auto p1 = probs(t.samplePosterior(1, 2, 1, false, 1, "ONEELECTIO"));
    CHECK(p1[0] + p1[1] + p1[2] == 1.0);
  }
  {  // test-posterior.R:71-93 — with replacement more than one candidate can win
    RDirichletTree t(letters(3), 0, 3, 1.0, false, "s");
    t.update(ballots({{1, 2, 3}, {1, 2, 3}, {3, 2, 1}}));
    auto p = probs(t.samplePosterior(1000, 3, 1, true, 2, "REPLACEAAA"));
    CHECK((p[0] > 0) + (p[1] > 0) + (p[2] > 0) > 1);
  }
  {  // test-update_and_reset.R:1-30 — update raises P(win), reset restores the prior (n_candidates 3..9)
    for (int n = 3; n <= 9; ++n) {
      RDirichletTree t(letters(n), n - 1, n - 1, 1.0, false, "s");
      std::vector<int> b;
      for (int c = 1; c <= n; ++c) b.push_back(c);
      t.update(ballots({b}));
      double p1 = probs(t.samplePosterior(2000, 5, 1, false, 2, "UPDATERESE"))[0];
      t.reset();
      double p0 = probs(t.samplePosterior(2000, 5, 1, false, 2, "UPDATERESE"))[0];
      CHECK(p1 > p0);
    }
  }
  {  // test-update_and_reset.R:32-57 — warnings for ballots shorter than min_depth, and on raising min_depth
    RDirichletTree t(letters(26), 3, 10, 0.0, false, "s");
    warnings.clear();
    t.update(ballots({{1, 2}}));
    CHECK(warnings.size() == 1 && warnings[0].find("fewer than `minDepth`") != std::string::npos);
    t.setMinDepth(2);
    CHECK(warnings.size() == 1);
    t.setMinDepth(3);
    CHECK(warnings.size() == 2);
    warnings.clear();
  }
  {  // test-update_and_reset.R:59-72 — an unknown candidate is an error and changes nothing
    RDirichletTree t(letters(3), 0, 3, 1.0, false, "s");
    std::string msg;
    CHECK(throws([&] { t.update(ballots({{1, 2}, {4}})); }, &msg));
    CHECK(msg == "Unknown candidate encountered in ballot!");
    auto p = probs(t.samplePosterior(100, 0 + 5, 1, false, 2, "s"));  // still no observation: 5 ballots are enough
    CHECK(p[0] + p[1] + p[2] == 1.0);
    Rcpp::List bad;  // a numeric vector where a character vector is expected (what Rcpp itself rejects)
    bad.push_back(Rcpp::NumericVector(2));
    CHECK(throws([&] { t.update(bad); }));
  }
  {  // test-minmaxdepth.R:1-27 — a0 = 0, max_depth = 3, one observed ballot: the bootstrap returns its first 3 preferences
    RDirichletTree t(letters(5), 0, 3, 0.0, false, "s");
    t.update(ballots({{3, 2, 1, 4, 5}}));
    Rcpp::List s = t.samplePredictive(1, "BOOTSTRAPA");
    CHECK(s.size() == 1);
    Rcpp::CharacterVector b = s[0];
    CHECK(b.size() == 3 && b.at(0) == "C" && b.at(1) == "B" && b.at(2) == "A");
  }
  {  // test-minmaxdepth.R:29-42, test-sampling.R:1-19 — sampled lengths within [min_depth, max_depth]
    RDirichletTree t(letters(10), 2, 8, 1.0, false, "s");
    Rcpp::List s = t.samplePredictive(1000, "LENGTHSAAA");
    CHECK(s.size() == 1000);
    for (R_xlen_t i = 0; i < s.size(); ++i) {
      Rcpp::CharacterVector b = s[i];
      CHECK(b.size() >= 2 && b.size() <= 8);
    }
  }
  {  // test-minmaxdepth.R:44-59 — update with own samples, then a0 <- 0 and max_depth <- 7: still 1000 ballots
    RDirichletTree t(letters(10), 3, 10, 1.0, false, "s");
    Rcpp::List s = t.samplePredictive(1000, "OWNSAMPLES");
    t.update(s);
    t.setA0(0.0);
    t.setMaxDepth(7);
    Rcpp::List s2 = t.samplePredictive(1000, "OWNSAMPLE2");
    CHECK(s2.size() == 1000);
    for (R_xlen_t i = 0; i < s2.size(); ++i) CHECK(Rcpp::CharacterVector(s2[i]).size() <= 7);
  }
  {  // test-minmaxdepth.R:61-75 — setters keep min_depth <= max_depth (R_tree.cpp:87-112)
    RDirichletTree lo(letters(26), 3, 26, 1.0, false, "s"), hi(letters(26), 0, 3, 1.0, false, "s");
    CHECK(throws([&] { lo.setMaxDepth(2); }));
    CHECK(throws([&] { hi.setMinDepth(4); }));
    CHECK(lo.getMaxDepth() == 26 && hi.getMinDepth() == 0 && lo.getNCandidates() == 26 && lo.getA0() == 1.0 && !lo.getVD());
    Rcpp::CharacterVector c = lo.getCandidates();
    CHECK(c.size() == 26 && c.at(0) == "A" && c.at(25) == "Z");  // constructor order (the reference's is unspecified)
    CHECK(throws([&] { RDirichletTree bad(letters(4), 4, 3, 1.0, false, "s"); }));
  }
  {  // test-determinism.R:5-42 — same seed string, same result (1 and 100 elections; predictive draws)
    RDirichletTree t(letters(10), 0, 10, 1.0, false, "s");
    t.update(ballots({{1, 2, 3}, {2, 1}, {5}}));
    for (unsigned n : {1u, 100u}) {
      auto a = probs(t.samplePosterior(n, 100, 1, false, 2, "DETERMINIS"));
      auto b = probs(t.samplePosterior(n, 100, 1, false, 2, "DETERMINIS"));
      CHECK(a == b);
    }
    Rcpp::List a = t.samplePredictive(1000, "DETERMINIS"), b = t.samplePredictive(1000, "DETERMINIS");
    CHECK(a.size() == b.size());
    for (R_xlen_t i = 0; i < a.size(); ++i) CHECK(*Rcpp::CharacterVector(a[i]).v.chr == *Rcpp::CharacterVector(b[i]).v.chr);
  }
  {  // RcppThread::checkUserInterrupt (R_tree.cpp:222): a pending interrupt aborts the call with an R error
    RDirichletTree t(letters(8), 0, 8, 1.0, false, "s");
    Rcpp::stub::interrupt_pending() = true;
    std::string msg;
    CHECK(throws([&] { t.samplePosterior(3000000, 50, 1, false, 2, "INTERRUPTA"); }, &msg));
    CHECK(msg.find("aborted") != std::string::npos);
    Rcpp::stub::interrupt_pending() = false;
    auto p = probs(t.samplePosterior(100, 50, 1, false, 2, "INTERRUPTA"));
    double s = 0;
    for (double x : p) s += x;
    CHECK(std::fabs(s - 1.0) < 1e-12);
  }
  {  // social_choice_irv: errors of R_social_choice.cpp:47-48,65-69 and a small known answer
    Rcpp::List bs = ballots({{1, 2}, {1, 2}, {2, 1}, {3, 2}, {3, 2}});  // C out first? tallies A2 B1 C2: B is eliminated
    Rcpp::List out = social_choice_irv(bs, 1, letters(3), "TIESEEDAAA");
    Rcpp::CharacterVector elim = out["elimination_order"], win = out["winners"];
    CHECK(elim.size() == 2 && win.size() == 1 && elim.at(0) == "B");
    CHECK(throws([&] { social_choice_irv(ballots({{4}}), 1, letters(3), "s"); }));
    CHECK(throws([&] { social_choice_irv(bs, 3, letters(3), "s"); }));
    Rcpp::List none;
    none.push_back_null();
    CHECK(throws([&] { social_choice_irv(none, 1, letters(3), "s"); }));
  }
  if (argc > 1) {  // test-socialchoice.R:11-23 — Wakehurst 2023: file of "count name name ...", expected order in argv[2..]
    std::ifstream in(argv[1]);
    Rcpp::List bs;
    Rcpp::CharacterVector cands;
    std::string line;
    std::getline(in, line);
    {
      std::istringstream ls(line);
      std::string nm;
      while (ls >> nm) cands.push_back(nm);
    }
    while (std::getline(in, line)) {
      std::istringstream ls(line);
      long count;
      ls >> count;
      Rcpp::CharacterVector b;
      std::string nm;
      while (ls >> nm) b.push_back(nm);
      for (long c = 0; c < count; ++c) bs.push_back(b);
    }
    Rcpp::List out = social_choice_irv(bs, 1, cands, "WAKEHURSTA");
    Rcpp::CharacterVector elim = out["elimination_order"], win = out["winners"];
    CHECK((int)elim.size() == argc - 3 && win.size() == 1);
    for (int i = 0; i + 3 < argc; ++i) CHECK(elim.at(i) == argv[2 + i]);
    CHECK(win.at(0) == argv[argc - 1]);
  }
  std::printf("shim ok: %d checks\n", n_checks);
  return 0;
}
