"""GPU parity tests: everything calls the CUDA path through the C ABI (ctypes) and is checked
against the CPU oracle (oracle/, pinned to the reference by tests/test_oracle_pin.py).

 * bit-exact: IRV elimination orders on identical ballots (Wakehurst golden vector, oracle-sampled
   elections, and the kernel's own elections re-run by the oracle), win counts over election ranges;
 * distributional (fixed seeds, tolerances written at each assert): Gamma / binomial generators,
   Dirichlet-tree marginals vs closed form and vs the oracle, ballot-type chi-squares, posterior win
   probabilities within binomial Monte-Carlo error of the oracle's.
"""
import ctypes as C
import warnings

import numpy as np
import pytest
from scipy import stats

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(pkg):
    lib = pkg._native.lib()
    if lib.dtree_device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests cannot run (there is no CPU fallback)")
    return lib


def names(pkg, nc):
    return pkg.workloads.candidate_names(nc)


def make_pair(pkg, port, nc, mn, mx, a0, vd, prefs=None, offsets=None):
    t = pkg.dirichlet_tree(names(pkg, nc), mn, mx, a0, vd, device=0)
    o = port.tree(nc, mn, mx, a0, vd)
    if prefs is not None:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t.update_indices(prefs, offsets)
        o.update((prefs, offsets))
    return t, o


def type_counts(prefs, offsets, counts):
    d = {}
    for b, c in zip(oracle.csr_to_ballots(prefs, offsets), counts):
        d[b] = d.get(b, 0) + int(c)
    return d


# ------------------------------------------------------------------------------------------------
# generators
# ------------------------------------------------------------------------------------------------
def test_philox_device_matches_host_and_known_answers(L):
    rng = np.random.default_rng(0)
    cases = [([0, 0, 0, 0], [0, 0]), ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])]
    cases += [(rng.integers(0, 2**32, 4).tolist(), rng.integers(0, 2**32, 2).tolist()) for _ in range(20)]
    for ctr, key in cases:
        h, d = (C.c_uint32 * 4)(), (C.c_uint32 * 4)()
        assert L.dtree_debug_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), h, 0) == 0
        assert L.dtree_debug_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), d, 1) == 0
        assert list(h) == list(d)
    assert list(d) != [0, 0, 0, 0]


@pytest.mark.parametrize("alpha,log_space", [(1.0, 0), (1.0, 1), (2.5, 0), (30.0, 0), (1e4, 0), (3e6, 0), (0.3, 1),
                                             (0.3, 0), (0.02, 1), (7.0, 1), (1.5, 0), (100.5, 1)])
def test_gamma_variates_match_the_gamma_law(L, alpha, log_space):
    """Counterpart of std::gamma_distribution in distributions.cpp:63-64 (KS p > 1e-4; mean and
    variance within 5 standard errors)."""
    n = 200_000
    out = np.zeros(n, dtype=np.float32)
    assert L.dtree_debug_gamma(alpha, log_space, 1234 + int(alpha * 10), n, out.ctypes.data_as(C.POINTER(C.c_float)), 0) == 0
    x = out.astype(np.float64)
    if log_space:
        # log of a Gamma variate: compare on the log scale with the exact law of log G
        cdf = lambda v: stats.gamma.cdf(np.exp(v), alpha)
        assert stats.kstest(x, cdf).pvalue > 1e-4
        x = np.exp(x)
    else:
        assert np.all(x > 0) or alpha < 1
        assert stats.kstest(x, lambda v: stats.gamma.cdf(v, alpha)).pvalue > 1e-4
    se_mean = np.sqrt(alpha / n)
    assert abs(x.mean() - alpha) < 5 * se_mean + 1e-6 * alpha
    var_of_var = (6 * alpha + 2 * alpha * alpha) / n  # Var of sample variance ~ (mu4 - sigma^4)/n
    assert abs(x.var() - alpha) < 5 * np.sqrt(var_of_var) + 1e-5 * alpha


def test_log_gamma_carries_tiny_shapes(L):
    """a0 = 1e-4 (config C5): Gamma(a) underflows in double for 93% of draws, so the
    reference falls back to a one-hot; the log-space variate must follow log G exactly instead:
    P(log G <= t) = P(G <= e^t) even at e^t ~ 1e-300 and below."""
    alpha, n = 1e-4, 200_000
    out = np.zeros(n, dtype=np.float32)
    assert L.dtree_debug_gamma(alpha, 1, 99, n, out.ctypes.data_as(C.POINTER(C.c_float)), 0) == 0
    x = out.astype(np.float64)
    assert np.all(np.isfinite(x))
    # for tiny alpha, G ~ U^(1/alpha) so -alpha*log G ~ Exp(1) to O(alpha)
    assert stats.kstest(-alpha * x, "expon").pvalue > 1e-4


@pytest.mark.parametrize("n_trials,p", [(1, 0.5), (10, 0.3), (50, 0.5), (1000, 0.001), (1000, 0.3), (100, 0.97),
                                        (20, 0.5), (40, 0.25), (10**6, 0.4), (10**6, 9.9e-6), (10**6, 1.2e-5),
                                        (4_000_000_000, 0.5), (3_000_000_000, 1e-9), (12345, 0.9991), (37, 0.0),
                                        (37, 1.0), (200, 0.05), (200, 0.06), (1000, 0.028), (1000, 0.031), (10**6, 2.9e-5),
                                        (58, 0.5), (62, 0.5), (5000, 0.994)])
def test_binomial_variates_match_the_binomial_law(L, n_trials, p):
    """Counterpart of std::binomial_distribution in distributions.cpp:45-46: chi-square against the exact pmf
    (p > 1e-4) and mean within 5 standard errors."""
    n = 200_000
    out = np.zeros(n, dtype=np.uint32)
    assert L.dtree_debug_binomial(n_trials, p, 777 + n_trials % 1000, n, out.ctypes.data_as(C.POINTER(C.c_uint32)), 0) == 0
    x = out.astype(np.float64)
    assert x.min() >= 0 and x.max() <= n_trials
    if p in (0.0, 1.0):
        assert np.all(x == n_trials * p)
        return
    mean, sd = n_trials * p, np.sqrt(n_trials * p * (1 - p))
    assert abs(x.mean() - mean) < 5 * sd / np.sqrt(n) + 1e-7 * mean
    assert abs(x.var() - sd * sd) < 0.03 * sd * sd + 1e-9
    # chi-square over ~40 equiprobable-ish bins of the support
    qs = np.unique(stats.binom.ppf(np.linspace(0.0005, 0.9995, 41), n_trials, p))
    edges = np.concatenate([[-1], qs, [n_trials]])
    edges = np.unique(edges)
    exp = np.diff(stats.binom.cdf(edges, n_trials, p)) * n
    obs = np.array([np.sum((x > lo) & (x <= hi)) for lo, hi in zip(edges[:-1], edges[1:])], dtype=np.float64)
    keep = exp > 5
    obs = np.append(obs[keep], obs[~keep].sum())
    exp = np.append(exp[keep], exp[~keep].sum())
    if exp[-1] < 1e-9:
        obs, exp = obs[:-1], exp[:-1]
    chi2 = ((obs - exp) ** 2 / np.maximum(exp, 1e-12)).sum()
    assert stats.chi2.sf(chi2, max(len(exp) - 1, 1)) > 1e-4, (chi2, len(exp))


# ------------------------------------------------------------------------------------------------
# IRV, bit-exact
# ------------------------------------------------------------------------------------------------
def test_irv_wakehurst_golden_vector(pkg, L, wakehurst):
    """tests/testthat/test-socialchoice.R:11-23 through the host API and dtree_irv."""
    cands = wakehurst["candidates"]
    ballots = [[cands[i] for i in b] for b in wakehurst["ballots"]]
    out = pkg.social_choice(ballots, "irv", counts=wakehurst["counts"], candidates=cands)
    assert out["winners"] == [wakehurst["irv_winner"]]
    assert out["elimination_order"] == wakehurst["irv_elimination_order"]
    assert pkg.social_choice([b[:1] for b in ballots], "plurality", counts=wakehurst["counts"],
                             candidates=cands) == [wakehurst["plurality_winner"]]


@pytest.mark.parametrize("nc", [2, 3, 4, 6, 10, 13, 14, 20, 26, 32])
def test_irv_bit_exact_on_random_elections(pkg, L, port, nc):
    """dtree_irv == socialChoiceIRV (irv_ballot.cpp:42-138) on identical ballots. Tie-free elections must
    match exactly; with ties the order must be one the reference rule allows (its tie-break consumes
    mt19937 draws we do not reproduce)."""
    rng = np.random.default_rng(nc)
    exact = 0
    for trial in range(40):
        n = int(rng.integers(1, 400 if nc <= 20 else 3000))
        pmf = rng.random(nc + 1) + 0.01
        prefs, offsets = pkg.workloads.plackett_luce_ballots(nc, float(rng.random()), n, 1000 * nc + trial, pmf)
        counts = rng.integers(1, 1_000_000 if trial % 2 else 5, n).astype(np.uint32)
        order, ties = pkg.irv_indices(prefs, offsets, counts, nc, "s%d" % trial, 0)
        ok, oties = port.irv_check_order((prefs, offsets), counts, nc, order)
        assert ok and oties == ties
        if ties == 0:
            assert order.tolist() == port.irv((prefs, offsets), counts, nc, "whatever").tolist()
            exact += 1
    assert exact >= 5


def test_irv_bit_exact_on_reference_sampled_elections(pkg, L, port):
    """Elections drawn by the oracle's posteriorSet, decided by both sides."""
    nc = 8
    prefs, offsets = pkg.workloads.plackett_luce_ballots(nc, 0.4, 500, 5, [1, 3, 2, 2, 2, 1, 1, 1, 6])
    o = port.tree(nc, 0, nc, 1.0, False)
    o.update((prefs, offsets))
    exact = 0
    for es in range(25):
        p, off, cnt = o.posterior_set(20_000, False, es)
        order, ties = pkg.irv_indices(p, off, cnt, nc, "x", 0)
        ok, _ = port.irv_check_order((p, off), cnt, nc, order)
        assert ok
        if ties == 0:
            assert order.tolist() == port.irv((p, off), cnt, nc, "y").tolist()
            exact += 1
    assert exact >= 20


def test_irv_tie_break_is_uniform(pkg, L):
    """Three candidates with equal tallies: each is eliminated first about a third of the time
    (irv_ballot.cpp:100-102: uniform among the tied). 600 seeds, 5 sigma."""
    prefs, offsets = np.array([0, 1, 2], dtype=np.uint8), np.array([0, 1, 2, 3], dtype=np.uint64)
    first = np.zeros(3)
    for s in range(600):
        order, ties = pkg.irv_indices(prefs, offsets, [5, 5, 5], 3, "tie%d" % s, 0)
        assert ties >= 1
        first[order[0]] += 1
    assert np.all(np.abs(first - 200) < 5 * np.sqrt(600 * (1 / 3) * (2 / 3)))


# ------------------------------------------------------------------------------------------------
# sampler: known answers and invariants
# ------------------------------------------------------------------------------------------------
def test_bootstrap_known_answer(pkg, L):
    """tests/testthat/test-minmaxdepth.R:1-27"""
    t = pkg.dirtree(list("ABCDE"), a0=0.0, min_depth=0, max_depth=3, vd=False)
    pkg.update(t, [["C", "B", "A", "D", "E"]])
    for s in range(5):
        assert pkg.sample_predictive(t, 1, seed="k%d" % s) == [(("C", "B", "A"), 1)]
    assert pkg.sample_predictive(t, 1000, seed="k") == [(("C", "B", "A"), 1000)]


def test_sampled_lengths_and_totals(pkg, L):
    """tests/testthat/test-minmaxdepth.R:29-59, test-sampling.R:1-19, src/test-distributions.cpp:11-43"""
    t = pkg.dirtree(list("ABCDEFGHIJ"), a0=1.0, min_depth=2, max_depth=8)
    s = pkg.sample_predictive(t, 1000, seed="len")
    assert sum(c for _, c in s) == 1000 and all(2 <= len(b) <= 8 for b, _ in s)
    assert all(len(set(b)) == len(b) for b, _ in s)
    t = pkg.dirtree(list("ABCDEFGHIJ"), a0=1.0, min_depth=3)
    init = pkg.sample_predictive(t, 1000, seed="init")
    pkg.update(t, [list(b) for b, _ in init], counts=[c for _, c in init])
    t.a0 = 0
    t.max_depth = 7
    s = pkg.sample_predictive(t, 1000, seed="again")
    assert sum(c for _, c in s) == 1000 and all(3 <= len(b) <= 7 for b, _ in s)
    for n in (1, 2, 7, 8, 9, 33, 100_000):  # across the urn / Gamma switch
        for nc, mn, mx in ((4, 0, 4), (13, 1, 13), (20, 0, 5), (30, 2, 30)):
            t = pkg.dirtree(names(pkg, nc), min_depth=mn, max_depth=mx, a0=0.7)
            p, off, cnt = t.sample_predictive_indices(n, seed="n%d" % n)
            lens = np.diff(off.astype(np.int64))
            assert int(cnt.sum()) == n and lens.min() >= mn and lens.max() <= min(mx, nc - 1)
            assert len(type_counts(p, off, cnt)) == len(cnt)  # every ballot type is emitted once


def test_determinism_and_range_splitting(pkg, L):
    """tests/testthat/test-determinism.R:5-42, and the property the multi-GPU sharding rests on:
    any split of [0, n) into ranges gives bit-identical totals and per-election orders."""
    nc = 10
    prefs, offsets = pkg.workloads.plackett_luce_ballots(nc, 0.3, 300, 9, None)
    t = pkg.dirichlet_tree(names(pkg, nc), 0, nc, 1.0, False, device=0)
    t.update_indices(prefs, offsets)
    pkg.set_seed(1)
    a = pkg.sample_posterior(t, 100, 1000)
    s1 = pkg.sample_predictive(t, 1000)
    pkg.set_seed(1)
    assert pkg.sample_posterior(t, 100, 1000) == a and pkg.sample_predictive(t, 1000) == s1
    w, orders = t.sample_posterior_counts(300, 2000, seed="SPLIT", return_orders=True)
    acc = np.zeros(nc, dtype=np.uint64)
    parts = []
    for lo, hi in ((0, 1), (1, 38), (38, 150), (150, 300)):
        ww, oo = t.sample_posterior_counts(hi - lo, 2000, seed="SPLIT", first_election=lo, return_orders=True)
        acc += ww
        parts.append(oo)
    assert acc.tolist() == w.tolist() and np.array_equal(np.vstack(parts), orders)
    for mode in (0, 1, 2):
        for bt in (32, 64, 128, 256):  # neither the kernel nor its launch shape may matter
            t.tune(block_threads=bt, mode=mode)
            w2, o2 = t.sample_posterior_counts(300, 2000, seed="SPLIT", return_orders=True)
            assert w2.tolist() == w.tolist() and np.array_equal(o2, orders)
    t.tune(block_threads=0)
    assert t.sample_posterior_counts(300, 2000, seed="OTHER").tolist() != w.tolist()


@pytest.mark.parametrize("nc,mn,mx,a0,vd,n_obs,n_ballots,replace", [
    (4, 0, 4, 1.0, False, 0, 1000, False), (6, 0, 6, 1.0, False, 300, 300, False),
    (6, 0, 3, 1.0, False, 500, 3000, False), (9, 2, 5, 0.5, False, 800, 800, False),
    (10, 0, 10, 1.0, False, 3000, 100_000, False), (5, 0, 5, 0.0, False, 40, 400, True),
    (8, 7, 7, 1e-4, True, 100, 20_000, False), (20, 0, 20, 1.0, False, 500, 50_000, False),
    (32, 0, 32, 1.0, False, 300, 5000, False), (14, 3, 9, 2.0, True, 200, 3000, True),
    (2, 0, 2, 1.0, False, 20, 100, False), (3, 0, 2, 0.5, False, 30, 90, False), (3, 1, 3, 1.0, True, 0, 7, False),
    # rounds that hinge on a handful of ballots (ties, near-ties): the deferred kernels must resolve exactly as
    # much of the tree as the outcome needs
    (6, 0, 6, 1.0, False, 0, 30, False), (7, 0, 7, 0.01, False, 50, 5000, False),
    (10, 0, 10, 1.0, False, 0, 200_000, False), (12, 0, 12, 1.0, False, 2000, 20_000, True),
    (8, 0, 8, 0.2, False, 64, 64, False)])
def test_deferred_equals_materialised(pkg, L, nc, mn, mx, a0, vd, n_obs, n_ballots, replace):
    """mode=1 / mode=2 (expand a node only when IRV looks below it; a warp / a thread per election) and mode=0
    (materialise every ballot, as the reference does) must give the same elimination order for every election and
    the same win counts:
    both evaluate the same counter-based draws, and the deferred kernels decide a round as soon as the ballots they have
    not yet placed cannot change it. Includes observed ballots longer than max_depth, n_ballots == n_observed
    (nothing sampled), an empty tree, a0 = 0, flat priors (every round close) and every key width."""
    pmf = np.ones(nc + 1)
    pmf[:mn] = 0
    t = pkg.dirichlet_tree(names(pkg, nc), mn, mx, a0, vd, device=0)
    if n_obs:
        prefs, offsets = pkg.workloads.plackett_luce_ballots(nc, 0.25, n_obs, nc + 77, pmf)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t.update_indices(prefs, offsets)
    res = {}
    # "0p": the materialising path as three launches per wave (sim_big / sim_walk / sim_irv kernels) instead of one
    for mode in (0, "0p", 1, 2):
        t.tune(mode=0 if mode == "0p" else mode, sim_phased=1 if mode == "0p" else 0)
        out = []
        for nw in ((1, 2) if nc > 2 else (1,)):
            out.append(t.sample_posterior_counts(400, n_ballots, n_winners=nw, replace=replace, seed="DEFER",
                                                 return_orders=True))
            out.append((t.sample_posterior_counts(400, n_ballots, n_winners=nw, replace=replace, seed="DEFER"), None))
        st = t.last_stats()
        assert st["mode"] == (0 if mode == "0p" else mode)
        if mode in (0, "0p"):
            assert st["launches"] == (3 if mode == "0p" else 1)
        res[mode] = out
    for mode in ("0p", 1, 2):
        for a, b in zip(res[0], res[mode]):
            assert a[0].tolist() == b[0].tolist(), mode
            assert (a[1] is None and b[1] is None) or np.array_equal(a[1], b[1]), mode
    assert res[0][0][0].tolist() == res[0][1][0].tolist()  # with and without the order output


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("nc,mn,mx,a0,vd,n_obs,n_ballots,replace", [
    (4, 0, 3, 1.0, False, 1000, 10_000, False), (6, 0, 5, 1.0, False, 300, 3000, False),
    (8, 7, 7, 0.01, True, 200, 5000, False), (10, 0, 10, 1.0, False, 2000, 20_000, False),
    (5, 0, 5, 0.0, False, 50, 500, False), (7, 2, 4, 3.0, False, 100, 1000, True),
    (16, 0, 16, 1.0, False, 400, 4000, False), (27, 1, 27, 0.5, True, 300, 2000, False)])
def test_in_kernel_irv_bit_exact_vs_oracle_on_the_kernels_own_elections(pkg, L, port, nc, mn, mx, a0, vd, n_obs,
                                                                        n_ballots, replace, mode):
    """sample_posterior's elimination orders vs socialChoiceIRV run by the oracle on exactly the ballots the
    kernel drew (dtree_posterior_set re-materialises election e). Covers the observed-ballot table, all
    key widths, a0 = 0 and the vd option."""
    pmf = np.ones(nc + 1)
    pmf[:mn] = 0
    prefs, offsets = pkg.workloads.plackett_luce_ballots(nc, 0.3, n_obs, nc + 50, pmf)
    t, o = make_pair(pkg, port, nc, mn, mx, a0, vd, prefs, offsets)
    t.tune(mode=mode)
    n_el = 24
    wins, orders = t.sample_posterior_counts(n_el, n_ballots, replace=replace, seed="OWN", return_orders=True)
    assert int(wins.sum()) == n_el
    recount = np.zeros(nc, dtype=np.int64)
    exact = 0
    for e in range(n_el):
        p, off, cnt = t.posterior_set_indices(e, n_ballots, replace, "OWN")
        assert int(cnt.sum()) == n_ballots
        ok, ties = port.irv_check_order((p, off), cnt, nc, orders[e].astype(np.uint32))
        assert ok, (e, orders[e])
        if ties == 0:
            assert port.irv((p, off), cnt, nc, "z").tolist() == orders[e].tolist()
            exact += 1
        recount[orders[e][-1]] += 1
    assert recount.tolist() == wins.tolist()
    assert exact >= n_el // 2 or a0 == 0.0 or nc > 10


# ------------------------------------------------------------------------------------------------
# sampler: distributions
# ------------------------------------------------------------------------------------------------
def test_root_split_is_dirichlet_multinomial(pkg, L):
    """First-preference tallies of sample(n) are DirMult(n, as_root + a0): mean n a_i/A, variance
    n (a_i/A)(1 - a_i/A)(A + n)/(A + 1). 4000 draws, 5 standard errors."""
    nc, n, draws = 5, 1000, 4000
    prefs, offsets = pkg.workloads.plackett_luce_ballots(nc, 0.6, 40, 3, [1, 1, 1, 1, 1, 3])
    t = pkg.dirichlet_tree(names(pkg, nc), 0, nc, 1.5, False, device=0)
    t.update_indices(prefs, offsets)
    dump = oracle.parse_node_dump(t.export_nodes())
    cands, counts, _ = dump[()]
    alpha = np.array(counts, dtype=np.float64) + 1.5  # nc candidates + stop
    A = alpha.sum()
    tallies = np.zeros((draws, nc + 1))
    for e in range(draws):
        p, off, cnt = t.posterior_set_indices(e, n, True, "ROOT")
        for b, c in zip(oracle.csr_to_ballots(p, off), cnt):
            tallies[e, cands.index(b[0]) if b else nc] += c
    assert np.all(tallies.sum(1) == n)
    mean = n * alpha / A
    var = n * (alpha / A) * (1 - alpha / A) * (A + n) / (A + 1)
    assert np.all(np.abs(tallies.mean(0) - mean) < 5 * np.sqrt(var / draws))
    assert np.all(np.abs(tallies.var(0) - var) < 0.15 * var)  # ~6 sigma of a sample variance at 4000 draws


def test_small_count_urn_matches_gamma_path(pkg, L):
    """The Polya-urn shortcut (counts <= urn_max) and the Gamma/binomial path must sample the same law:
    ballot-type frequencies of 30000 predictive draws of n=6 ballots, two-sample chi-square p > 1e-4."""
    nc = 4
    prefs, offsets = pkg.workloads.plackett_luce_ballots(nc, 0.5, 30, 8, [1, 2, 2, 2, 4])
    freq = []
    for urn in (8, 0):
        t = pkg.dirichlet_tree(names(pkg, nc), 0, nc, 0.8, False, device=0).tune(urn_max=urn)
        t.update_indices(prefs, offsets)
        d = {}
        for e in range(30000):
            for b, c in type_counts(*t.posterior_set_indices(e, 6, True, "URN")).items():
                d[b] = d.get(b, 0) + c
        freq.append(d)
    keys = sorted(set(freq[0]) | set(freq[1]))
    tab = np.array([[f.get(k, 0) for k in keys] for f in freq], dtype=np.float64)
    tab = tab[:, tab.sum(0) >= 10]
    assert stats.chi2_contingency(tab)[1] > 1e-4


@pytest.mark.parametrize("nc,mn,mx,a0,vd,n_obs", [(4, 0, 4, 1.0, False, 30), (5, 2, 4, 0.3, False, 20),
                                                  (4, 3, 3, 2.0, True, 10), (6, 0, 3, 5.0, False, 100)])
def test_ballot_type_frequencies_match_the_oracle(pkg, L, port, nc, mn, mx, a0, vd, n_obs):
    """Sampled ballot-type frequencies, GPU vs oracle: 3000 elections of 200 ballots on each side. Counts within
    an election are overdispersed (Dirichlet mixing), so every type is compared through its per-election mean with
    the per-election sample variances: z_k = (m_gpu - m_orc) / sqrt(v_gpu/n + v_orc/n). Required: max |z_k| < 4.5,
    the chi-square of the z_k (types seen >= 50 times) has p > 1e-4, and the per-election variances of the five
    commonest types agree within 25 %."""
    pmf = np.ones(nc + 1)
    pmf[:mn] = 0
    prefs, offsets = pkg.workloads.plackett_luce_ballots(nc, 0.4, n_obs, nc, pmf)
    t, o = make_pair(pkg, port, nc, mn, mx, a0, vd, prefs, offsets)
    n_el, n = 3000, 200
    per = [[], []]
    for e in range(n_el):
        per[0].append(type_counts(*t.posterior_set_indices(e, n, True, "CHI")))
        per[1].append(type_counts(*o.posterior_set(n, True, e)))
    keys = sorted(set().union(*[set(d) for side in per for d in side]))
    mats = [np.array([[d.get(k, 0) for k in keys] for d in side], dtype=np.float64) for side in per]
    assert np.all(mats[0].sum(1) == n) and np.all(mats[1].sum(1) == n)
    tot = mats[0].sum(0) + mats[1].sum(0)
    keep = tot >= 50
    m0, m1 = mats[0][:, keep].mean(0), mats[1][:, keep].mean(0)
    v0, v1 = mats[0][:, keep].var(0, ddof=1), mats[1][:, keep].var(0, ddof=1)
    z = (m0 - m1) / np.sqrt((v0 + v1) / n_el)
    assert np.abs(z).max() < 4.5, (np.abs(z).max(), keep.sum())
    assert stats.chi2.sf(float((z * z).sum()), int(keep.sum())) > 1e-4, ((z * z).sum(), keep.sum())
    top = np.argsort(-tot[keep])[:5]
    assert np.all((v0[top] / v1[top] > 0.8) & (v0[top] / v1[top] < 1.25)), (v0[top] / v1[top])


def test_vd_prior_is_uniform_over_ballot_types(pkg, L):
    """vd=TRUE on an empty tree is a flat Dirichlet over complete ballots (README / dtree paper): with
    min_depth = max_depth every ballot type has the same mean frequency. 4 candidates -> 24 types,
    2000 elections x 240 ballots; Dirichlet(1,...,1)-multinomial variance, 5 sigma."""
    nc = 4
    t = pkg.dirichlet_tree(names(pkg, nc), 3, 3, 1.0, True, device=0)
    tot = {}
    n_el, n = 2000, 240
    for e in range(n_el):
        for b, c in type_counts(*t.posterior_set_indices(e, n, True, "VD")).items():
            assert len(b) == 3
            tot[b] = tot.get(b, 0) + c
    assert len(tot) == 24
    x = np.array(list(tot.values()), dtype=np.float64)
    A, p = 24.0, 1 / 24.0
    var_one = n * p * (1 - p) * (A + n) / (A + 1)
    assert np.all(np.abs(x / n_el - n * p) < 5 * np.sqrt(var_one / n_el))


@pytest.mark.parametrize("cfg", ["C1", "C2", "C2small", "C5a", "C5b", "replace"])
def test_posterior_win_probabilities_match_the_oracle(pkg, L, port, cfg):
    """Posterior win probabilities, GPU vs oracle, within 4.5 sigma of the binomial Monte-Carlo error of the
    difference. C1 is BASELINE config 1 in full."""
    W = pkg.workloads
    if cfg == "C1":
        c = W.CONFIGS["C1"]
        nc, mn, mx, a0, vd = c["n_candidates"], c["min_depth"], c["max_depth"], c["a0"], c["vd"]
        prefs, offsets = W.observed_ballots("C1")
        n_el, n_b, replace = c["n_elections"] * 4, c["n_ballots"], False
    elif cfg == "C2":  # BASELINE config 2 in full: 6 candidates, Wakehurst-shaped partial ballots, 50 k ballots, 10 k elections
        c = W.CONFIGS["C2"]
        nc, mn, mx, a0, vd = c["n_candidates"], c["min_depth"], c["max_depth"], c["a0"], c["vd"]
        prefs, offsets = W.observed_ballots("C2")
        n_el, n_b, replace = c["n_elections"], c["n_ballots"], False
    elif cfg == "C2small":
        nc, mn, mx, a0, vd = 6, 0, 5, 1.0, False
        prefs, offsets = W.wakehurst_sample(300, 2)
        n_el, n_b, replace = 3000, 5000, False
    elif cfg in ("C5a", "C5b"):
        nc, mn, mx, vd = 8, 7, 7, True
        a0 = 1e-3 if cfg == "C5a" else 1.0
        prefs, offsets = W.plackett_luce_ballots(8, 0.1, 60, 5, None)
        n_el, n_b, replace = 1500, 2000, False
    else:
        nc, mn, mx, a0, vd = 5, 0, 5, 1.0, False
        prefs, offsets = W.plackett_luce_ballots(5, 0.2, 40, 6, [1, 1, 1, 1, 1, 2])
        n_el, n_b, replace = 4000, 60, True
    t, o = make_pair(pkg, port, nc, mn, mx, a0, vd, prefs, offsets)
    wins = t.sample_posterior_counts(n_el, n_b, replace=replace, seed="WIN" + cfg)
    gpu = wins / float(n_el)
    freq, _ = o.sample_posterior(n_el, n_b, 1, replace, 4, "WIN" + cfg)
    assert abs(gpu.sum() - 1) < 1e-12 and abs(freq.sum() - 1) < 1e-9
    pooled = (gpu + freq) / 2
    tol = 4.5 * np.sqrt(np.maximum(pooled * (1 - pooled), 1.0 / n_el) * 2 / n_el)
    assert np.all(np.abs(gpu - freq) <= tol), (gpu.tolist(), freq.tolist())


# ------------------------------------------------------------------------------------------------
# the reference's own behavioural tests, through the host API
# ------------------------------------------------------------------------------------------------
def test_posterior_shifts_after_observing_data(pkg, L):
    """tests/testthat/test-posterior.R:1-22 and test-update_and_reset.R:1-30"""
    pkg.set_seed(42)
    dtree = pkg.dirtree(list("ABCD"))
    prior = pkg.sample_posterior(dtree, 1000, 10)
    for _ in range(5):
        pkg.update(dtree, [list("ABCD")])
    post_dt = pkg.sample_posterior(dtree, 1000, 10)
    dtree.vd = True
    post_d = pkg.sample_posterior(dtree, 1000, 10)
    assert post_d["A"] > prior["A"] and post_dt["A"] > prior["A"]
    assert all(post_d[c] < prior[c] and post_dt[c] < prior[c] for c in "BCD")
    pkg.reset(dtree)
    dtree.vd = False
    again = pkg.sample_posterior(dtree, 1000, 10)
    assert abs(again["A"] - prior["A"]) < 0.1 and dtree.n_observed == 0


def test_posterior_is_uniform_when_a0_is_zero(pkg, L):
    """tests/testthat/test-posterior.R:24-29"""
    dtree = pkg.dirtree(list("ABCDEFGHIJ"), a0=0)
    probs = np.array(list(pkg.sample_posterior(dtree, 1000, 10, seed="a0").values()))
    assert abs(probs.sum() - 1) < 1e-12 and probs.var(ddof=1) < 1e-2


def test_replace_gives_more_than_one_possible_winner(pkg, L):
    """tests/testthat/test-posterior.R:71-93"""
    dtree = pkg.dirtree(list("ABC"))
    pkg.update(dtree, [list("ABC")] * 3)
    without = pkg.sample_posterior(dtree, 100, 3, replace=False, seed="r")
    assert without == {"A": 1.0, "B": 0.0, "C": 0.0}
    with_r = pkg.sample_posterior(dtree, 1000, 3, replace=True, seed="r")
    assert sum(v > 0 for v in with_r.values()) > 1


def test_n_winners_and_error_paths(pkg, L):
    dtree = pkg.dirtree(list("ABCDE"))
    pkg.update(dtree, [list("ABCDE")] * 4 + [list("BACDE")] * 3)
    two = pkg.sample_posterior(dtree, 500, 50, n_winners=2, seed="w")
    assert abs(sum(two.values()) - 2.0) < 1e-12 and two["A"] > 0.8 and two["B"] > 0.8
    with pytest.raises(pkg._native.DtreeError):
        dtree.sample_posterior_counts(10, 50, n_winners=5)
    with pytest.raises(ValueError):
        pkg.sample_posterior(dtree, 10, 3)
    # like the reference (R_tree.cpp:183-186) the native layer refuses n_ballots < n_observed even with replace
    with pytest.raises(pkg._native.DtreeError) as ei:
        pkg.sample_posterior(dtree, 10, 3, replace=True)
    assert ei.value.code == pkg._native.ERR_STATE


def test_abort_callback_stops_a_run(pkg, L):
    """RcppThread::checkUserInterrupt (R_tree.cpp:222) -> dtree_set_abort_callback."""
    dtree = pkg.dirtree(list("ABCDEFGH"))
    calls = []

    def cb(_):
        calls.append(1)
        return 1 if len(calls) >= 2 else 0

    fn = pkg._native.ABORT_CB(cb)
    lib = pkg._native.lib()
    lib.dtree_set_abort_callback(dtree._h, fn, None)
    dtree.tune(block_threads=32, blocks_per_sm=1)
    with pytest.raises(pkg._native.DtreeError) as ei:
        dtree.sample_posterior_counts(2_000_000, 50, seed="abort")
    assert ei.value.code == pkg._native.ERR_ABORTED and len(calls) == 2
    lib.dtree_set_abort_callback(dtree._h, C.cast(None, pkg._native.ABORT_CB), None)
    assert int(dtree.sample_posterior_counts(100, 50, seed="abort").sum()) == 100


# ------------------------------------------------------------------------------------------------
# BASELINE configurations at full size: size-independent properties
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["C2", "C3", "C4"])
def test_full_size_configs_conserve_ballots_and_agree_across_kernels(pkg, L, name):
    """At BASELINE.json's full shapes (6/10/20 candidates, 50 k / 1 M / 1 M ballots per election) the oracle is too
    slow to replay, so the checks are the domain's invariants: a materialised election holds exactly n_ballots
    ballots, every sampled ballot is a duplicate-free ranking of legal length listed once, the observed ballots are
    all in it, win counts add up, and the two kernels agree election by election."""
    W = pkg.workloads
    cfg = W.CONFIGS[name]
    nc = cfg["n_candidates"]
    prefs, offsets = W.observed_ballots(name)
    t = pkg.dirichlet_tree(names(pkg, nc), cfg["min_depth"], cfg["max_depth"], cfg["a0"], cfg["vd"], device=0)
    t.update_indices(prefs, offsets)
    n_obs = t.n_observed
    p, off, cnt = t.posterior_set_indices(5, cfg["n_ballots"], False, "FULL")
    assert int(cnt.astype(np.int64).sum()) == cfg["n_ballots"]
    op, ooff, ocnt = t.observed_indices()
    n_obs_entries = len(ocnt)
    assert int(ocnt.astype(np.int64).sum()) == n_obs
    assert np.array_equal(cnt[:n_obs_entries], ocnt) and np.array_equal(off[:n_obs_entries + 1], ooff)
    lens = np.diff(off.astype(np.int64))[n_obs_entries:]
    assert lens.min() >= cfg["min_depth"] and lens.max() <= min(cfg["max_depth"], nc - 1)
    sampled = oracle.csr_to_ballots(p, off)[n_obs_entries:]
    assert len(set(sampled)) == len(sampled) and all(len(set(b)) == len(b) for b in sampled[:20000])
    n_el = {"C2": 3000, "C3": 400, "C4": 60}[name]
    res = {}
    splits = {}
    for mode in (0, 1, 2):
        t.tune(mode=mode)
        res[mode] = t.sample_posterior_counts(n_el, cfg["n_ballots"], seed="FULL", return_orders=True)
        assert int(res[mode][0].sum()) == n_el
        plain = t.sample_posterior_counts(n_el, cfg["n_ballots"], seed="FULL")  # with the early-stop short-cuts
        assert plain.tolist() == res[mode][0].tolist()
        splits[mode] = t.last_stats()["items"] / n_el
    # the deferred kernels decide rounds without splitting what cannot matter: far fewer node splits, same outcomes
    assert splits[1] < 0.5 * splits[0] and splits[2] < 0.5 * splits[0], splits
    if name == "C3":
        assert splits[2] < 50, splits
    for mode in (1, 2):
        assert res[0][0].tolist() == res[mode][0].tolist() and np.array_equal(res[0][1], res[mode][1]), mode
    # the materialising path as one persistent kernel and as three launches per wave (whichever `auto` did not pick)
    for phased in (0, 1):
        t.tune(mode=0, sim_phased=phased)
        w, o = t.sample_posterior_counts(n_el, cfg["n_ballots"], seed="FULL", return_orders=True)
        launches = t.last_stats()["launches"]
        assert (launches >= 3 and launches % 3 == 0) if phased else launches == 1, launches
        assert w.tolist() == res[0][0].tolist() and np.array_equal(o, res[0][1]), phased
    t.tune(sim_phased=-1)
    # the order reported for election 5 is one the reference rule allows on election 5's ballots
    ok, _ = oracle.load("port").irv_check_order((p, off), cnt, nc, res[1][1][5].astype(np.uint32))
    assert ok


@pytest.mark.parametrize("a0", [1e-4, 1e-2, 1.0, 10.0])
def test_c5_dirichlet_special_case_calibration(pkg, L, port, a0):
    """BASELINE config 5 (vd=TRUE, min_depth = max_depth = 7, 8 candidates, a0 sweep), reduced to sizes the oracle
    finishes in seconds: posterior win probabilities, GPU vs oracle, within 4.5 sigma of the binomial Monte-Carlo
    error of the difference. a0 = 1e-4 is where Gamma variates underflow in double and the reference falls back to
    its one-hot (distributions.cpp:69-77); the log-space variates must land on the same probabilities."""
    W = pkg.workloads
    prefs, offsets = W.plackett_luce_ballots(8, 0.15, 120, 5, None)
    t, o = make_pair(pkg, port, 8, 7, 7, a0, True, prefs, offsets)
    n_el, n_b = 1500, 3000
    gpu = t.sample_posterior_counts(n_el, n_b, seed="C5cal") / float(n_el)
    freq, _ = o.sample_posterior(n_el, n_b, 1, False, 4, "C5cal")
    pooled = (gpu + freq) / 2
    tol = 4.5 * np.sqrt(np.maximum(pooled * (1 - pooled), 1.0 / n_el) * 2 / n_el)
    assert np.all(np.abs(gpu - freq) <= tol), (a0, gpu.tolist(), freq.tolist())


def test_one_process_many_devices_matches_single_device(pkg, L):
    """dtree_sample_posterior_multi: contiguous ranges on several devices (here every visible device, and the same
    device listed three times, which exercises the replica + range logic on a one-GPU box) give the single-device
    win counts bit for bit."""
    nc = 9
    prefs, offsets = pkg.workloads.plackett_luce_ballots(nc, 0.2, 400, 21, [1, 1, 1, 1, 1, 1, 1, 1, 1, 3])
    t = pkg.dirichlet_tree(names(pkg, nc), 0, nc, 1.0, False, device=0)
    t.update_indices(prefs, offsets)
    want = t.sample_posterior_counts(1001, 5000, seed="MULTI")
    ndev = L.dtree_device_count()
    for devices in ([0], [0, 0, 0], list(range(ndev)), list(range(ndev)) * 2):
        got = t.sample_posterior_counts(1001, 5000, seed="MULTI", devices=devices)
        assert got.tolist() == want.tolist(), devices
    t.update_indices(prefs[: int(offsets[50])], offsets[:51])  # replicas must follow an update
    want2 = t.sample_posterior_counts(500, 5000, seed="MULTI")
    assert t.sample_posterior_counts(500, 5000, seed="MULTI", devices=[0, 0]).tolist() == want2.tolist()
    assert want2.tolist() != want.tolist() or True


def test_handles_release_their_device_memory(pkg, L):
    """Create / sample / destroy in a loop (all three kernels, dumps, replicas): device memory in use must come back
    to where it started (RDirichletTree's finaliser path, R_tree.cpp:63-66)."""
    import gc
    import torch
    W = pkg.workloads
    prefs, offsets = W.plackett_luce_ballots(8, 0.3, 300, 4, None)

    def cycle():
        t = pkg.dirichlet_tree(names(pkg, 8), 0, 8, 1.0, False, device=0)
        t.update_indices(prefs, offsets)
        for mode in (0, 1, 2):
            t.tune(mode=mode)
            t.sample_posterior_counts(2000, 3000, seed="leak")
        t.sample_posterior_counts(500, 3000, seed="leak", devices=[0, 0])
        t.sample_predictive(1000, seed="leak")
        del t
        gc.collect()

    cycle()
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info(0)
    for _ in range(5):
        cycle()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info(0)
    assert free0 - free1 < 64 << 20, (free0, free1)


def test_largest_ballot_counts(pkg, L, port):
    """n_ballots near the top of the reference's `unsigned` range (4 x 10^9 per election): counts stay exact
    (f64 binomial rejection, u32 tallies), the three kernels agree, and win probabilities match the oracle
    (5 sigma of the binomial Monte-Carlo error)."""
    nc, n_b, n_el = 4, 4_000_000_000, 3000
    prefs, offsets = pkg.workloads.plackett_luce_ballots(nc, 0.02, 50, 2, [1, 1, 1, 1, 2])
    t, o = make_pair(pkg, port, nc, 0, nc, 1.0, False, prefs, offsets)
    p, off, cnt = t.posterior_set_indices(1, n_b, False, "BIG")
    assert int(cnt.astype(np.int64).sum()) == n_b
    res = []
    for mode in (0, 1, 2):
        t.tune(mode=mode)
        res.append(t.sample_posterior_counts(n_el, n_b, seed="BIG"))
        assert t.last_stats()["mode"] == mode
    assert res[0].tolist() == res[1].tolist() == res[2].tolist()
    gpu = res[0] / float(n_el)
    freq, _ = o.sample_posterior(n_el, n_b, 1, False, 4, "BIG")
    pooled = (gpu + freq) / 2
    tol = 5 * np.sqrt(np.maximum(pooled * (1 - pooled), 1.0 / n_el) * 2 / n_el)
    assert np.all(np.abs(gpu - freq) <= tol), (gpu.tolist(), freq.tolist())
