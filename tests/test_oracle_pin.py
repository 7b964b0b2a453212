"""Pins the CPU oracle (oracle/dtree_oracle.cpp, the "port") before anything trusts it:

* against the reference's golden vector and known answers (Wakehurst IRV, a0=0 bootstrap);
* against tests/golden/ref_vectors.json, outputs of the UNMODIFIED reference core recorded by
  tests/golden/make_golden.py (travels to boxes without /root/reference);
* live, bit-for-bit, against oracle/_ref/libdtree_ref.so wherever that library exists.
"""
import numpy as np
import pytest

from oracle import oracle


def _take(x):
    prefs, offsets, counts = x
    return [list(b) for b in oracle.csr_to_ballots(prefs, offsets)], counts.tolist()


def _build(lib, case):
    t = lib.tree(case["n_candidates"], case["min_depth"], case["max_depth"], case["a0"], case["vd"], "GOLDENSEED")
    t.update(case["observed"])
    return t


def test_packaged_wakehurst_data_is_the_golden_fixture(pkg, wakehurst):
    assert pkg.workloads.wakehurst() == wakehurst


def test_wakehurst_golden(port, wakehurst):
    """reference tests/testthat/test-socialchoice.R:11-23"""
    names = wakehurst["candidates"]
    order = port.irv(wakehurst["ballots"], wakehurst["counts"], 6, "anyseed")
    assert [names[i] for i in order[:5]] == wakehurst["irv_elimination_order"]
    assert names[order[5]] == wakehurst["irv_winner"]
    ok, ties = port.irv_check_order(wakehurst["ballots"], wakehurst["counts"], 6, order)
    assert ok and ties == 0


def test_bootstrap_known_answer(port):
    """reference tests/testthat/test-minmaxdepth.R:1-27: a0=0, max_depth=3, one observed 5-ballot."""
    t = port.tree(5, 0, 3, 0.0, False)
    t.update([[2, 1, 0, 3, 4]])
    ballots, counts = _take(t.sample_predictive(1, "x"))
    assert ballots == [[2, 1, 0]] and counts == [1]


def test_depth_factors_probe(port):
    """SURVEY §8 a1 probes of irv_node.cpp:13-25, and the stale-factor quirk of irv_node.h:127."""
    assert port.tree(4, 0, 3).depth_factors().tolist() == [12, 3, 1]
    assert port.tree(4, 2, 3).depth_factors().tolist() == [9, 3, 1]
    assert port.tree(4, 3, 3).depth_factors().tolist() == [6, 2, 1]
    assert port.tree(4, 0, 4).depth_factors().tolist() == [24, 6, 2, 1]
    t = port.tree(4, 0, 3)
    t.set(min_depth=3)
    assert t.depth_factors().tolist() == [12, 3, 1]


def test_port_matches_recorded_reference(port, ref_vectors):
    for case in ref_vectors:
        t = _build(port, case)
        assert t.depth_factors().tolist() == case["depth_factors"]
        assert t.dump_nodes().tolist() == case["node_dump"]
        b, c = _take(t.observed())
        assert b == case["observed_map"]["ballots"] and c == case["observed_map"]["counts"]
        g = case["posterior_set"]
        ps = t.posterior_set(g["n_ballots"], False, g["engine_seed"])
        b, c = _take(ps)
        assert b == g["ballots"] and c == g["counts"]
        assert sum(c) == g["n_ballots"]
        g = case["posterior_set_replace"]
        b, c = _take(t.posterior_set(g["n_ballots"], True, g["engine_seed"]))
        assert b == g["ballots"] and c == g["counts"]
        assert port.irv((ps[0], ps[1]), ps[2], case["n_candidates"], "IRVSEED").tolist() == case["irv_of_posterior_set"]
        g = case["sample_predictive"]
        b, c = _take(t.sample_predictive(g["n_ballots"], g["seed"]))
        assert b == g["ballots"] and c == g["counts"]
        g = case["sample_posterior"]
        freq, _ = t.sample_posterior(g["n_elections"], g["n_ballots"], 1, False, g["n_threads"], g["seed"])
        assert freq.tolist() == g["freq"]
        t.close()


@pytest.mark.parametrize("nc,mn,mx,a0,vd", [(4, 0, 3, 1.0, False), (6, 0, 6, 1.0, False), (8, 7, 7, 1e-3, True),
                                            (10, 2, 8, 1.0, False), (5, 0, 5, 0.0, False), (12, 0, 4, 0.3, True)])
def test_port_matches_live_reference(port, ref, pkg, nc, mn, mx, a0, vd):
    """Bit-for-bit on fresh random inputs: tree, posteriorSet, IRV, samplePosterior (1..3 threads)."""
    rng = np.random.default_rng(nc * 1000 + mn)
    pmf = rng.random(nc + 1) + 0.05
    pmf[:mn] = 0 if vd else pmf[:mn] * 0.1
    prefs, offsets = pkg.workloads.plackett_luce_ballots(nc, 0.4, 500, nc + 7, pmf)
    tp, tr = port.tree(nc, mn, mx, a0, vd, "S1"), ref.tree(nc, mn, mx, a0, vd, "S1")
    for t in (tp, tr):
        t.update((prefs, offsets))
    assert tp.dump_nodes().tolist() == tr.dump_nodes().tolist()
    assert tp.n_observed() == tr.n_observed() == 500
    for replace, n in ((False, 900), (True, 300), (False, 500)):
        for es in (1, 99):
            a, b = tp.posterior_set(n, replace, es), tr.posterior_set(n, replace, es)
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
            assert port.irv((a[0], a[1]), a[2], nc, "T").tolist() == ref.irv((b[0], b[1]), b[2], nc, "T").tolist()
    a, b = tp.sample_predictive(1000, "PRED"), tr.sample_predictive(1000, "PRED")
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    for threads, nel in ((1, 7), (2, 25), (3, 10), (2, 1)):
        fa, _ = tp.sample_posterior(nel, 900, 1, False, threads, "POST")
        fb, _ = tr.sample_posterior(nel, 900, 1, False, threads, "POST")
        assert fa.tolist() == fb.tolist()
    # parameter changes take effect on the next sample without touching counts (irv_node.h:127-153)
    for t in (tp, tr):
        t.set(a0=a0 * 2 + 0.1, max_depth=max(mn, mx - 1), vd=not vd if mx > 1 else vd)
    a, b = tp.posterior_set(800, False, 5), tr.posterior_set(800, False, 5)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    for t in (tp, tr):
        t.reset()
        t.update((prefs[: int(offsets[10])], offsets[:11]))
    assert tp.dump_nodes().tolist() == tr.dump_nodes().tolist()


def test_invariants_sum_and_depth_bounds(port):
    """reference src/test-distributions.cpp:11-43 (sum to count) and test-minmaxdepth.R:29-42 (lengths)."""
    t = port.tree(10, 2, 8, 1.0, False)
    prefs, offsets, counts = t.sample_predictive(1000, "abc")
    lens = np.diff(offsets.astype(np.int64))
    assert counts.sum() == 1000 and lens.min() >= 2 and lens.max() <= 8


def test_irv_check_order_rejects_wrong_order(port, wakehurst):
    order = port.irv(wakehurst["ballots"], wakehurst["counts"], 6, "s").tolist()
    bad = order[:]
    bad[0], bad[1] = bad[1], bad[0]
    ok, _ = port.irv_check_order(wakehurst["ballots"], wakehurst["counts"], 6, bad)
    assert not ok
