"""GPU tests of the paths an election takes when its scratch arena is too small.

In normal runs the arenas are sized from measured usage with a margin, so these paths are almost never taken:
 * deferred_kernel (mode 1): the election is re-run in one of a few shared worst-case arenas (try-lock);
 * lane_kernel (mode 2): the election's index goes to a retry list that a deferred_kernel launch enqueued right
   behind replays, a warp per election, with the worst-case arenas behind it; an election is also handed over when
   it needs more node splits than `lane_max_splits`.
The tuning key `arena_cap` forces arenas far too small; outcomes must stay bit-identical, election by election, to
the materialising kernel (mode 0), which has no arenas at all.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(pkg):
    lib = pkg._native.lib()
    if lib.dtree_device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests cannot run (there is no CPU fallback)")
    return lib


def build(pkg, name, gamma=None):
    W = pkg.workloads
    cfg = dict(W.CONFIGS[name])
    nc = cfg["n_candidates"]
    if gamma is None:
        prefs, offsets = W.observed_ballots(name)
    else:  # same shape, flatter candidate strengths: a close race
        rec = cfg["observed"]
        prefs, offsets = W.plackett_luce_ballots(nc, gamma, rec[2], rec[3])
    t = pkg.dirichlet_tree(W.candidate_names(nc), cfg["min_depth"], cfg["max_depth"], cfg["a0"], cfg["vd"], device=0)
    t.update_indices(prefs, offsets)
    return t, cfg


@pytest.mark.parametrize("name,gamma,n_el", [("C2", None, 3000), ("C3", 0.05, 160), ("C3", None, 4000)])
def test_overflowing_arenas_change_nothing(pkg, L, name, gamma, n_el):
    t, cfg = build(pkg, name, gamma)
    nb = cfg["n_ballots"]
    t.tune(mode=0)
    want_w, want_o = t.sample_posterior_counts(n_el, nb, seed="OVERFLOW", return_orders=True)
    for cap in (8, 48):
        t.tune(arena_cap=cap)
        # warp per election: own arena first, then a shared worst-case arena
        t.tune(mode=1)
        w, o = t.sample_posterior_counts(n_el, nb, seed="OVERFLOW", return_orders=True)
        st = t.last_stats()
        assert st["mode"] == 1 and w.tolist() == want_w.tolist() and np.array_equal(o, want_o), (cap, st)
        runs1 = st["arena_runs"]
        w = t.sample_posterior_counts(n_el, nb, seed="OVERFLOW")  # with the early-stop short cuts
        assert w.tolist() == want_w.tolist()
        # thread per election: retry list -> replay launch -> worst-case arenas
        t.tune(mode=2)
        w, o = t.sample_posterior_counts(n_el, nb, seed="OVERFLOW", return_orders=True)
        st = t.last_stats()
        assert st["mode"] == 2 and w.tolist() == want_w.tolist() and np.array_equal(o, want_o), (cap, st)
        retries2 = st["lane_retries"]
        w = t.sample_posterior_counts(n_el, nb, seed="OVERFLOW")
        assert w.tolist() == want_w.tolist()
        if cap == 8:  # nothing fits 8 records: every path was taken
            assert runs1 > 0 and retries2 > 0 and st["arena_runs"] > 0, (runs1, retries2, st)
    t.tune(arena_cap=0)


@pytest.mark.parametrize("name,gamma,n_el,budget", [("C3", None, 6000, 3), ("C3", 0.2, 2000, 40), ("C2", None, 5000, 1)])
def test_lane_split_budget_hands_elections_to_the_replay(pkg, L, name, gamma, n_el, budget):
    """lane_max_splits: a thread gives an election up after that many node splits; the replay launch (a warp per
    election in right-sized arenas) finishes it. Same outcomes as without a budget and as the materialising kernel."""
    t, cfg = build(pkg, name, gamma)
    nb = cfg["n_ballots"]
    t.tune(mode=0)
    want_w, want_o = t.sample_posterior_counts(min(n_el, 400), nb, seed="BUDGET", return_orders=True)
    t.tune(mode=2, lane_max_splits=0)
    free_w = t.sample_posterior_counts(n_el, nb, seed="BUDGET")
    assert t.last_stats()["lane_retries"] == 0
    t.tune(lane_max_splits=budget)
    w = t.sample_posterior_counts(n_el, nb, seed="BUDGET")
    st = t.last_stats()
    assert st["mode"] == 2 and 0 < st["lane_retries"] <= n_el, st
    assert w.tolist() == free_w.tolist()
    w, o = t.sample_posterior_counts(min(n_el, 400), nb, seed="BUDGET", return_orders=True)
    assert w.tolist() == want_w.tolist() and np.array_equal(o, want_o)
    t.tune(lane_max_splits=-1)


def test_multi_device_call_polls_the_abort_callback_and_collects(pkg, L):
    """dtree_sample_posterior_multi: the parent's abort callback is honoured on the multi-device path,
    an aborted call leaves nothing running (the next call on the same handle is correct), and a bad ordinal is
    refused before anything is launched."""
    nc = 8
    prefs, offsets = pkg.workloads.plackett_luce_ballots(nc, 0.2, 300, 5)
    t = pkg.dirichlet_tree(pkg.workloads.candidate_names(nc), 0, nc, 1.0, False, device=0)
    t.update_indices(prefs, offsets)
    want = t.sample_posterior_counts(3000, 4000, seed="MABORT")
    calls = []

    def cb(_):
        calls.append(1)
        return 1 if len(calls) >= 2 else 0

    fn = pkg._native.ABORT_CB(cb)
    L.dtree_set_abort_callback(t._h, fn, None)
    t.tune(block_threads=64, blocks_per_sm=1, mode=1)
    with pytest.raises(pkg._native.DtreeError) as ei:
        t.sample_posterior_counts(3_000_000, 4000, seed="MABORT", devices=[0, 0])
    assert ei.value.code == pkg._native.ERR_ABORTED and len(calls) == 2
    L.dtree_set_abort_callback(t._h, C.cast(None, pkg._native.ABORT_CB), None)
    t.tune(block_threads=0, blocks_per_sm=0, mode=-1)
    assert t.sample_posterior_counts(3000, 4000, seed="MABORT", devices=[0, 0]).tolist() == want.tolist()
    with pytest.raises(pkg._native.DtreeError) as ei:
        t.sample_posterior_counts(3000, 4000, seed="MABORT", devices=[0, L.dtree_device_count()])
    assert ei.value.code == pkg._native.ERR_ARG
    assert t.sample_posterior_counts(3000, 4000, seed="MABORT", devices=[0]).tolist() == want.tolist()


@pytest.mark.parametrize("nc,mn,mx,a0,vd,n_obs,n_ballots,replace", [
    (4, 0, 4, 1.0, False, 30, 400, False), (9, 2, 6, 0.3, False, 200, 200, False), (20, 0, 20, 1.0, True, 100, 5000, True),
    (6, 0, 6, 0.0, False, 12, 300, False), (32, 0, 32, 1.0, False, 64, 640, False), (5, 0, 5, 1.0, False, 0, 7, False)])
def test_plurality_on_the_device_is_bit_exact(pkg, L, nc, mn, mx, a0, vd, n_obs, n_ballots, replace):
    """dtree_sample_posterior_sc(DTREE_SC_PLURALITY) (R/social_choice.R:58-83): for every simulated
    election, every candidate with the maximal first-preference tally (blank ballots dropped) gets a win. Checked
    bit for bit against tallies of the very same elections materialised by dtree_posterior_set, including ties
    (small elections, a0 = 0) and elections with nothing sampled."""
    import warnings
    W = pkg.workloads
    pmf = np.ones(nc + 1)
    pmf[:mn] = 0
    t = pkg.dirichlet_tree(W.candidate_names(nc), mn, mx, a0, vd, device=0)
    if n_obs:
        prefs, offsets = W.plackett_luce_ballots(nc, 0.1, n_obs, nc + 5, pmf)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t.update_indices(prefs, offsets)
    n_el = 300
    got = t.sample_posterior_counts(n_el, n_ballots, replace=replace, seed="PLUR", sc_function="plurality")
    want = np.zeros(nc, dtype=np.int64)
    ties = 0
    for e in range(n_el):
        p, off, cnt = t.posterior_set_indices(e, n_ballots, replace, "PLUR")
        lens = np.diff(off.astype(np.int64))
        tally = np.zeros(nc, dtype=np.int64)
        has = lens > 0
        np.add.at(tally, p[off[:-1].astype(np.int64)[has]], cnt.astype(np.int64)[has])
        if tally.max() > 0:
            winners = np.flatnonzero(tally == tally.max())
            ties += len(winners) > 1
            want[winners] += 1
    assert got.astype(np.int64).tolist() == want.tolist(), (got.tolist(), want.tolist())
    assert int(got.sum()) >= n_el or n_ballots == 0
    if n_ballots <= 7:
        assert ties > 0  # the tie rule was exercised
    # range splitting and determinism hold as for IRV
    a = t.sample_posterior_counts(100, n_ballots, replace=replace, seed="PLUR", sc_function="plurality")
    b = t.sample_posterior_counts(200, n_ballots, replace=replace, seed="PLUR", sc_function="plurality", first_election=100)
    assert (a.astype(np.int64) + b.astype(np.int64)).tolist() == want.tolist()
    with pytest.raises(NotImplementedError):
        t.sample_posterior_counts(10, n_ballots, replace=replace, seed="PLUR", sc_function="stv")


def test_random_configurations_agree_across_kernels_and_tunings(pkg, L):
    """60 random small configurations (2..14 candidates, every combination of min/max depth, a0 from 0 to 5, vd,
    replace, empty to dense trees, 0..6000 ballots) x the three kernels x random arena caps, split budgets, urn limits
    and resolve thresholds: elimination orders and win counts must be identical to the materialising kernel's, and
    win counts must add up. A cheap net under everything the tunings can reach."""
    import warnings
    rng = np.random.default_rng(2024)
    W = pkg.workloads
    for case in range(60):
        nc = int(rng.integers(2, 15))
        mx = int(rng.integers(1, nc + 1))
        mn = int(rng.integers(0, mx + 1))
        a0 = float(rng.choice([0.0, 1e-3, 0.3, 1.0, 5.0]))
        vd = bool(rng.integers(0, 2)) and mx >= 1
        replace = bool(rng.integers(0, 2))
        n_obs = int(rng.choice([0, 1, 7, 60, 400]))
        t = pkg.dirichlet_tree(W.candidate_names(nc), mn, mx, a0, vd, device=0)
        if n_obs:
            pmf = rng.random(nc + 1) + 0.05
            prefs, offsets = W.plackett_luce_ballots(nc, float(rng.choice([0.0, 0.3, 1.0])), n_obs, 100 + case, pmf)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                t.update_indices(prefs, offsets)
        n_ballots = n_obs + int(rng.choice([0, 1, 9, 300, 6000]))
        if n_ballots == 0:
            n_ballots = 1
        nw = int(rng.integers(1, nc)) if nc > 2 else 1
        n_el = 160
        t.tune(mode=0, arena_cap=0, lane_max_splits=-1, urn_max=8, tau_quarters=3)
        want_w, want_o = t.sample_posterior_counts(n_el, n_ballots, n_winners=nw, replace=replace, seed="RND%d" % case,
                                                   return_orders=True)
        assert int(want_w.sum()) == n_el * nw
        for mode in (1, 2):
            t.tune(mode=mode, arena_cap=int(rng.choice([0, 8, 40])), lane_max_splits=int(rng.choice([-1, 0, 1, 6])),
                   urn_max=int(rng.choice([8, 8, 3, 0])), tau_quarters=int(rng.integers(1, 4)))
            # the urn limit changes which sampler a node uses, hence the draws: compare within the same urn_max
            um = int(rng.choice([8, 8, 8, 2]))
            t.tune(urn_max=um)
            if um != 8:
                t.tune(mode=0, arena_cap=0)
                ref_w, ref_o = t.sample_posterior_counts(n_el, n_ballots, n_winners=nw, replace=replace,
                                                         seed="RND%d" % case, return_orders=True)
                t.tune(mode=mode)
            else:
                ref_w, ref_o = want_w, want_o
            w, o = t.sample_posterior_counts(n_el, n_ballots, n_winners=nw, replace=replace, seed="RND%d" % case,
                                             return_orders=True)
            assert w.tolist() == ref_w.tolist() and np.array_equal(o, ref_o), (case, mode, nc, mn, mx, a0, vd, replace, n_obs, n_ballots)
            w2 = t.sample_posterior_counts(n_el, n_ballots, n_winners=nw, replace=replace, seed="RND%d" % case)
            assert w2.tolist() == ref_w.tolist(), (case, mode)
