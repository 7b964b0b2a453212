"""A tiny run through every kernel (all three simulation kernels, dump, IRV, variate hooks) for compute-sanitizer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
import numpy as np
pkg = load_pkg()
W = pkg.workloads
for nc, mn, mx, a0, vd in ((6, 0, 6, 1.0, False), (14, 1, 9, 0.5, True), (27, 0, 27, 1.0, False)):
    pmf = np.ones(nc + 1); pmf[:mn] = 0
    prefs, offsets = W.plackett_luce_ballots(nc, 0.2, 120, nc, pmf)
    t = pkg.dirichlet_tree(W.candidate_names(nc), mn, mx, a0, vd, device=0)
    t.update_indices(prefs, offsets)
    ref = None
    for mode in (0, 1, 2):
        t.tune(mode=mode)
        w, o = t.sample_posterior_counts(96, 1500, seed="SAN", return_orders=True)
        w2 = t.sample_posterior_counts(96, 1500, seed="SAN")
        assert w.tolist() == w2.tolist()
        if ref is None:
            ref = (w, o)
        assert ref[0].tolist() == w.tolist() and np.array_equal(ref[1], o)
    # the overflow paths: arenas of 8 records, so that elections re-run in the shared worst-case arenas (try-lock)
    # and lane_kernel hands them to its replay launch; a split budget of 1; the three tree-build paths; plurality
    t.tune(arena_cap=8)
    for mode in (1, 2):
        t.tune(mode=mode)
        w, o = t.sample_posterior_counts(96, 1500, seed="SAN", return_orders=True)
        assert ref[0].tolist() == w.tolist() and np.array_equal(ref[1], o)
    t.tune(arena_cap=0, mode=2, lane_max_splits=1)
    assert t.sample_posterior_counts(96, 1500, seed="SAN").tolist() == ref[0].tolist()
    t.tune(lane_max_splits=-1, mode=-1)
    for bm in (1, 2, 0):
        t.tune(build_mode=bm)
        assert t.sample_posterior_counts(96, 1500, seed="SAN").tolist() == ref[0].tolist()
    assert int(t.sample_posterior_counts(96, 1500, seed="SAN", sc_function="plurality").sum()) >= 96
    p, off, cnt = t.posterior_set_indices(3, 1500, False, "SAN")
    assert int(cnt.sum()) == 1500
    order, ties = pkg.irv_indices(p, off, cnt, nc, "s", 0)
    assert order.tolist() == ref[1][3].tolist() or ties > 0
    t.sample_predictive(500, seed="p")
print("sanitize case ok")
