import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_pkg
pkg = load_pkg(); W = pkg.workloads
import ctypes as C, numpy as np
from elections_dtree_b200 import _native as N
L = N.lib()
cfg = W.CONFIGS["C3"]; prefs, offsets = W.observed_ballots("C3")
t = pkg.dirichlet_tree(W.candidate_names(10), 0, 10, 1.0, False, device=0); t.update_indices(prefs, offsets)
wins = (C.c_uint64 * 10)()
seed = b"plur"
for n in (100000, 1000000):
    best = 1e9
    for r in range(6):
        ts = time.perf_counter()
        N.check(L.dtree_sample_posterior_sc(t._h, 1, 0, n, cfg["n_ballots"], 0, seed, len(seed), wins))
        best = min(best, time.perf_counter() - ts)
    print(n, "plurality", round(best * 1e3, 3), "ms", round(n / best / 1e6, 1), "M/s", list(wins)[:3])
