"""Host-side overhead of one dtree_sample_posterior call on C3 (few elections, so the kernel is negligible)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
pkg = load_pkg()
from elections_dtree_b200 import _native as N
W = pkg.workloads
cfg = W.CONFIGS["C3"]
prefs, offsets = W.observed_ballots("C3")
t = pkg.dirichlet_tree(W.candidate_names(10), 0, 10, 1.0, False, device=0)
t.update_indices(prefs, offsets)
L = N.lib()
for mode in (1, 2):
    t.tune(mode=mode)
    t.sample_posterior_counts(100000, cfg["n_ballots"], seed="w")
    for label, inval, n in (("resident", False, 256), ("invalidate", True, 256), ("resident", False, 100000), ("invalidate", True, 100000)):
        ts = []
        for i in range(8):
            if inval:
                L.dtree_invalidate_device(t._h)
            t0 = time.perf_counter()
            t.sample_posterior_counts(n, cfg["n_ballots"], seed="x%d" % i)
            ts.append((time.perf_counter() - t0) * 1e3)
        print("mode", mode, label, "n", n, "ms per call:", " ".join("%.2f" % x for x in ts), flush=True)
