"""Host-side breakdown of the e2e call (dtree_invalidate_device + dtree_sample_posterior) on C3."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
import numpy as np
pkg = load_pkg()
W = pkg.workloads
from elections_dtree_b200 import _native as N
L = N.lib()
cfg = W.CONFIGS["C3"]
prefs, offsets = W.observed_ballots("C3")
t = pkg.dirichlet_tree(W.candidate_names(10), 0, 10, 1.0, False, device=0)
t.update_indices(prefs, offsets)
for inval in (1, 0):
    for rep in range(2):
        acc = {}
        n = 50
        t0 = time.perf_counter()
        for i in range(n):
            if inval:
                N.check(L.dtree_invalidate_device(t._h))
            t.sample_posterior_counts(100000, cfg["n_ballots"], seed="P", first_election=i * 100000)
            st = t.last_stats()
            for k in ("mirror_us", "plan_us", "launch_us", "finish_us"):
                acc[k] = acc.get(k, 0) + st[k]
        dt = (time.perf_counter() - t0) / n * 1e3
        print("invalidate=%d rep %d: %.3f ms/call" % (inval, rep, dt), {k: round(v / n, 1) for k, v in acc.items()}, flush=True)
