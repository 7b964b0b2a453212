"""Where do sporadic slow e2e steps come from? Times the pieces of a step separately, many repetitions."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
pkg = load_pkg()
W = pkg.workloads
L = pkg._native.lib()
cfg = W.CONFIGS["C3"]
prefs, offsets = W.observed_ballots("C3")
epg = 100000


def pct(v):
    v = np.array(v)
    return "median %.2f p90 %.2f p99 %.2f max %.2f ms" % (np.median(v), np.percentile(v, 90), np.percentile(v, 99), v.max())


for hb in (0, 1):
    t = pkg.dirichlet_tree(W.candidate_names(10), 0, 10, 1.0, False, device=0).tune(host_build=hb)
    t.update_indices(prefs, offsets)
    for i in range(3):
        t.sample_posterior_counts(epg, cfg["n_ballots"], seed="BENCHSEEDx", first_election=i * epg)
    a, b, b0 = [], [], []
    for i in range(60):
        ts = time.perf_counter()
        t.sample_posterior_counts(epg, cfg["n_ballots"], seed="BENCHSEEDx", first_election=(1 << 41) + i * epg)
        b0.append((time.perf_counter() - ts) * 1e3)
    for i in range(300):
        L.dtree_invalidate_device(t._h)
        ts = time.perf_counter(); L.dtree_sync_device(t._h); a.append((time.perf_counter() - ts) * 1e3)
    for i in range(60):
        ts = time.perf_counter()
        t.sample_posterior_counts(epg, cfg["n_ballots"], seed="BENCHSEEDx", first_election=(1 << 41) + i * epg)
        b.append((time.perf_counter() - ts) * 1e3)
    print("host_build", hb, "| mirror:", pct(a), "| sample before:", pct(b0), "| sample after:", pct(b), flush=True)
    print("   after, per step:", " ".join("%.0f" % x for x in b), flush=True)
    del t
