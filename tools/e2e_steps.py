"""Per-step wall time of invalidate + sample_posterior (the bench's e2e step) for both mirror builders."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
pkg = load_pkg()
W = pkg.workloads
L = pkg._native.lib()
cfg = W.CONFIGS["C3"]
prefs, offsets = W.observed_ballots("C3")
epg = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
for hb in (0, 1, 0):
    t = pkg.dirichlet_tree(W.candidate_names(10), 0, 10, 1.0, False, device=0).tune(host_build=hb)
    t.update_indices(prefs, offsets)
    for i in range(3):
        t.sample_posterior_counts(epg, cfg["n_ballots"], seed="BENCHSEEDx", first_election=i * epg)
    out = []
    for i in range(10):
        ts = time.perf_counter()
        L.dtree_invalidate_device(t._h)
        t.sample_posterior_counts(epg, cfg["n_ballots"], seed="BENCHSEEDx", first_election=(1 << 41) + i * epg)
        out.append("%.1f/%d" % ((time.perf_counter() - ts) * 1e3, t.last_stats()["scratch_bytes"] >> 20))
    print("host_build", hb, "ms/scratchMiB:", " ".join(out), flush=True)
    del t
