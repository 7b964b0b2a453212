"""Per-step time and library counters over many different election batches (tree resident)."""
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
epg = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
t = pkg.dirichlet_tree(W.candidate_names(10), 0, 10, 1.0, False, device=0)
if len(sys.argv) > 3:
    t.tune(mode=int(sys.argv[3]))
t.update_indices(prefs, offsets)
for i in range(3):
    t.sample_posterior_counts(epg, cfg["n_ballots"], seed="BENCHSEEDx", first_election=i * epg)
rows = []
for i in range(steps):
    ts = time.perf_counter()
    t.sample_posterior_counts(epg, cfg["n_ballots"], seed="BENCHSEEDx", first_election=(1 << 41) + i * epg)
    dt = (time.perf_counter() - ts) * 1e3
    s = t.last_stats()
    rows.append((dt, s["launches"], s["arena_runs"], s["scratch_bytes"] >> 20, s["mode"]))
for r in rows:
    print("%.1f ms launches %d arena_runs %d scratch %d MiB mode %d" % r)
