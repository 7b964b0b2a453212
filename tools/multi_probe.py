"""dtree_sample_posterior_multi (one process, every visible GPU) against the single-device call on C3: wall clock of
the host-buffer calls, tree resident on every device, third call of each shape; win counts must be identical."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
pkg = load_pkg()
W = pkg.workloads
L = pkg._native.lib()
nd = L.dtree_device_count()
cfg = W.CONFIGS["C3"]
prefs, offsets = W.observed_ballots("C3")
t = pkg.dirichlet_tree(W.candidate_names(10), 0, 10, 1.0, False, device=0)
t.update_indices(prefs, offsets)
per = 100_000
for devices in [None] + [list(range(k)) for k in (1, 2, 4, 8) if k <= nd]:
    n = per * (len(devices) if devices else 1)
    best = None
    for rep in range(4):
        t0 = time.perf_counter()
        w = t.sample_posterior_counts(n, cfg["n_ballots"], seed="MULTIPROBE", devices=devices)
        dt = time.perf_counter() - t0
        best = dt if best is None or rep > 0 and dt < best else best
    print("devices=%s: %d elections in %.3f ms = %.1f M elections/s, wins[0]=%d" % (devices, n, best * 1e3, n / best / 1e6, int(w[0])), flush=True)
