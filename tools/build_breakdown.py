"""Wall time of one device tree rebuild (C3 tree + 1000 new ballots per step), for profiling the build kernels."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
pkg = load_pkg()
W = pkg.workloads
L = pkg._native.lib()
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = W.CONFIGS[name]
nc = cfg["n_candidates"]
prefs, offsets = W.observed_ballots(name)
extra_p, extra_o = W.plackett_luce_ballots(nc, 0.5, 8000, 33)
t = pkg.dirichlet_tree(W.candidate_names(nc), 0, nc, 1.0, False, device=0)
t.update_indices(prefs, offsets)
t0 = time.perf_counter(); L.dtree_sync_device(t._h); print("first build %.3f ms" % ((time.perf_counter() - t0) * 1e3))
for k in range(8):
    lo, hi = int(extra_o[k * 1000]), int(extra_o[(k + 1) * 1000])
    t.update_indices(extra_p[lo:hi], (extra_o[k * 1000:(k + 1) * 1000 + 1] - lo).astype(np.uint64))
    t0 = time.perf_counter(); L.dtree_sync_device(t._h); dt = (time.perf_counter() - t0) * 1e3
    print("rebuild %d: %.3f ms, %d launches" % (k, dt, t.last_stats()["build_launches"]), flush=True)
L.dtree_invalidate_device(t._h)
t0 = time.perf_counter(); L.dtree_sync_device(t._h); print("rebuild from host memory %.3f ms" % ((time.perf_counter() - t0) * 1e3))
