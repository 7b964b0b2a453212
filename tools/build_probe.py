"""Device tree rebuild time on C3 without the H2D of the log (build_mode toggling re-mirrors from the device log)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
pkg = load_pkg()
W = pkg.workloads
from elections_dtree_b200 import _native as N
L = N.lib()
prefs, offsets = W.observed_ballots("C3")
t = pkg.dirichlet_tree(W.candidate_names(10), 0, 10, 1.0, False, device=0)
t.update_indices(prefs, offsets)
N.check(L.dtree_sync_device(t._h))
for mode in (0, 1, 0, 1):
    n = 50
    t0 = time.perf_counter()
    for i in range(n):
        t.tune(build_mode=mode)
        N.check(L.dtree_sync_device(t._h))
    dt = (time.perf_counter() - t0) / n * 1e3
    print("build_mode=%d: %.3f ms per rebuild (log already on the device), %d launches" % (mode, dt, t.last_stats()["build_launches"]), flush=True)
n = 50
t0 = time.perf_counter()
for i in range(n):
    N.check(L.dtree_invalidate_device(t._h))
    N.check(L.dtree_sync_device(t._h))
print("with H2D of the 3.6 MB log: %.3f ms" % ((time.perf_counter() - t0) / n * 1e3))
