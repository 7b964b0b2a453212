"""Three rebuilds of the C3 tree with one launch per phase (build_mode=1), for an ncu launch list of the phases."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
pkg = load_pkg()
W = pkg.workloads
from elections_dtree_b200 import _native as N
L = N.lib()
prefs, offsets = W.observed_ballots("C3")
t = pkg.dirichlet_tree(W.candidate_names(10), 0, 10, 1.0, False, device=0)
t.tune(build_mode=1)
t.update_indices(prefs, offsets)
for i in range(3):
    t.tune(build_mode=1)
    N.check(L.dtree_sync_device(t._h))
print("ok", t.last_stats()["build_launches"])
