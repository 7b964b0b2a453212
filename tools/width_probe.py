"""Small jobs on C3: time per job by kernel and, for lane_kernel, by elections per warp (lane_width).
usage: width_probe.py [case]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
pkg = load_pkg()
W = pkg.workloads
case = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = W.CONFIGS[case]
prefs, offsets = W.observed_ballots(case)
t = pkg.dirichlet_tree(W.candidate_names(cfg["n_candidates"]), cfg["min_depth"], cfg["max_depth"], cfg["a0"], cfg["vd"], device=0)
t.update_indices(prefs, offsets)
def run(n, **tune):
    t.tune(**tune)
    best = None
    ref = None
    for rep in range(12):
        ts = time.perf_counter()
        w = t.sample_posterior_counts(n, cfg["n_ballots"], seed="wp", first_election=7)
        dt = (time.perf_counter() - ts) * 1e3
        if rep >= 2:
            best = dt if best is None else min(best, dt)
    return best, w.tolist(), t.last_stats()["mode"]
for n in (500, 1000, 2368, 5000, 9472, 12500, 25000, 50000, 100000):
    row = {"n": n}
    ms, ref, _ = run(n, mode=1, lane_width=0)
    row["deferred"] = round(ms, 3)
    for wd in (0, 32, 16, 8, 4, 2, 1):
        ms, w, m = run(n, mode=2, lane_width=wd)
        assert w == ref and m == 2
        row["lane_w%d" % wd] = round(ms, 3)
    ms, w, m = run(n, mode=-1, lane_width=0)
    assert w == ref
    row["auto"] = (round(ms, 3), m)
    print(json.dumps(row), flush=True)
