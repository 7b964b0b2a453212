"""C3-shaped elections with flatter candidate strengths (closer races: later majorities, more of the tree expanded)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
pkg = load_pkg()
W = pkg.workloads
for gamma in (0.5, 0.2, 0.05, 0.0):
    prefs, offsets = W.plackett_luce_ballots(10, gamma, 200_000, 3)
    t = pkg.dirichlet_tree(W.candidate_names(10), 0, 10, 1.0, False, device=0)
    t.update_indices(prefs, offsets)
    for mode, n in ((1, 20000), (2, 100000), (0, 1200)):
        t.tune(mode=mode)
        t.sample_posterior_counts(max(n // 4, 300), 1_000_000, seed="w")
        t.sample_posterior_counts(max(n // 4, 300), 1_000_000, seed="w2")
        t0 = time.perf_counter()
        w = t.sample_posterior_counts(n, 1_000_000, seed="close")
        dt = time.perf_counter() - t0
        st = t.last_stats()
        print(json.dumps({"gamma": gamma, "mode": st["mode"], "elections": n, "elections_per_s": round(n / dt),
                          "nodes_split_per_election": round(st["items"] / n), "arena_runs": st["arena_runs"],
                          "scratch_GB": round(st["scratch_bytes"] / 1e9, 1), "p_win": (w / n).round(3).tolist()}), flush=True)
