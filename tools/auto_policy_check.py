"""Which kernel the automatic choice picks, and how fast it is, on C3-shaped elections with closer races."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
pkg = load_pkg()
W = pkg.workloads
rows = []
for gamma, n in ((0.5, 100000), (0.2, 100000), (0.05, 20000), (0.0, 20000)):
    prefs, offsets = W.plackett_luce_ballots(10, gamma, 200_000, 3)
    t = pkg.dirichlet_tree(W.candidate_names(10), 0, 10, 1.0, False, device=0)
    t.update_indices(prefs, offsets)
    best = None
    for rep in range(3):
        t0 = time.perf_counter()
        w = t.sample_posterior_counts(n, 1_000_000, seed="auto%d" % rep)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    st = t.last_stats()
    r = {"gamma": gamma, "mode_chosen": st["mode"], "elections": n, "ms": round(best * 1e3, 2), "elections_per_s": round(n / best),
         "nodes_split_per_election": round(st["items"] / n, 1), "arena_runs": st["arena_runs"],
         "scratch_GB": round(st["scratch_bytes"] / 1e9, 2), "p_win": (w / n).round(3).tolist()}
    rows.append(r)
    print(json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "auto_policy.json"), "w"), indent=1)
