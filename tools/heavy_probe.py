"""Heavy-election regime (C4, close-race C3 shape, C5): elections/s and node splits per election of one kernel
for a list of tuning settings. usage: heavy_probe.py [mode] [key=v1,v2,...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
pkg = load_pkg()
W = pkg.workloads
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
key, vals = (sys.argv[2].split("=")[0], [int(v) for v in sys.argv[2].split("=")[1].split(",")]) if len(sys.argv) > 2 else ("tau_quarters", [2])
cases = sys.argv[3].split(",") if len(sys.argv) > 3 else ["C4", "C3g005", "C5", "C3g02"]
rows = []
for case in cases:
    name = case[:2]
    cfg = dict(W.CONFIGS[name])
    nc = cfg["n_candidates"]
    if "g" in case:
        gamma = {"g005": 0.05, "g02": 0.2, "g0": 0.0, "g01": 0.1, "g03": 0.3}[case[2:]]
        prefs, offsets = W.plackett_luce_ballots(nc, gamma, cfg["observed"][2], cfg["observed"][3])
    else:
        prefs, offsets = W.observed_ballots(name)
    t = pkg.dirichlet_tree(W.candidate_names(nc), cfg["min_depth"], cfg["max_depth"], cfg["a0"], cfg["vd"], device=0)
    t.update_indices(prefs, offsets)
    n = int(sys.argv[4]) if len(sys.argv) > 4 else {"C4": 2400, "C3": 20000, "C5": 50000, "C1": 1000000, "C2": 400000}[name]
    ref = None
    for v in vals:
        t.tune(mode=mode, **{key: v})
        best = None
        for rep in range(3):
            t0 = time.perf_counter()
            w = t.sample_posterior_counts(n, cfg["n_ballots"], seed="hp")
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        if ref is None:
            ref = w.tolist()
        assert w.tolist() == ref, "win counts changed with the tuning"
        st = t.last_stats()
        r = {"case": case, "mode": st["mode"], key: v, "elections": n, "ms": round(best * 1e3, 2),
             "elections_per_s": round(n / best), "splits_per_election": round(st["items"] / n, 1),
             "records_rw_per_election": round((st["entries"] + st["irv_rereads"]) / n, 1),
             "urn_splits_per_election": round(st["urn_items"] / n, 1), "gammas_per_election": round(st["gammas"] / n, 1),
             "retries": st["lane_retries"], "arena_runs": st["arena_runs"], "scratch_GB": round(st["scratch_bytes"] / 1e9, 2)}
        rows.append(r)
        print(json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "heavy_probe.json"), "w"), indent=1)
