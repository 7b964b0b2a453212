"""Times sample_posterior for the BASELINE configurations over launch shapes (block size, blocks/SM, urn_max).
Usage: python tools/calibrate.py [C1 C2 ...]   — prints one JSON line per measurement."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg  # noqa: E402

pkg = load_pkg()
W = pkg.workloads
N_EL = {"C1": 200_000, "C2": 40_000, "C3": 1200, "C4": 300, "C5": 4000}
names = [x for x in sys.argv[1:] if x.startswith("C")] or ["C1", "C2", "C3", "C4", "C5"]
MODE = int(os.environ.get("DTREE_MODE", "1"))
if MODE >= 1:
    N_EL = {k: (v * 10 if k != "C4" else v) for k, v in N_EL.items()}
shapes = [(0, 0, 8)] + [(b, 0, 8) for b in (32, 64, 128, 256)] + [(0, 0, 0), (0, 0, 2), (0, 0, 4)]
for name in names:
    cfg = W.CONFIGS[name]
    prefs, offsets = W.observed_ballots(name)
    t = pkg.dirichlet_tree(W.candidate_names(cfg["n_candidates"]), cfg["min_depth"], cfg["max_depth"], cfg["a0"],
                           cfg["vd"], device=0)
    t0 = time.perf_counter()
    t.update_indices(prefs, offsets)
    t_update = time.perf_counter() - t0
    n = N_EL[name]
    for block, bps, urn in shapes:
        if MODE >= 1 and block == 32:
            continue
        t.tune(block_threads=block, blocks_per_sm=bps, urn_max=urn, mode=MODE)
        try:
            t.sample_posterior_counts(max(n // 20, 1), cfg["n_ballots"], seed="warm")
            t0 = time.perf_counter()
            w = t.sample_posterior_counts(n, cfg["n_ballots"], seed="calib")
            dt = time.perf_counter() - t0
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"config": name, "mode": MODE, "block": block, "urn_max": urn, "error": str(e)}), flush=True)
            continue
        st = t.last_stats()
        print(json.dumps({"config": name, "mode": MODE, "block": block, "blocks_per_sm": bps, "urn_max": urn, "elections": n,
                          "seconds": round(dt, 4), "elections_per_s": round(n / dt, 1), "update_s": round(t_update, 3),
                          "nodes": int(pkg._native.lib().dtree_n_nodes(t._h)),
                          "per_election": {k: round(st[k] / n, 1) for k in ("entries", "items", "gammas",
                                                                            "binomial_attempts", "irv_rereads",
                                                                            "slots", "urn_items")},
                          "scratch_GB": round(st["scratch_bytes"] / 1e9, 2), "wins": w.tolist()}), flush=True)
