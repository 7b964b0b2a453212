"""Elections per second of the warp-per-election (1) and thread-per-election (2) kernels on every benchmark
configuration, on close races (flatter candidate strengths) and for small batches. Wall clock around the
host-buffer call, tree resident, third call of each shape."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
pkg = load_pkg()
W = pkg.workloads
rows = []


def run(label, t, n_ballots, batches, modes=(1, 2)):
    for mode in modes:
        for n in batches[mode]:
            t.tune(mode=mode)
            best = None
            for rep in range(3):
                t0 = time.perf_counter()
                w = t.sample_posterior_counts(n, n_ballots, seed="mm%d" % rep)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            st = t.last_stats()
            r = {"case": label, "mode": st["mode"], "elections": n, "ms": round(best * 1e3, 2),
                 "elections_per_s": round(n / best), "nodes_split_per_election": round(st["items"] / n, 1),
                 "arena_runs": st["arena_runs"], "scratch_GB": round(st["scratch_bytes"] / 1e9, 2)}
            rows.append(r)
            print(json.dumps(r), flush=True)


which = sys.argv[1:] or ["C1", "C2", "C3", "C4", "C5", "close", "small"]
for name in [c for c in which if c in W.CONFIGS]:
    cfg = W.CONFIGS[name]
    nc = cfg["n_candidates"]
    prefs, offsets = W.observed_ballots(name)
    t = pkg.dirichlet_tree(W.candidate_names(nc), cfg["min_depth"], cfg["max_depth"], cfg["a0"], cfg["vd"], device=0)
    t.update_indices(prefs, offsets)
    big = {"C1": 1000000, "C2": 400000, "C3": 100000, "C4": 2000, "C5": 100000}[name]
    warp = {"C1": 200000, "C2": 100000, "C3": 20000, "C4": 2000, "C5": 50000}[name]
    run(name, t, cfg["n_ballots"], {1: [warp], 2: [big]})
if "close" in which:
    for gamma in (0.2, 0.05, 0.0):
        prefs, offsets = W.plackett_luce_ballots(10, gamma, 200_000, 3)
        t = pkg.dirichlet_tree(W.candidate_names(10), 0, 10, 1.0, False, device=0)
        t.update_indices(prefs, offsets)
        run("C3 shape, gamma=%g" % gamma, t, 1_000_000, {1: [20000], 2: [100000]})
if "small" in which:
    cfg = W.CONFIGS["C3"]
    prefs, offsets = W.observed_ballots("C3")
    t = pkg.dirichlet_tree(W.candidate_names(10), 0, 10, 1.0, False, device=0)
    t.update_indices(prefs, offsets)
    run("C3 small batches", t, cfg["n_ballots"], {1: [1, 32, 1000, 5000], 2: [1, 32, 1000, 5000]})
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "mode_matrix.json"), "w"), indent=1)
