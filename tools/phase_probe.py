"""simulate_kernel by phase. Run twice: with the shipped library (true counters) and with the -DDTB_PHASE_CLOCKS
variant (tools/build_variant.sh clk -DDTB_PHASE_CLOCKS; DTREE_B200_LIB=...), whose gammas / binomial_attempts / slots /
irv_rereads counters hold cycles/64 summed over the warps' lane 0: small-item walk, big-item chunks, level-end barrier,
IRV stage. usage: phase_probe.py [case] [elections] [mode]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
pkg = load_pkg()
W = pkg.workloads
case = sys.argv[1] if len(sys.argv) > 1 else "C3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cfg = dict(W.CONFIGS[case])
nc = cfg["n_candidates"]
prefs, offsets = W.observed_ballots(case)
t = pkg.dirichlet_tree(W.candidate_names(nc), cfg["min_depth"], cfg["max_depth"], cfg["a0"], cfg["vd"], device=0)
t.update_indices(prefs, offsets)
t.tune(mode=mode)
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    t.tune(**{k: int(v)})
best = None
for rep in range(3):
    t0 = time.perf_counter()
    w = t.sample_posterior_counts(n, cfg["n_ballots"], seed="pp")
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
st = t.last_stats()
clk = "clk" in os.environ.get("DTREE_B200_LIB", "")
out = {"launches": st["launches"], "scratch_GB": round(st["scratch_bytes"] / 1e9, 2), "case": case, "elections": n, "ms": round(best * 1e3, 2), "elections_per_s": round(n / best), "wins": w.tolist()}
if clk:
    warps = 296 * 8
    tot = best * 1.965e9  # cycles of the whole call at 1965 MHz
    for k, name in (("gammas", "small_walk"), ("binomial_attempts", "big_chunks"), ("slots", "level_barrier"), ("irv_rereads", "irv")):
        out[name + "_share"] = round(st[k] * 64.0 / warps / tot, 3)
    out["frames_1_per_election"] = round(st["entries"] / n, 1)
    out["frames_2_3_per_election"] = round(st["items"] / n, 1)
    out["walker_passes_per_election"] = round(st["urn_items"] / n, 1)
else:
    for k in ("entries", "items", "urn_items", "gammas", "binomial_attempts", "irv_rereads", "slots"):
        out[k + "_per_election"] = round(st[k] / n, 1)
print(json.dumps(out), flush=True)
