"""Latency of observing ballots and getting the tree onto the GPU, C3 shape.

For each way of building the device mirror — on the device from the ballot log (default) or with the host
trie + upload (tuning host_build=1) — times: dtree_update of the 200 000 observed ballots, the first
dtree_sync_device after it, then an audit-style loop of `update(+1000 ballots); sample_posterior(1000
elections)`. The reference's own update (DirichletTree::update, unmodified sources in oracle/_ref) is timed
beside them on this host. Wall clock; every timed region ends in a call that synchronises the device.
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
pkg = load_pkg()
from oracle import oracle
N = pkg._native
W = pkg.workloads
cfg = W.CONFIGS["C3"]
nc = cfg["n_candidates"]
prefs, offsets = W.observed_ballots("C3")
extra_p, extra_o = W.plackett_luce_ballots(nc, 0.5, 20000, 33)
L = N.lib()
out = {"workload": "C3 observed ballots (200000 x 10 preferences)"}


def ms(f, *a):
    t0 = time.perf_counter()
    r = f(*a)
    return (time.perf_counter() - t0) * 1e3, r


for label, hb in (("device_build", 0), ("host_build", 1)):
    res = {"update_ms": [], "sync_ms": [], "audit_step_ms": [], "audit_update_ms": [], "audit_sample_ms": []}
    for rep in range(4):
        t = pkg.dirichlet_tree(W.candidate_names(nc), 0, nc, 1.0, False, device=0).tune(host_build=hb)
        t.sample_posterior_counts(8, 100, seed="warm")  # context, kernels loaded
        a, _ = ms(t.update_indices, prefs, offsets)
        b, _ = ms(L.dtree_sync_device, t._h)
        res["update_ms"].append(a)
        res["sync_ms"].append(b)
        t.sample_posterior_counts(1000, cfg["n_ballots"], seed="size")  # scratch sized
        for k in range(6):
            lo, hi = int(extra_o[k * 1000]), int(extra_o[(k + 1) * 1000])
            p = extra_p[lo:hi]
            o = (extra_o[k * 1000:(k + 1) * 1000 + 1] - lo).astype(np.uint64)
            u, _ = ms(t.update_indices, p, o)
            s, _ = ms(t.sample_posterior_counts, 1000, cfg["n_ballots"], 1, False, None, "audit%d" % k)
            if k >= 1:
                res["audit_update_ms"].append(u)
                res["audit_sample_ms"].append(s)
                res["audit_step_ms"].append(u + s)
        res["build_launches"] = t.last_stats()["build_launches"]
    out[label] = {k: (float(np.median(v)) if isinstance(v, list) else v) for k, v in res.items()}
    out[label]["first_build_total_ms"] = out[label]["update_ms"] + out[label]["sync_ms"]
    print(label, json.dumps(out[label]), flush=True)

if oracle.available("reference"):
    ref = oracle.load("reference")
    ts = []
    for rep in range(3):
        o = ref.tree(nc, 0, nc, 1.0, False)
        a, _ = ms(o.update, (prefs, offsets))
        ts.append(a)
        o.close()
    out["reference_update_ms"] = float(np.median(ts))
    print("reference update", out["reference_update_ms"], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "update_latency.json"), "w"), indent=1)
print(json.dumps(out))
