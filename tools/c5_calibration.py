"""BASELINE config 5: vd=TRUE Dirichlet special case, 8 candidates, a0 sweep 1e-4..10 — posterior win-probability
calibration of the GPU sampler against the reference (oracle/_ref when present, else the port).

    python tools/c5_calibration.py [n_elections_gpu] [n_elections_cpu] [n_ballots] > profiles/r01_c5_calibration.json

For each a0 the GPU simulates n_elections_gpu elections at the config's size; the reference simulates a bounded number
(it needs ~60 ms per 100k-ballot election and thread); z = (p_gpu - p_ref) / binomial standard error of the difference.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg  # noqa: E402
from oracle import oracle  # noqa: E402

pkg = load_pkg()
W = pkg.workloads
cfg = W.CONFIGS["C5"]
n_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["n_elections"]
n_cpu = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
n_ballots = int(sys.argv[3]) if len(sys.argv) > 3 else cfg["n_ballots"]
kind = "reference" if oracle.available("reference") else "port"
olib = oracle.load(kind)
threads = max(1, len(os.sched_getaffinity(0)))
prefs, offsets = W.observed_ballots("C5")
rows = []
for a0 in (1e-4, 1e-3, 1e-2, 0.1, 1.0, 10.0):
    t = pkg.dirichlet_tree(W.candidate_names(8), 7, 7, a0, True, device=0)
    t.update_indices(prefs, offsets)
    t.sample_posterior_counts(1000, n_ballots, seed="warm")
    t0 = time.perf_counter()
    gpu = t.sample_posterior_counts(n_gpu, n_ballots, seed="C5SWEEP") / float(n_gpu)
    gpu_s = time.perf_counter() - t0
    o = olib.tree(8, 7, 7, a0, True)
    o.update((prefs, offsets))
    ref, ref_s = o.sample_posterior(n_cpu, n_ballots, 1, False, threads, "C5SWEEP")
    o.close()
    se = np.sqrt(np.maximum(gpu * (1 - gpu), 1e-12) / n_gpu + np.maximum(ref * (1 - ref), 1.0 / n_cpu) / n_cpu)
    rows.append({"a0": a0, "gpu": gpu.round(5).tolist(), kind: ref.round(5).tolist(),
                 "max_abs_z": float(np.max(np.abs(gpu - ref) / se)), "gpu_elections_per_s": n_gpu / gpu_s,
                 kind + "_elections_per_s": n_cpu / ref_s})
print(json.dumps({"config": "C5: 8 candidates, vd=TRUE, min_depth=max_depth=7, 1000 observed PL(0.4) ballots, n_ballots=%d" % n_ballots,
                  "n_elections_gpu": n_gpu, "n_elections_" + kind: n_cpu, "threads": threads, "rows": rows}, indent=1))
