"""Times the deferred kernel on C3 for several batch sizes (elections per launch), device-resident API."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
import numpy as np
import torch
pkg = load_pkg()
from elections_dtree_b200 import _native as N
W = pkg.workloads
cfg = W.CONFIGS["C3"]
prefs, offsets = W.observed_ballots("C3")
t = pkg.dirichlet_tree(W.candidate_names(10), 0, 10, 1.0, False, device=0)
t.update_indices(prefs, offsets)
t.tune(mode=int(os.environ.get("DTREE_MODE", "1")))
L = N.lib()
wins = torch.zeros(10, dtype=torch.int64, device="cuda")
stream = torch.cuda.current_stream()
seed = b"BENCHSEEDx"
first = 0
for epg in [int(x) for x in (sys.argv[1:] or ["2368", "4736", "9472", "12500", "25000", "4736"])]:
    times = []
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        N.check(L.dtree_sample_posterior_device(t._h, first, epg, cfg["n_ballots"], 1, 0, seed, len(seed), wins.data_ptr(), stream.cuda_stream))
        e1.record(stream)
        N.check(L.dtree_sample_posterior_finish(t._h, stream.cuda_stream))
        first += epg
        times.append(e0.elapsed_time(e1))
    st = t.last_stats()
    print(json.dumps({"epg": epg, "ms": [round(x, 1) for x in times], "el_per_s": round(epg / (min(times) / 1e3)),
                      "arena_runs": st["arena_runs"], "scratch_GB": round(st["scratch_bytes"] / 1e9, 2),
                      "items_per_el": round(st["items"] / st["elections"])}), flush=True)
