"""Regenerates the committed fixtures under tests/golden/. Run HERE (needs /root/reference):

    python tests/golden/make_golden.py

1. wakehurst2023.json — the reference's only data fixture (tests/data/wakehurst2023.soi, PrefLib
   00058-00000273) as 0-based index ballots + counts, with the answer pinned by the reference's
   tests/testthat/test-socialchoice.R:5-23 (checked below against the compiled reference).
2. ref_vectors.json — outputs of the UNMODIFIED reference core (oracle/_ref/libdtree_ref.so) on
   small seeded inputs: node dumps, posteriorSet lists, IRV orders, sample_posterior frequencies,
   depth factors. tests/test_oracle_pin.py replays them against the port (and the CUDA tree build),
   so the pin travels to boxes where /root/reference does not exist.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SOI = "/root/reference/tests/data/wakehurst2023.soi"


def make_wakehurst(ref):
    names, ballots, counts = [], [], []
    for line in open(SOI):
        line = line.strip()
        if line.startswith("# ALTERNATIVE NAME"):
            names.append(line.split(":", 1)[1].strip())
        elif line and not line.startswith("#"):
            c, b = line.split(":")
            counts.append(int(c))
            ballots.append([int(x) - 1 for x in b.strip().split(",")])
    assert len(names) == 6 and sum(counts) == 51334 and len(ballots) == 1189
    expected_elim = ["MAWSON Greg", "SORENSEN Susan", "HRNJAK Ethan", "WRIGHT Sue", "WILLIAMS Toby"]
    expected_win = "REGAN Michael"
    order = ref.irv(ballots, counts, 6, "seed")
    assert [names[i] for i in order[:5]] == expected_elim and names[order[5]] == expected_win
    fp = np.zeros(6, dtype=np.int64)
    for b, c in zip(ballots, counts):
        fp[b[0]] += c
    assert fp.tolist() == [4000, 1127, 18430, 1220, 18940, 7617]  # README.md:150-156
    json.dump(dict(source="reference tests/data/wakehurst2023.soi (PrefLib 00058-00000273)",
                   pinned_by="reference tests/testthat/test-socialchoice.R:5-23",
                   candidates=names, ballots=ballots, counts=counts,
                   irv_elimination_order=expected_elim, irv_winner=expected_win,
                   plurality_winner="WILLIAMS Toby", first_preference_tallies=fp.tolist()),
              open(os.path.join(HERE, "wakehurst2023.json"), "w"), separators=(",", ":"))


def take(x):
    prefs, offsets, counts = x
    return dict(ballots=[list(b) for b in oracle.csr_to_ballots(prefs, offsets)], counts=counts.tolist())


def make_ref_vectors(ref):
    import importlib.util
    spec = importlib.util.spec_from_file_location("workloads", os.path.join(ROOT, "elections.dtree_b200", "workloads.py"))
    wl = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(wl)
    cases = []
    specs = [
        # nc, min, max, a0, vd, n_obs, gamma, length pmf, ballot seed
        (4, 0, 3, 1.0, False, 60, 0.3, None, 11),
        (4, 2, 4, 0.5, True, 40, 0.2, [0, 0, 1, 1, 2], 12),
        (5, 0, 3, 0.0, False, 1, 0.0, None, 13),            # the a0=0 bootstrap known answer
        (6, 0, 5, 1.0, False, 200, 0.4, [1, 6, 1, 1, 1, 1, 4], 14),
        (7, 1, 7, 2.0, False, 150, 0.5, [0, 2, 1, 1, 1, 1, 1, 5], 15),
        (3, 0, 2, 1.0, False, 30, 0.1, [1, 1, 1, 1], 16),
        (2, 0, 1, 1.0, False, 10, 0.3, [1, 2, 2], 17),
        (9, 3, 6, 0.01, True, 100, 0.3, None, 18),
    ]
    for nc, mn, mx, a0, vd, nobs, gamma, pmf, bseed in specs:
        prefs, offsets = wl.plackett_luce_ballots(nc, gamma, nobs, bseed, pmf)
        t = ref.tree(nc, mn, mx, a0, vd, "GOLDENSEED")
        t.update((prefs, offsets))
        case = dict(n_candidates=nc, min_depth=mn, max_depth=mx, a0=a0, vd=vd,
                    observed=[list(b) for b in oracle.csr_to_ballots(prefs, offsets)],
                    depth_factors=t.depth_factors().tolist(),
                    node_dump=t.dump_nodes().tolist(),
                    observed_map=take(t.observed()))
        n_ballots = nobs + 300
        ps = t.posterior_set(n_ballots, False, 4242)
        case["posterior_set"] = dict(n_ballots=n_ballots, engine_seed=4242, **take(ps))
        case["posterior_set_replace"] = dict(n_ballots=50, engine_seed=77, **take(t.posterior_set(50, True, 77)))
        case["irv_of_posterior_set"] = ref.irv((ps[0], ps[1]), ps[2], nc, "IRVSEED").tolist()
        case["sample_predictive"] = dict(n_ballots=25, seed="PREDICTIVE", **take(t.sample_predictive(25, "PREDICTIVE")))
        freq, _ = t.sample_posterior(64, n_ballots, 1, False, 2, "POSTERIOR")
        case["sample_posterior"] = dict(n_elections=64, n_ballots=n_ballots, n_threads=2, seed="POSTERIOR",
                                        freq=freq.tolist())
        cases.append(case)
        t.close()
    json.dump(dict(generated_by="tests/golden/make_golden.py from oracle/_ref/libdtree_ref.so "
                                "(unmodified reference core, g++ 13.3 libstdc++)", cases=cases),
              open(os.path.join(HERE, "ref_vectors.json"), "w"), separators=(",", ":"))


if __name__ == "__main__":
    oracle.build()
    ref = oracle.load("reference")
    make_wakehurst(ref)
    make_ref_vectors(ref)
    print("ok")
