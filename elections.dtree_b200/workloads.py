"""Pinned synthetic inputs for the five BASELINE.json configurations.

Host-side utility shared by tests/ and bench.py; pure numpy, nothing here touches the GPU.
Ballots are tuples of 0-based candidate indices in preference order.
"""
import json
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def plackett_luce_ballots(n_candidates, gamma, n, seed, lengths=None):
    """G(nc, gamma, L): Plackett-Luce rankings with weights w_k = exp(-gamma k), truncated.

    lengths: None for full rankings, or a pmf over lengths 0..nc (index = length).
    Returns (prefs u8 [sum len], offsets u64 [n+1]) in CSR form.
    """
    rng = np.random.default_rng(seed)
    logw = -gamma * np.arange(n_candidates)
    # Gumbel-max: sorting log w + Gumbel noise in decreasing order is a Plackett-Luce draw
    keys = logw[None, :] + rng.gumbel(size=(n, n_candidates))
    perm = np.argsort(-keys, axis=1).astype(np.uint8)
    if lengths is None:
        lens = np.full(n, n_candidates, dtype=np.uint64)
    else:
        pmf = np.asarray(lengths, dtype=np.float64)
        lens = rng.choice(len(pmf), size=n, p=pmf / pmf.sum()).astype(np.uint64)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    mask = np.arange(n_candidates)[None, :] < lens[:, None]
    prefs = perm[mask]
    if prefs.size == 0:
        prefs = np.zeros(1, dtype=np.uint8)
    return np.ascontiguousarray(prefs), offsets


def wakehurst():
    """The reference's only fixture (tests/data/wakehurst2023.soi), shipped with the package as data/wakehurst2023.json
    (the tests check it against the copy under tests/golden/)."""
    with open(os.path.join(_DATA, "wakehurst2023.json")) as f:
        return json.load(f)


def wakehurst_sample(n, seed):
    """n ballots drawn without replacement from the Wakehurst multiset -> CSR."""
    w = wakehurst()
    counts = np.asarray(w["counts"], dtype=np.int64)
    rng = np.random.default_rng(seed)
    idx = rng.choice(np.repeat(np.arange(len(counts)), counts), size=n, replace=False)
    ballots = [w["ballots"][i] for i in idx]
    lens = np.fromiter((len(b) for b in ballots), dtype=np.uint64, count=n)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    prefs = np.fromiter((c for b in ballots for c in b), dtype=np.uint8, count=int(offsets[-1]))
    return prefs, offsets


CONFIGS = {
    # name: tree parameters, observed-ballot recipe, sample_posterior arguments
    "C1": dict(n_candidates=4, min_depth=0, max_depth=3, a0=1.0, vd=False,
               observed=("pl", 0.3, 1000, 1), n_elections=1000, n_ballots=10_000),
    "C2": dict(n_candidates=6, min_depth=0, max_depth=5, a0=1.0, vd=False,
               observed=("wakehurst", 1000, 2), n_elections=10_000, n_ballots=50_000),
    "C3": dict(n_candidates=10, min_depth=0, max_depth=10, a0=1.0, vd=False,
               observed=("pl", 0.5, 200_000, 3), n_elections=100_000, n_ballots=1_000_000),
    "C4": dict(n_candidates=20, min_depth=0, max_depth=20, a0=1.0, vd=False,
               observed=("pl", 0.3, 5000, 4), n_elections=1_000_000, n_ballots=1_000_000),
    "C5": dict(n_candidates=8, min_depth=7, max_depth=7, a0=1.0, vd=True,
               observed=("pl", 0.4, 1000, 5), n_elections=100_000, n_ballots=100_000),
}


def observed_ballots(name):
    cfg = CONFIGS[name]
    rec = cfg["observed"]
    if rec[0] == "pl":
        return plackett_luce_ballots(cfg["n_candidates"], rec[1], rec[2], rec[3])
    return wakehurst_sample(rec[1], rec[2])


def candidate_names(n):
    return ["C%02d" % i for i in range(n)]
