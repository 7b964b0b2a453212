"""Host-side mirror of the reference's R-facing API for the hot path, over the C ABI.

The reference's host layer for this path is R (R/dtree.R, R/social_choice.R) over an Rcpp class
(src/R_tree.cpp). Neither R nor Rcpp exists in this image, so this module restates that layer in
Python with the same names, argument meaning, validation, warnings and errors, so that tests read
like tests/testthat/*.R. All sampling and IRV work happens in lib/libdtree_b200.so on the GPU.

    dtree = dirtree(candidates=list("ABCD"), a0=1.0)
    update(dtree, [["A", "B"], ["C"]])
    sample_posterior(dtree, n_elections=1000, n_ballots=100)   # {"A": 0.61, ...}
"""
import ctypes as C
import string
import warnings

import numpy as np

from . import _native as N

_seed_rng = np.random.RandomState()


def set_seed(seed):
    """Counterpart of R's set.seed(): makes the seed strings drawn by gseed() reproducible."""
    global _seed_rng
    _seed_rng = np.random.RandomState(seed)


def gseed():
    """R/dtree.R:682-684: a random 10-letter seed string for the native layer."""
    return "".join(_seed_rng.choice(list(string.ascii_uppercase), 10))


def _csr(index_ballots):
    n = len(index_ballots)
    lens = np.fromiter((len(b) for b in index_ballots), dtype=np.uint64, count=n)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    prefs = np.zeros(max(int(offsets[-1]), 1), dtype=np.uint8)
    o = 0
    for b in index_ballots:
        k = len(b)
        if k:
            prefs[o:o + k] = b
            o += k
    return prefs, offsets


def _take_ballots(handle):
    L = N.lib()
    n = L.dtree_ballots_size(handle)
    tp = L.dtree_ballots_total_prefs(handle)
    prefs = np.zeros(max(tp, 1), dtype=np.uint8)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    counts = np.zeros(max(n, 1), dtype=np.uint32)
    N.check(L.dtree_ballots_export(handle, prefs.ctypes.data_as(N.u8p), offsets.ctypes.data_as(N.u64p),
                                   counts.ctypes.data_as(N.u32p)))
    L.dtree_ballots_free(handle)
    return prefs[:tp], offsets, counts[:n]


class dirichlet_tree:
    """R6 class `dirichlet_tree` (R/dtree.R:74-472) over RDirichletTree (src/R_tree.h:32-96)."""

    def __init__(self, candidates, min_depth=0, max_depth=None, a0=1.0, vd=False, device=-1):
        candidates = list(candidates)
        if not all(isinstance(c, str) for c in candidates):
            raise TypeError("`candidates` must be a character vector, with each element representing a single candidate.")
        if len(set(candidates)) != len(candidates):
            raise ValueError("All `candidates` must be unique.")
        if max_depth is None:
            max_depth = len(candidates) - 1  # R/dtree.R:161
        if not (min_depth >= 0 and max_depth >= min_depth and len(candidates) >= max_depth):
            raise ValueError("min_depth and max_depth must satisfy: 0 <= min_depth <= max_depth <= n_candidates")
        if a0 < 0:
            raise ValueError("`a0` must be >= 0.")
        if not isinstance(vd, (bool, np.bool_)):
            raise TypeError("`vd` must be a logical.")
        self._candidates = candidates
        self._index = {c: i for i, c in enumerate(candidates)}
        self._L = N.lib()
        h = N.vp()
        N.check(self._L.dtree_create(len(candidates), int(min_depth), int(max_depth), float(a0), int(vd), int(device),
                                     C.byref(h)))
        self._h = h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.dtree_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- properties (R/dtree.R:81-147 active bindings -> RcppModules.cpp:21-28) -------------------
    @property
    def candidates(self):
        return list(self._candidates)

    @property
    def n_candidates(self):
        return int(self._L.dtree_n_candidates(self._h))

    @property
    def a0(self):
        return float(self._L.dtree_a0(self._h))

    @a0.setter
    def a0(self, v):
        if v < 0:
            raise ValueError("`a0` must be >= 0.")
        N.check(self._L.dtree_set_a0(self._h, float(v)))

    @property
    def min_depth(self):
        return int(self._L.dtree_min_depth(self._h))

    @min_depth.setter
    def min_depth(self, v):
        if v < 0:
            raise ValueError("`min_depth` must be >= 0.")
        if v > self.max_depth:  # R_tree.cpp:88-89
            raise ValueError("Cannot set `minDepth` to a value larger than `maxDepth`.")
        N.check(self._L.dtree_set_min_depth(self._h, int(v)))
        mask = int(self._L.dtree_observed_depth_mask(self._h))
        if any((mask >> d) & 1 for d in range(1, int(v))):  # R_tree.cpp:95-105
            warnings.warn("Ballots with fewer than `minDepth` preferences specified have been observed. Some "
                          "sampling techniques could now exhibit undefined behaviour. A Dirichlet Posterior can no "
                          "longer reduce to a tree of height 1. Consider setting `minDepth` to a value lower than "
                          "the length of the smallest ballot.")

    @property
    def max_depth(self):
        return int(self._L.dtree_max_depth(self._h))

    @max_depth.setter
    def max_depth(self, v):
        if v < self.min_depth:  # R_tree.cpp:109-110
            raise ValueError("Cannot set `maxDepth` to a value less than `minDepth`.")
        if v > self.n_candidates:
            raise ValueError("`max_depth` must be <= n_candidates.")
        N.check(self._L.dtree_set_max_depth(self._h, int(v)))

    @property
    def vd(self):
        return bool(self._L.dtree_vd(self._h))

    @vd.setter
    def vd(self, v):
        if not isinstance(v, (bool, np.bool_)):
            raise TypeError("`vd` must be a logical.")
        N.check(self._L.dtree_set_vd(self._h, int(v)))

    @property
    def n_observed(self):
        return int(self._L.dtree_n_observed(self._h))

    def tune(self, **kw):
        for k, v in kw.items():
            N.check(self._L.dtree_set_tuning(self._h, k.encode(), int(v)))
        return self

    def last_stats(self):
        buf = (C.c_uint64 * 24)()
        N.check(self._L.dtree_last_stats(self._h, buf, 24))
        keys = ["launches", "elections", "entries", "items", "gammas", "binomial_attempts", "irv_rereads",
                "scratch_bytes", "h2d_bytes", "d2h_bytes", "slots", "urn_items", "mode", "arena_runs", "build_launches",
                "lane_retries", "mirror_us", "plan_us", "launch_us", "finish_us"]
        return {k: int(buf[i]) for i, k in enumerate(keys)}

    _IMAGE_ARRAYS = {"node_base": (0, "uint32"), "node_total": (1, "uint32"), "slot_count": (2, "uint32"),
                     "slot_child": (3, "int32"), "slot_cand": (4, "uint8"), "obs_key": (5, "uint64"),
                     "obs_cnt": (6, "uint32"), "obs_first": (7, "uint8"), "obs_tally0": (8, "uint32")}

    def device_image(self):
        """The device mirror of the tree, copied back array by array (test hook, dtree_b200_debug.h)."""
        import numpy as np
        out = {}
        for name, (which, dt) in self._IMAGE_ARRAYS.items():
            need = C.c_uint64(0)
            N.check(self._L.dtree_debug_device_image(self._h, which, None, 0, C.byref(need)))
            a = np.zeros(need.value // np.dtype(dt).itemsize, dtype=dt)
            if a.size:
                N.check(self._L.dtree_debug_device_image(self._h, which, a.ctypes.data_as(C.c_void_p), a.nbytes, C.byref(need)))
            out[name] = a
        return out

    # ---- ballots ------------------------------------------------------------------------------------
    def _to_indices(self, ballots):
        out = []
        for b in ballots:
            if b is None:
                b = ()
            if isinstance(b, str):
                b = (b,)
            try:
                out.append([self._index[c] for c in b])
            except (KeyError, TypeError):
                raise ValueError("Unknown candidate encountered in ballot!")  # R_tree.cpp:30-31
        return out

    def update(self, ballots, counts=None):
        """R/dtree.R:263-303 -> R_tree.cpp:125-153. `ballots`: sequences of candidate names (an empty one is a
        blank ballot); `counts` optionally repeats each."""
        idx = self._to_indices(ballots)
        if counts is not None:
            idx = [b for b, c in zip(idx, counts) for _ in range(int(c))]
        self.update_indices(*_csr(idx))
        return self

    def update_indices(self, prefs, offsets):
        """Same with candidate indices in CSR form (what crosses the C ABI)."""
        prefs = np.ascontiguousarray(prefs, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n_short = C.c_uint64(0)
        N.check(self._L.dtree_update(self._h, prefs.ctypes.data_as(N.u8p), offsets.ctypes.data_as(N.u64p),
                                     len(offsets) - 1, C.byref(n_short)))
        if n_short.value:  # R_tree.cpp:139-147
            warnings.warn("Updating a Dirichlet-tree distribution with a ballot specifying fewer than `minDepth` "
                          "preferences. This introduces undefined behaviour to the sampling methods, and the "
                          "resulting posterior can no longer reduce to a Dirichlet distribution when using the `vd` "
                          "option. Consider setting `minDepth` to a value lower than the length of the smallest "
                          "ballot.")
        return self

    def reset(self):
        N.check(self._L.dtree_reset(self._h))
        return self

    def export_nodes(self):
        need = C.c_uint64(0)
        N.check(self._L.dtree_export_nodes(self._h, None, 0, C.byref(need)))
        buf = np.zeros(max(need.value, 1), dtype=np.uint32)
        N.check(self._L.dtree_export_nodes(self._h, buf.ctypes.data_as(N.u32p), need.value, C.byref(need)))
        return buf[:need.value]

    def observed_indices(self):
        h = N.vp()
        N.check(self._L.dtree_observed(self._h, C.byref(h)))
        return _take_ballots(h)

    def depth_factors(self):
        buf = np.zeros(64, dtype=np.float64)
        n = self._L.dtree_depth_factors(self._h, buf.ctypes.data_as(N.f64p), 64)
        return buf[:n].copy()

    # ---- sampling -----------------------------------------------------------------------------------
    def sample_posterior(self, n_elections, n_ballots, n_winners=1, replace=False, n_threads=None, seed=None):
        """R/dtree.R:365-401 -> R_tree.cpp:177-257. Returns {candidate: probability of being elected}."""
        counts = self.sample_posterior_counts(n_elections, n_ballots, n_winners, replace, n_threads, seed)
        return {c: counts[i] / float(n_elections) for i, c in enumerate(self._candidates)}

    def sample_posterior_counts(self, n_elections, n_ballots, n_winners=1, replace=False, n_threads=None, seed=None,
                                first_election=0, return_orders=False, devices=None, sc_function="irv"):
        if n_elections <= 0:
            raise ValueError("`n_elections` must be an integer > 0.")
        if n_elections > 0xFFFFFFFF or n_ballots > 0xFFFFFFFF:  # unsigned in the reference too (R_tree.h:86-88)
            raise ValueError("`n_elections` and `n_ballots` must be < 2^32.")
        if n_ballots < self.n_observed and not replace:
            raise ValueError("`n_ballots` must be an integer >= the number of observed ballots unless sampling "
                             "with replacement.")
        if n_threads is not None and n_threads < 1:
            raise ValueError("`n_threads` must be >= 1.")  # accepted for compatibility; the GPU ignores it
        seed = (gseed() if seed is None else seed).encode()
        nc = self.n_candidates
        wins = np.zeros(nc, dtype=np.uint64)
        if sc_function != "irv":  # R/social_choice.R:35-97 on every simulated election
            if sc_function == "stv":
                raise NotImplementedError("'stv' social choice not implemented.")  # social_choice.R:94-96
            if sc_function != "plurality":
                raise ValueError("`sc_function` must be one of 'plurality', 'irv' or 'stv'.")
            if return_orders or devices is not None or n_winners != 1:
                raise ValueError("plurality: single device, n_winners = 1, no elimination orders")
            N.check(self._L.dtree_sample_posterior_sc(self._h, 1, int(first_election), int(n_elections), int(n_ballots),
                                                      int(replace), seed, len(seed), wins.ctypes.data_as(N.u64p)))
            return wins
        if devices is not None:  # one process, several GPUs (dtree_sample_posterior_multi)
            if return_orders:
                raise ValueError("elimination orders are only returned by single-device calls")
            devs = (C.c_int * len(devices))(*[int(d) for d in devices])
            N.check(self._L.dtree_sample_posterior_multi(self._h, devs, len(devices), int(first_election),
                                                         int(n_elections), int(n_ballots), int(n_winners), int(replace),
                                                         seed, len(seed), wins.ctypes.data_as(N.u64p)))
            return wins
        orders = np.zeros((int(n_elections), nc), dtype=np.uint8) if return_orders else None
        N.check(self._L.dtree_sample_posterior(self._h, int(first_election), int(n_elections), int(n_ballots),
                                               int(n_winners), int(replace), seed, len(seed),
                                               wins.ctypes.data_as(N.u64p),
                                               orders.ctypes.data_as(N.u8p) if return_orders else None))
        return (wins, orders) if return_orders else wins

    def sample_predictive(self, n_ballots, seed=None):
        """R/dtree.R:429-470 -> R_tree.cpp:155-175. Returns the aggregated form the R side ends with
        (prefio::preferences(aggregate=TRUE)): a list of (ballot as a tuple of names, frequency)."""
        if n_ballots <= 0:
            raise ValueError("n_ballots must be an integer > 0")
        prefs, offsets, counts = self.sample_predictive_indices(n_ballots, seed)
        agg = {}
        for i in range(len(counts)):
            b = tuple(self._candidates[c] for c in prefs[int(offsets[i]):int(offsets[i + 1])])
            agg[b] = agg.get(b, 0) + int(counts[i])
        return sorted(agg.items(), key=lambda kv: (-kv[1], kv[0]))

    def sample_predictive_indices(self, n_ballots, seed=None):
        seed = (gseed() if seed is None else seed).encode()
        h = N.vp()
        N.check(self._L.dtree_sample_predictive(self._h, int(n_ballots), seed, len(seed), C.byref(h)))
        return _take_ballots(h)

    def posterior_set_indices(self, election, n_ballots, replace=False, seed="seed"):
        """All ballots of simulated election `election` of sample_posterior(seed=seed): (prefs, offsets, counts)."""
        seed = seed.encode()
        h = N.vp()
        N.check(self._L.dtree_posterior_set(self._h, int(election), int(n_ballots), int(replace), seed, len(seed),
                                            C.byref(h)))
        return _take_ballots(h)

    def __repr__(self):
        return ("Dirichlet-tree (a0=%g, min_depth=%d, max_depth=%d, vd=%s)\nCandidates: %s\nObserved ballots: %d"
                % (self.a0, self.min_depth, self.max_depth, self.vd, " ".join(self._candidates), self.n_observed))


# ---- S3 wrappers (R/dtree.R:521-679) -------------------------------------------------------------
def dirtree(candidates, min_depth=0, max_depth=None, a0=1.0, vd=False, device=-1):
    """R/dtree.R:521-533 (note: this wrapper defaults max_depth to length(candidates), the R6 class to one less)."""
    if max_depth is None:
        max_depth = len(candidates)
    return dirichlet_tree(candidates, min_depth, max_depth, a0, vd, device)


def update(dtree, ballots, counts=None):
    return dtree.update(ballots, counts)


def reset(dtree):
    return dtree.reset()


def sample_posterior(dtree, n_elections, n_ballots, n_winners=1, replace=False, n_threads=None, seed=None):
    return dtree.sample_posterior(n_elections, n_ballots, n_winners, replace, n_threads, seed)


def sample_predictive(dtree, n_ballots, seed=None):
    return dtree.sample_predictive(n_ballots, seed)


def irv_indices(prefs, offsets, counts, n_candidates, seed="seed", device=-1):
    """dtree_irv on index ballots in CSR form -> (elimination order incl. the winner last, tie rounds)."""
    prefs = np.ascontiguousarray(prefs, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    order = np.zeros(n_candidates, dtype=np.uint32)
    ties = C.c_uint32(0)
    s = seed.encode()
    N.check(N.lib().dtree_irv(prefs.ctypes.data_as(N.u8p), offsets.ctypes.data_as(N.u64p),
                              counts.ctypes.data_as(N.u32p), len(offsets) - 1, n_candidates, s, len(s), device,
                              order.ctypes.data_as(N.u32p), C.byref(ties)))
    return order, int(ties.value)


def social_choice(ballots, sc_function="plurality", n_winners=1, candidates=None, counts=None, seed=None):
    """R/social_choice.R:35-97. `ballots`: sequences of candidate names; `counts` optional multiplicities.
    irv -> {"elimination_order": [...], "winners": [...]} (R_social_choice.cpp:16-95)."""
    if sc_function not in ("plurality", "irv", "stv"):
        raise ValueError("`sc_function` must be one of 'plurality', 'irv' or 'stv'.")
    if sc_function == "stv":
        raise NotImplementedError("The stv social choice function has not been implemented yet.")  # social_choice.R:94-96
    if candidates is None:
        candidates = []
        for b in ballots:
            for c in (b or ()):
                if c not in candidates:
                    candidates.append(c)
    index = {c: i for i, c in enumerate(candidates)}
    counts = [1] * len(ballots) if counts is None else list(counts)
    if sc_function == "plurality":  # social_choice.R:58-83, done on the host in the reference too
        tally = np.zeros(len(candidates), dtype=np.int64)
        for b, c in zip(ballots, counts):
            if b:
                tally[index[b[0]]] += c
        return [candidates[i] for i in np.flatnonzero(tally == tally.max())]
    idx, cnt = [], []
    for b, c in zip(ballots, counts):
        if b is None:
            continue  # R_social_choice.cpp:40-41
        try:
            idx.append([index[x] for x in b])
        except KeyError:
            raise ValueError("Invalid candidate found during social-choice evaluation.")
        cnt.append(c)
    if n_winners < 1 or n_winners >= len(candidates):
        raise ValueError("`nWinners` must be >= 1 and <= the number of candidates.")
    if not idx:
        raise ValueError("No valid ballots for the IRV social choice function.")
    prefs, offsets = _csr(idx)
    order, _ = irv_indices(prefs, offsets, cnt, len(candidates), gseed() if seed is None else seed)
    k = len(candidates) - n_winners
    return {"elimination_order": [candidates[i] for i in order[:k]], "winners": [candidates[i] for i in order[k:]]}
