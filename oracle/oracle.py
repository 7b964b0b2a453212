"""ctypes loader for the two CPU checkers behind oracle/oracle_api.h.

TEST INFRASTRUCTURE ONLY. Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from the product package.

    port = load("port")        # oracle/libdtree_oracle.so  (restatement)
    ref  = load("reference")   # oracle/_ref/libdtree_ref.so (unmodified reference core)
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {
    "port": os.path.join(_HERE, "libdtree_oracle.so"),
    "reference": os.path.join(_HERE, "_ref", "libdtree_ref.so"),
}

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)


def build(quiet=True):
    """Compile the checkers (oracle/Makefile). _ref is rebuilt only where /root/reference exists."""
    subprocess.run(["make", "-C", _HERE] + (["-s"] if quiet else []), check=True)


def available(kind):
    return os.path.exists(_PATHS[kind])


def ballots_to_csr(ballots):
    """list of sequences of candidate indices -> (prefs u8, offsets u64)."""
    lens = np.fromiter((len(b) for b in ballots), dtype=np.uint64, count=len(ballots))
    offsets = np.zeros(len(ballots) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    prefs = np.zeros(max(int(offsets[-1]), 1), dtype=np.uint8)
    o = 0
    for b in ballots:
        n = len(b)
        if n:
            prefs[o:o + n] = b
            o += n
    return prefs, offsets


def csr_to_ballots(prefs, offsets):
    return [tuple(int(c) for c in prefs[int(offsets[i]):int(offsets[i + 1])]) for i in range(len(offsets) - 1)]


def parse_node_dump(words):
    """orc_dump_nodes / dtree_export_nodes words -> {prefix: (slot_candidates, counts, has_child)}."""
    out = {}
    i = 0
    w = [int(x) for x in words]
    while i < len(w):
        depth = w[i]; i += 1
        prefix = tuple(w[i:i + depth]); i += depth
        K = w[i]; i += 1
        cands = tuple(w[i:i + K]); i += K
        counts = tuple(w[i:i + K + 1]); i += K + 1
        has = tuple(w[i:i + K]); i += K
        assert prefix not in out
        out[prefix] = (cands, counts, has)
    return out


class OracleLib:
    def __init__(self, kind):
        self.kind = kind
        self.lib = L = C.CDLL(_PATHS[kind])
        L.orc_kind.restype = C.c_char_p
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int, C.c_char_p, C.c_size_t]
        L.orc_destroy.argtypes = [C.c_void_p]
        for nm, ty in (("min_depth", C.c_uint32), ("max_depth", C.c_uint32), ("a0", C.c_double), ("vd", C.c_int)):
            getattr(L, "orc_set_" + nm).argtypes = [C.c_void_p, ty]
        L.orc_depth_factors.restype = C.c_uint32
        L.orc_depth_factors.argtypes = [C.c_void_p, f64p, C.c_uint32]
        L.orc_reset.argtypes = [C.c_void_p]
        L.orc_update.argtypes = [C.c_void_p, u8p, u64p, C.c_uint64]
        L.orc_n_observed.restype = C.c_uint64
        L.orc_n_observed.argtypes = [C.c_void_p]
        L.orc_dump_nodes.restype = C.c_uint64
        L.orc_dump_nodes.argtypes = [C.c_void_p, u32p, C.c_uint64]
        L.orc_observed.restype = C.c_void_p
        L.orc_observed.argtypes = [C.c_void_p]
        L.orc_posterior_set.restype = C.c_void_p
        L.orc_posterior_set.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_uint32]
        L.orc_sample_predictive.restype = C.c_void_p
        L.orc_sample_predictive.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_size_t]
        L.orc_ballots_size.restype = C.c_uint64
        L.orc_ballots_size.argtypes = [C.c_void_p]
        L.orc_ballots_total_prefs.restype = C.c_uint64
        L.orc_ballots_total_prefs.argtypes = [C.c_void_p]
        L.orc_ballots_export.argtypes = [C.c_void_p, u8p, u64p, u32p]
        L.orc_ballots_free.argtypes = [C.c_void_p]
        L.orc_irv.argtypes = [u8p, u64p, u32p, C.c_uint64, C.c_uint32, C.c_char_p, C.c_size_t, u32p]
        L.orc_sample_posterior.restype = C.c_int
        L.orc_sample_posterior.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32,
                                           C.c_char_p, C.c_size_t, f64p, f64p]
        if kind == "port":
            L.orc_election_stats.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, u64p]
            L.orc_irv_check_order.restype = C.c_int
            L.orc_irv_check_order.argtypes = [u8p, u64p, u32p, C.c_uint64, C.c_uint32, u32p, u32p]
        assert L.orc_kind().decode() == kind

    # ---- tree -----------------------------------------------------------------------
    def tree(self, n_candidates, min_depth=0, max_depth=None, a0=1.0, vd=False, seed="12345"):
        return OracleTree(self, n_candidates, min_depth, n_candidates - 1 if max_depth is None else max_depth,
                          a0, vd, seed)

    # ---- stateless ------------------------------------------------------------------
    def irv(self, ballots, counts, n_candidates, seed="seed"):
        prefs, offsets = ballots if isinstance(ballots, tuple) else ballots_to_csr(ballots)
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        order = np.zeros(n_candidates, dtype=np.uint32)
        s = seed.encode()
        self.lib.orc_irv(prefs.ctypes.data_as(u8p), offsets.ctypes.data_as(u64p), counts.ctypes.data_as(u32p),
                         len(offsets) - 1, n_candidates, s, len(s), order.ctypes.data_as(u32p))
        return order

    def irv_check_order(self, ballots, counts, n_candidates, order):
        """(valid, tie_rounds) — port only."""
        prefs, offsets = ballots if isinstance(ballots, tuple) else ballots_to_csr(ballots)
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        order = np.ascontiguousarray(order, dtype=np.uint32)
        ties = C.c_uint32(0)
        ok = self.lib.orc_irv_check_order(prefs.ctypes.data_as(u8p), offsets.ctypes.data_as(u64p),
                                          counts.ctypes.data_as(u32p), len(offsets) - 1, n_candidates,
                                          order.ctypes.data_as(u32p), C.byref(ties))
        return bool(ok), int(ties.value)


class OracleTree:
    def __init__(self, lib, nc, min_depth, max_depth, a0, vd, seed):
        self.o = lib
        self.L = lib.lib
        self.nc = nc
        s = seed.encode()
        self.h = self.L.orc_create(nc, min_depth, max_depth, a0, int(vd), s, len(s))

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, **kw):
        for k, v in kw.items():
            getattr(self.L, "orc_set_" + k)(self.h, int(v) if k != "a0" else float(v))

    def depth_factors(self):
        buf = np.zeros(64, dtype=np.float64)
        n = self.L.orc_depth_factors(self.h, buf.ctypes.data_as(f64p), 64)
        return buf[:n].copy()

    def reset(self):
        self.L.orc_reset(self.h)

    def update(self, ballots):
        prefs, offsets = ballots if isinstance(ballots, tuple) else ballots_to_csr(ballots)
        self.L.orc_update(self.h, prefs.ctypes.data_as(u8p), offsets.ctypes.data_as(u64p), len(offsets) - 1)

    def n_observed(self):
        return int(self.L.orc_n_observed(self.h))

    def dump_nodes(self):
        n = self.L.orc_dump_nodes(self.h, None, 0)
        buf = np.zeros(n, dtype=np.uint32)
        self.L.orc_dump_nodes(self.h, buf.ctypes.data_as(u32p), n)
        return buf

    def _take(self, b):
        n = self.L.orc_ballots_size(b)
        tp = self.L.orc_ballots_total_prefs(b)
        prefs = np.zeros(max(tp, 1), dtype=np.uint8)
        offsets = np.zeros(n + 1, dtype=np.uint64)
        counts = np.zeros(n, dtype=np.uint32)
        self.L.orc_ballots_export(b, prefs.ctypes.data_as(u8p), offsets.ctypes.data_as(u64p), counts.ctypes.data_as(u32p))
        self.L.orc_ballots_free(b)
        return prefs, offsets, counts

    def observed(self):
        return self._take(self.L.orc_observed(self.h))

    def posterior_set(self, n_ballots, replace=False, engine_seed=1):
        return self._take(self.L.orc_posterior_set(self.h, n_ballots, int(replace), engine_seed))

    def sample_predictive(self, n_ballots, seed="seed"):
        s = seed.encode()
        return self._take(self.L.orc_sample_predictive(self.h, n_ballots, s, len(s)))

    def sample_posterior(self, n_elections, n_ballots, n_winners=1, replace=False, n_threads=2, seed="seed"):
        """-> (win frequencies [nc], seconds spent in the thread pool)."""
        out = np.zeros(self.nc, dtype=np.float64)
        secs = C.c_double(0)
        s = seed.encode()
        rc = self.L.orc_sample_posterior(self.h, n_elections, n_ballots, n_winners, int(replace), n_threads,
                                         s, len(s), out.ctypes.data_as(f64p), C.byref(secs))
        if rc:
            raise ValueError("`nBallots` must be larger than the number of ballots observed to obtain the posterior.")
        return out, secs.value

    def election_stats(self, n_ballots, replace=False, engine_seed=1):
        """port only: dict(S, E, R, V_gamma, dirmult, binomials) for one election."""
        out = np.zeros(8, dtype=np.uint64)
        self.L.orc_election_stats(self.h, n_ballots, int(replace), engine_seed, out.ctypes.data_as(u64p))
        return dict(S=int(out[0]), E=int(out[1]), R=int(out[2]), V_gamma=int(out[3]), dirmult=int(out[4]),
                    binomials=int(out[5]))


_cache = {}


def load(kind="port"):
    if kind not in _cache:
        if not available(kind):
            build()
        _cache[kind] = OracleLib(kind)
    return _cache[kind]
