"""ctypes declarations for lib/libdtree_b200.so (include/dtree_b200.h, include/dtree_b200_debug.h).

Loading fails loudly when the library has not been built: there is no Python or CPU fallback for
anything that samples or runs a social choice function.
"""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# DTREE_B200_LIB selects another build of the same library (profiling variants); there is still no fallback.
LIB_PATH = os.environ.get("DTREE_B200_LIB") or os.path.join(_HERE, "lib", "libdtree_b200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")

OK, ERR_ARG, ERR_CUDA, ERR_STATE, ERR_ABORTED = 0, 1, 2, 3, 4

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)
vp = C.c_void_p
ABORT_CB = C.CFUNCTYPE(C.c_int, C.c_void_p)

_SIGS = {
    "dtree_last_error": (C.c_char_p, []),
    "dtree_version": (C.c_char_p, []),
    "dtree_device_count": (C.c_int, []),
    "dtree_create": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int, C.c_int, C.POINTER(vp)]),
    "dtree_destroy": (C.c_int, [vp]),
    "dtree_set_min_depth": (C.c_int, [vp, C.c_uint32]),
    "dtree_set_max_depth": (C.c_int, [vp, C.c_uint32]),
    "dtree_set_a0": (C.c_int, [vp, C.c_double]),
    "dtree_set_vd": (C.c_int, [vp, C.c_int]),
    "dtree_n_candidates": (C.c_uint32, [vp]),
    "dtree_min_depth": (C.c_uint32, [vp]),
    "dtree_max_depth": (C.c_uint32, [vp]),
    "dtree_a0": (C.c_double, [vp]),
    "dtree_vd": (C.c_int, [vp]),
    "dtree_depth_factors": (C.c_uint32, [vp, f64p, C.c_uint32]),
    "dtree_update": (C.c_int, [vp, u8p, u64p, C.c_uint64, u64p]),
    "dtree_reset": (C.c_int, [vp]),
    "dtree_n_observed": (C.c_uint64, [vp]),
    "dtree_observed_depth_mask": (C.c_uint64, [vp]),
    "dtree_n_nodes": (C.c_uint64, [vp]),
    "dtree_export_nodes": (C.c_int, [vp, u32p, C.c_uint64, u64p]),
    "dtree_observed": (C.c_int, [vp, C.POINTER(vp)]),
    "dtree_sample_posterior": (C.c_int, [vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_char_p,
                                         C.c_size_t, u64p, u8p]),
    "dtree_sample_posterior_sc": (C.c_int, [vp, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_char_p,
                                            C.c_size_t, u64p]),
    "dtree_sample_posterior_device": (C.c_int, [vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                                C.c_char_p, C.c_size_t, vp, vp]),
    "dtree_sample_posterior_multi": (C.c_int, [vp, C.POINTER(C.c_int), C.c_int, C.c_uint64, C.c_uint32, C.c_uint32,
                                               C.c_uint32, C.c_int, C.c_char_p, C.c_size_t, u64p]),
    "dtree_sample_posterior_finish": (C.c_int, [vp, vp]),
    "dtree_sync_device": (C.c_int, [vp]),
    "dtree_invalidate_device": (C.c_int, [vp]),
    "dtree_sample_predictive": (C.c_int, [vp, C.c_uint32, C.c_char_p, C.c_size_t, C.POINTER(vp)]),
    "dtree_posterior_set": (C.c_int, [vp, C.c_uint64, C.c_uint32, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(vp)]),
    "dtree_irv": (C.c_int, [u8p, u64p, u32p, C.c_uint64, C.c_uint32, C.c_char_p, C.c_size_t, C.c_int, u32p, u32p]),
    "dtree_ballots_size": (C.c_uint64, [vp]),
    "dtree_ballots_total_prefs": (C.c_uint64, [vp]),
    "dtree_ballots_export": (C.c_int, [vp, u8p, u64p, u32p]),
    "dtree_ballots_free": (None, [vp]),
    "dtree_set_abort_callback": (C.c_int, [vp, ABORT_CB, vp]),
    "dtree_set_tuning": (C.c_int, [vp, C.c_char_p, C.c_int64]),
    "dtree_last_stats": (C.c_int, [vp, u64p, C.c_uint32]),
    # debug hooks
    "dtree_debug_philox": (C.c_int, [u32p, u32p, u32p, C.c_int]),
    "dtree_debug_hash_seed": (C.c_uint64, [C.c_char_p, C.c_size_t]),
    "dtree_debug_gamma": (C.c_int, [C.c_double, C.c_int, C.c_uint64, C.c_uint32, f32p, C.c_int]),
    "dtree_debug_binomial": (C.c_int, [C.c_uint32, C.c_double, C.c_uint64, C.c_uint32, u32p, C.c_int]),
    "dtree_debug_gamma_ex": (C.c_int, [C.c_double, C.c_int, C.c_uint64, C.c_uint32, f32p, C.c_int, C.c_int]),
    "dtree_debug_binomial_ex": (C.c_int, [C.c_uint32, C.c_double, C.c_uint64, C.c_uint32, u32p, C.c_int, C.c_int]),
    "dtree_debug_binv_law": (C.c_int, [C.c_uint32, C.c_double, u32p, C.c_int]),
    "dtree_debug_device_image": (C.c_int, [vp, C.c_int, vp, C.c_uint64, u64p]),
    "dtree_debug_lane_shape": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, u32p, u32p]),
}

_lib = None


def declared_symbols():
    """Every function the public headers declare (used by the CPU test that the .so exports them all)."""
    names = []
    for h in ("dtree_b200.h", "dtree_b200_debug.h"):
        src = open(os.path.join(INCLUDE_DIR, h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(dtree_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
                "`make -C elections.dtree_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


class DtreeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def check(rc):
    if rc != OK:
        raise DtreeError(rc, lib().dtree_last_error().decode())
