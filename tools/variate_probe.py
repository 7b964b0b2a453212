"""The variate stages on their own (north_star: "FP32/SFU pipe utilisation for the Gamma kernel"): the debug hooks
run the device Gamma and binomial samplers of csrc/variates.cuh over 2^25 independent draws; timed here, profiled by
tools/profile_variates.sh. usage: variate_probe.py gamma|binomial"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
import numpy as np
pkg = load_pkg()
L = pkg._native.lib()
which = sys.argv[1] if len(sys.argv) > 1 else "gamma"
n = 1 << 25
if which == "gamma":
    out = np.zeros(n, dtype=np.float32)
    for alpha in (1.0, 1.0e5):
        for rep in range(2):
            t0 = time.perf_counter()
            assert L.dtree_debug_gamma(alpha, 0, 11, n, out.ctypes.data_as(C.POINTER(C.c_float)), 0) == 0
            dt = time.perf_counter() - t0
        print("gamma alpha=%g: %d variates, %.1f ms incl. 128 MB read-back, mean %.4g" % (alpha, n, dt * 1e3, float(out[:1 << 20].mean())))
else:
    out = np.zeros(n, dtype=np.uint32)
    for trials, p in ((800_000, 0.39), (800_000, 3e-5), (5000, 0.2)):
        for rep in range(2):
            t0 = time.perf_counter()
            assert L.dtree_debug_binomial(trials, p, 12, n, out.ctypes.data_as(C.POINTER(C.c_uint32)), 0) == 0
            dt = time.perf_counter() - t0
        print("binomial n=%d p=%g: %d variates, %.1f ms incl. 128 MB read-back, mean %.6g" % (trials, p, n, dt * 1e3, float(out[:1 << 20].mean())))
