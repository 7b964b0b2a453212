"""CPU check of the formula the device binomial sampler's f32 acceptance test relies on (csrc/variates.cuh,
btrs32_accept), against double-precision lgamma. No GPU: this validates the mathematics and its f32 conditioning;
tests/test_gpu_variates.py validates the device arithmetic (guard-banded f32 decisions == f64 decisions, draw for draw).

Loader's saddle-point form of log f(k) - log f(m) is evaluated in numpy float32 arithmetic, with the 2-ulp divisions
of the device emulated by a random +-2 ulp perturbation, and must stay well inside the guard band.
"""
import numpy as np
from scipy.special import gammaln

f32 = np.float32


def cases(rng, n_cases, hi_exp):
    out = 0
    while out < n_cases:
        n = min(int(10 ** rng.uniform(1.8, hi_exp)), 2 ** 20 - 1)
        p = float(f32(10 ** rng.uniform(-5.5, np.log10(0.5))))
        if n * p < 30:
            continue
        out += 1
        yield n, p


SF = np.array([0.0] + [float(gammaln(x + 1) - ((x + 0.5) * np.log(x) - x + 0.5 * np.log(2 * np.pi))) for x in range(1, 16)])


def test_stirling_error_table_of_the_kernel():
    src = open(__file__.replace("tests/test_variates_cpu.py", "elections.dtree_b200/csrc/variates.cuh")).read()
    body = src[src.index("kStirlErr[16] = {") + len("kStirlErr[16] = {"):]
    vals = [float(v.strip().rstrip("f")) for v in body[:body.index("}")].split(",")]
    assert np.allclose(vals, SF, rtol=3e-7, atol=1e-9)


def test_saddle_point_form_in_f32_stays_inside_the_guard_band():
    rng = np.random.default_rng(4)

    def fdiv(a, b):
        qt = (a / b).astype(f32)
        return (qt * (f32(1) + f32(2.4e-7) * rng.uniform(-1, 1, qt.shape).astype(f32))).astype(f32)

    def flog(x):  # __logf: absolute error 2^-21 around 1, 3 ulp elsewhere
        y = np.log(x.astype(np.float64))
        return (y + rng.uniform(-1, 1, y.shape) * np.maximum(2.0 ** -21, 3 * 1.2e-7 * np.abs(y))).astype(f32)

    def stirlerr(x):
        r = fdiv(np.full_like(x, 1), np.maximum(x, f32(1)))
        r2 = r * r
        ser = r * (f32(1 / 12) - r2 * f32(1 / 360))
        return np.where(x < 16, SF[np.minimum(x.astype(np.int64), 15)].astype(f32), ser)

    def bd0(x, dx, s):
        v = fdiv(dx, s)
        v2 = v * v
        t = f32(1 / 21)
        for c in (1 / 19, 1 / 17, 1 / 15, 1 / 13, 1 / 11, 1 / 9, 1 / 7, 1 / 5, 1 / 3):
            t = t * v2 + f32(c)
        t3 = (f32(1 / 7) * v2 + f32(1 / 5)) * v2 + f32(1 / 3)  # the kernel's short series for |v| < 0.05
        t = np.where(v2 < f32(2.5e-3), t3, t)
        return dx * v + f32(2) * x * (v * v2 * t), np.abs(v)

    def bd0_near(x, dx, s):
        v = fdiv(dx, s)
        return dx * v + f32(2 / 3) * x * (v * v * v)

    worst, checked = 0.0, 0
    for n, p in cases(rng, 6000, 6.02):
        q = 1 - p
        m = np.floor((n + 1) * p)
        sd = np.sqrt(n * p * q)
        ks = np.unique(np.clip(np.round(m + rng.normal(0, 2.2, 200) * sd), 1, n - 1))
        lf = gammaln(m + 1) + gammaln(n - m + 1) - gammaln(ks + 1) - gammaln(n - ks + 1) + (ks - m) * np.log(p / q)
        npd = n * p
        dxk = (ks - npd).astype(f32)
        dxm = np.full(len(ks), m - npd).astype(f32)
        kf, mf = ks.astype(f32), np.full(len(ks), m).astype(f32)
        nk, nm = (n - ks).astype(f32), np.full(len(ks), n - m).astype(f32)
        npf, nqf = f32(npd), f32(n - npd)
        delta = (ks - m).astype(f32)
        ra, rb = -fdiv(delta, kf), fdiv(delta, nk)
        half = f32(0.5) * flog(f32(1) + (ra + rb + ra * rb))
        st = stirlerr(kf) + stirlerr(nk) - stirlerr(mf) - stirlerr(nm)
        b1, v1 = bd0(kf, dxk, kf + npf)
        b2, v2 = bd0(nk, -dxk, nk + nqf)
        lf32 = (half - st - (b1 + b2 - bd0_near(mf, dxm, mf + npf) - bd0_near(nm, -dxm, nm + nqf))).astype(np.float64)
        ok = (v1 <= 0.4) & (v2 <= 0.4)
        if not ok.any():
            continue
        guard = 3e-5 + 8e-6 * np.abs(lf32[ok])
        worst = max(worst, float(np.max(np.abs(lf32 - lf)[ok] / guard)))
        checked += int(ok.sum())
    assert checked > 300_000 and worst < 0.25, (checked, worst)  # measured 0.06: a factor 16 inside the band
