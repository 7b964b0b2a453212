"""GPU tests that bound the arithmetic of the device samplers (csrc/variates.cuh).

 * guard bands: the f32 accept/reject tests with f64 fall-back must make exactly the decisions of the f64 tests —
   both variants are run over ~10^8 draws and must give identical outputs;
 * BINV: its exact output law (all 2^23 uniforms pushed through it) against the binomial pmf, total variation;
 * shapes >= 2^22 (all-double path), the vd prior at 20 candidates (20! through the float cast), a node whose slot
   observed more than 2^25 ballots.
"""
import ctypes as C
import warnings

import numpy as np
import pytest
from scipy import stats

pytestmark = pytest.mark.gpu

F32P = C.POINTER(C.c_float)
U32P = C.POINTER(C.c_uint32)


@pytest.fixture(scope="module")
def L(pkg):
    lib = pkg._native.lib()
    if lib.dtree_device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests cannot run (there is no CPU fallback)")
    return lib


@pytest.mark.parametrize("n_trials,p", [(60, 0.5), (64, 0.47), (200, 0.2), (1000, 0.03), (1000, 0.5), (5000, 0.007),
                                        (50_000, 0.3), (50_000, 0.0007), (300_000, 0.5), (300_000, 0.011),
                                        (1_000_000, 0.4), (1_048_575, 0.5), (1_000_000, 3.1e-5), (999_999, 0.25),
                                        (2_000_000, 0.3), (123_456, 0.9)])
def test_btrs_guard_bands_decide_as_the_double_test(L, n_trials, p):
    """4 M draws per case: guard-banded f32 decisions (what the kernels run) == f64 decisions, draw for draw."""
    n = 4_000_000
    a = np.zeros(n, dtype=np.uint32)
    b = np.zeros(n, dtype=np.uint32)
    assert L.dtree_debug_binomial_ex(n_trials, p, 4242, n, a.ctypes.data_as(U32P), 0, 0) == 0
    assert L.dtree_debug_binomial_ex(n_trials, p, 4242, n, b.ctypes.data_as(U32P), 0, 1) == 0
    diff = int(np.count_nonzero(a != b))
    assert diff == 0, (diff, n_trials, p)
    assert a.max() <= n_trials
    mean, sd = n_trials * p, np.sqrt(n_trials * p * (1 - p))
    assert abs(a.astype(np.float64).mean() - mean) < 5 * sd / np.sqrt(n) + 1e-6 * mean


@pytest.mark.parametrize("alpha,log_space", [(1.0, 0), (1.37, 0), (2.5, 0), (8.0, 0), (40.0, 0), (60.0, 0), (75.0, 0),
                                             (1000.0, 0), (1e5, 0), (3e6, 0), (0.3, 1), (0.05, 1), (1.0, 1)])
def test_gamma_guard_bands_decide_as_the_double_test(L, alpha, log_space):
    n = 4_000_000
    a = np.zeros(n, dtype=np.float32)
    b = np.zeros(n, dtype=np.float32)
    assert L.dtree_debug_gamma_ex(alpha, log_space, 9, n, a.ctypes.data_as(F32P), 0, 0) == 0
    assert L.dtree_debug_gamma_ex(alpha, log_space, 9, n, b.ctypes.data_as(F32P), 0, 1) == 0
    assert np.array_equal(a, b), int(np.count_nonzero(a != b))


BINV_TV_BOUND = 6e-6


@pytest.mark.parametrize("n_trials,p", [(1, 0.5), (2, 0.3), (9, 0.5), (50, 0.5), (59, 0.5), (192, 0.15), (1000, 0.029),
                                        (1000, 0.001), (100_000, 2.9e-4), (100_000, 1e-6), (1_000_000, 2.99e-5),
                                        (4_000_000_000, 7e-9), (4_000_000_000, 1e-12), (300, 0.0999), (37, 0.4)])
def test_binv_law_total_variation(L, n_trials, p):
    """BINV's exact output law (it is a deterministic function of one 23-bit uniform: every one of the 2^23 inputs is
    evaluated on the device) against the binomial pmf at the f32 value of p the sampler is given. The bound covers
    the 2^-23 resolution of the uniform and the f32 recurrence; it is what `variates.cuh` states."""
    counts = np.zeros(256, dtype=np.uint32)
    assert L.dtree_debug_binv_law(n_trials, p, counts.ctypes.data_as(U32P), 0) == 0
    assert int(counts.sum()) == 1 << 23
    law = counts.astype(np.float64) / float(1 << 23)
    pf = float(np.float32(p))
    k = np.arange(256)
    pmf = stats.binom.pmf(k, n_trials, pf)
    tail = max(0.0, 1.0 - float(stats.binom.cdf(255, n_trials, pf)))
    tv = 0.5 * (np.abs(law - pmf).sum() + tail)
    assert tv < BINV_TV_BOUND, (tv, n_trials, p)
    assert counts[min(n_trials, 255) + 1:].sum() == 0  # never more successes than trials


def test_shapes_of_two_to_the_22_and_more_use_the_double_path(L):
    """outcome_gamma with 2^25 + 7 observed ballots and a prior of 0.5: f32 cannot even represent the shape (spacing
    4). KS against the exact law and mean within 5 standard errors; a shift by the f32 rounding (up to 2.0) would be
    0.0003 sigma per draw — not detectable by sampling, which is why the path is all-double by construction."""
    alpha, n = float((1 << 25) + 7) + 0.5, 1_000_000
    out = np.zeros(n, dtype=np.float32)
    assert L.dtree_debug_gamma(alpha, 0, 5, n, out.ctypes.data_as(F32P), 0) == 0
    x = out.astype(np.float64)
    z = (x - alpha) / np.sqrt(alpha)
    assert abs(z.mean()) < 5 / np.sqrt(n) + 2e-3  # f32 output spacing is 4 = 7e-4 sigma
    assert abs(z.var() - 1) < 0.01
    assert stats.kstest(x, lambda v: stats.gamma.cdf(v, alpha)).pvalue > 1e-4
    out_l = np.zeros(n, dtype=np.float32)
    assert L.dtree_debug_gamma(alpha, 1, 5, n, out_l.ctypes.data_as(F32P), 0) == 0
    assert np.allclose(np.exp(out_l.astype(np.float64)), x, rtol=1e-5)


def test_slot_with_more_than_two_to_the_25_observed_ballots(pkg, L):
    """A tree whose root slot observed 2^25 + 5 ballots (f32 spacing 4 up there): counts stay exact integers on the
    device, the three kernels agree, ballots are conserved, and the first-preference split of the sampled ballots has
    the Dirichlet-multinomial mean n alpha_i / A computed in double."""
    nc = 3
    big = (1 << 25) + 5
    prefs = np.concatenate([np.tile(np.array([0, 1], dtype=np.uint8), big), np.array([1, 2, 2, 0, 1], dtype=np.uint8)])
    offsets = np.concatenate([np.arange(0, 2 * big + 1, 2, dtype=np.uint64), 2 * big + np.array([2, 4, 5], dtype=np.uint64)])
    t = pkg.dirichlet_tree(pkg.workloads.candidate_names(nc), 0, nc, 1.0, False, device=0)
    t.update_indices(prefs, offsets)
    n_obs = big + 3
    assert t.n_observed == n_obs
    img = t.device_image()
    assert int(img["slot_count"][0]) == big  # u32 on the device, not a float
    n_new = 200_000
    tot = np.zeros(nc + 1)
    n_el = 300
    for e in range(n_el):
        p, off, cnt = t.posterior_set_indices(e, n_obs + n_new, False, "BIG25")
        assert int(cnt.astype(np.int64).sum()) == n_obs + n_new
    for e in range(n_el):
        p, off, cnt = t.posterior_set_indices(e, n_new, True, "BIG25")
        first = np.where(np.diff(off.astype(np.int64)) > 0, p[np.minimum(off[:-1].astype(np.int64), len(p) - 1)], nc)
        np.add.at(tot, first, cnt)
    alpha = np.array([big + 1.0, 2.0 + 1.0, 1.0 + 1.0, 0.0 + 1.0])
    A = alpha.sum()
    mean = n_new * alpha / A
    var = n_new * (alpha / A) * (1 - alpha / A) * (A + n_new) / (A + 1)
    assert np.all(np.abs(tot / n_el - mean) < 5 * np.sqrt(var / n_el) + 1e-9), (tot / n_el, mean)
    res = []
    for mode in (0, 1, 2):
        t.tune(mode=mode)
        res.append(t.sample_posterior_counts(500, n_obs + n_new, seed="BIG25").tolist())
    assert res[0] == res[1] == res[2] and sum(res[0]) == 500


def test_vd_prior_at_twenty_candidates(pkg, L):
    """vd with 20 candidates: the root's prior parameter is a0 * 20!/1 = 2.4e18 (beyond 2^53, irv_node.cpp:13-25) and
    crosses the ABI's float cast; sampling must stay finite and, on an empty tree with min_depth = max_depth, flat:
    first preferences of 400 x 20 000 ballots uniform over the candidates (they are Dirichlet(2.4e18/20 ...)-multinomial,
    i.e. multinomial to 1e-14), and a0 so large that a0 * depthFactor leaves the float range is refused."""
    nc = 20
    t = pkg.dirichlet_tree(pkg.workloads.candidate_names(nc), nc - 1, nc - 1, 1.0, True, device=0)
    f = t.depth_factors()
    assert f[0] > 2.0 ** 53 and np.isfinite(f).all()
    tot = np.zeros(nc)
    n_el, n = 400, 20_000
    for e in range(n_el):
        p, off, cnt = t.posterior_set_indices(e, n, True, "VD20")
        lens = np.diff(off.astype(np.int64))
        assert lens.min() == nc - 1 and lens.max() == nc - 1 and int(cnt.sum()) == n
        np.add.at(tot, p[off[:-1].astype(np.int64)], cnt)
    exp = n_el * n / nc
    chi2 = ((tot - exp) ** 2 / exp).sum()
    assert stats.chi2.sf(chi2, nc - 1) > 1e-4, chi2
    w = t.sample_posterior_counts(2000, 5000, seed="VD20")
    assert int(w.sum()) == 2000 and (w > 40).all()  # 20 exchangeable candidates: 100 wins each on average
    t.a0 = 1e20
    with pytest.raises(pkg._native.DtreeError) as ei:
        t.sample_posterior_counts(10, 100, seed="VD20")
    assert ei.value.code == pkg._native.ERR_ARG
