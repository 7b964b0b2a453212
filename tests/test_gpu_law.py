"""GPU tests of the sampler's law deep in the tree and at 20 candidates.

The round-1 distribution tests sat at <= 6 candidates and at the root. Here:
 * closed form: the split of the ballots that reach an interior INITIALISED node and a LAZY (never observed) node is
   Beta-binomial given how many reached it (IRVNode::sample, irv_node.cpp:107-174; lazyIRVBallots, :27-84);
 * against the oracle on a reduced C4 shape (20 candidates, lazy subtrees below a small observed tree): ballot
   lengths, the candidate ranked at depth 4, the mass below the commonest observed prefixes, distinct ballot types.
Elections are materialised by dtree_posterior_set (simulate_kernel); the deferred kernels are tied to it election by
election by tests/test_gpu_parity.py::test_deferred_equals_materialised.
"""
import warnings

import numpy as np
import pytest
from scipy import stats

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(pkg):
    lib = pkg._native.lib()
    if lib.dtree_device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests cannot run (there is no CPU fallback)")
    return lib


def ballots_of(p, off, cnt):
    return list(zip(oracle.csr_to_ballots(p, off), (int(c) for c in cnt)))


def beta_binomial_z(x, n, a, A):
    """Standardised Beta-binomial residuals of x successes out of n with parameters (a, A - a); rows with n = 0 dropped."""
    x, n = np.asarray(x, dtype=np.float64), np.asarray(n, dtype=np.float64)
    keep = n > 0
    x, n = x[keep], n[keep]
    mean = n * a / A
    var = n * (a / A) * (1 - a / A) * (A + n) / (A + 1)
    return (x - mean) / np.sqrt(var)


@pytest.mark.parametrize("nc,mn,mx,a0,vd", [(20, 0, 20, 1.0, False), (12, 2, 9, 0.4, False), (20, 3, 12, 2.0, True)])
def test_interior_and_lazy_nodes_split_beta_binomially(pkg, L, nc, mn, mx, a0, vd):
    """For a node with prefix P: given that n sampled ballots reach it, the number continuing with outcome j is
    Beta-binomial(n, alpha_j, A - alpha_j) with alpha = observed counts + a0' (initialised node) or a0' everywhere
    (lazy node), a0' = a0 * depthFactor if vd. The standardised residuals over 1500 elections must have mean 0 and
    second moment 1, each within 4.5 standard errors (the latter's from the sample's own fourth moment)."""
    pmf = np.ones(nc + 1)
    pmf[:max(mn, 3)] = 0
    W = pkg.workloads
    prefs, offsets = W.plackett_luce_ballots(nc, 0.9, 150, nc + 3, pmf)  # steep: few distinct first preferences
    t = pkg.dirichlet_tree(W.candidate_names(nc), mn, mx, a0, vd, device=0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t.update_indices(prefs, offsets)
    dump = oracle.parse_node_dump(t.export_nodes())
    factors = t.depth_factors() if vd else None
    prior = lambda depth: a0 * (factors[depth] if vd else 1.0)
    # an initialised node at depth 2 (the commonest observed 2-prefix) and lazy nodes at depths 1 and 3
    obs = oracle.csr_to_ballots(prefs, offsets)
    two = {}
    for b in obs:
        two[b[:2]] = two.get(b[:2], 0) + 1
    P_init = max(two, key=two.get)
    assert P_init in dump
    firsts = {b[0] for b in obs}
    c_lazy = next(c for c in range(nc - 1, -1, -1) if c not in firsts)
    P_lazy1 = (c_lazy,)
    assert P_lazy1 not in dump
    P_lazy3 = P_init + (next(c for c in range(nc - 1, -1, -1) if c not in P_init and (P_init + (c,)) not in dump),)
    assert P_lazy3 not in dump
    n_el, n = 1500, 4000
    nodes = {P_init: "init", P_lazy1: "lazy", P_lazy3: "lazy"}
    reach = {P: np.zeros(n_el) for P in nodes}
    cont = {P: {} for P in nodes}  # outcome -> per-election counts (candidate, or "stop")
    for e in range(n_el):
        for b, c in ballots_of(*t.posterior_set_indices(e, n, True, "BB")):
            for P in nodes:
                d = len(P)
                if b[:d] != P:
                    continue
                reach[P][e] += c
                key = "stop" if len(b) == d else b[d]
                cont[P].setdefault(key, np.zeros(n_el))[e] += c
    for P, kind in nodes.items():
        d = len(P)
        K, T = nc - d, d >= mn
        if kind == "init":
            cands, counts, _ = dump[P]
            alpha = {c: counts[i] + prior(d) for i, c in enumerate(cands)}
            if T:
                alpha["stop"] = counts[K] + prior(d)
        else:
            alpha = {c: prior(d) for c in range(nc) if c not in P}
            if T:
                alpha["stop"] = prior(d)
        A = sum(alpha.values())
        assert reach[P].sum() > 200, (P, reach[P].sum())  # the test must have something to look at
        if not T:
            assert "stop" not in cont[P]
        # the two heaviest outcomes and the stop outcome (or the lightest one)
        keys = sorted(alpha, key=lambda k: -alpha[k] if k != "stop" else 1.0)[:2] + (["stop"] if T else [])
        for k in keys:
            z = beta_binomial_z(cont[P].get(k, np.zeros(n_el)), reach[P], alpha[k], A)
            m = len(z)
            assert abs(z.mean()) < 4.5 / np.sqrt(m), (P, kind, k, z.mean(), m)
            # counts are small integers, z^2 is heavy-tailed: the tolerance comes from the sample's own 4th moment
            if m >= 500:
                tol = 4.5 * np.sqrt(max(float(np.mean(z ** 4)) - 1.0, 2.0) / m)
                assert abs(float(np.mean(z * z)) - 1) < tol, (P, kind, k, float(np.mean(z * z)), tol, m)


def test_reduced_c4_shape_matches_the_oracle_in_law(pkg, L, port):
    """20 candidates, 300 observed full ballots, 3000-ballot elections sampled with the tree's lazy subtrees doing most
    of the work (as in BASELINE config 4): 600 elections on each side, statistics compared through per-election means
    with per-election variances (ballots of one election are correlated), |z| < 4.5 each:
    ballot-length histogram, the candidate ranked at depth 4, ballots under the 10 commonest observed 2-prefixes,
    ballots that leave the initialised tree at depth 1 (lazy), distinct ballot types per election."""
    nc, n_obs, n, n_el = 20, 300, 3000, 600
    W = pkg.workloads
    prefs, offsets = W.plackett_luce_ballots(nc, 0.3, n_obs, 4)
    t = pkg.dirichlet_tree(W.candidate_names(nc), 0, nc, 1.0, False, device=0)
    t.update_indices(prefs, offsets)
    o = port.tree(nc, 0, nc, 1.0, False)
    o.update((prefs, offsets))
    obs = oracle.csr_to_ballots(prefs, offsets)
    two = {}
    for b in obs:
        two[b[:2]] = two.get(b[:2], 0) + 1
    top2 = sorted(two, key=two.get, reverse=True)[:10]
    firsts = {b[0] for b in obs}
    feats = [[], []]
    for side in (0, 1):
        for e in range(n_el):
            bs = ballots_of(*(t.posterior_set_indices(e, n, True, "C4LAW") if side == 0 else o.posterior_set(n, True, e)))
            f = np.zeros(20 + 20 + 10 + 2)
            for b, c in bs:
                f[len(b)] += c  # lengths 0..19
                if len(b) > 4:
                    f[20 + b[4]] += c
                if b[:2] in top2:
                    f[40 + top2.index(b[:2])] += c
                if b and b[0] not in firsts:
                    f[50] += c
            f[51] = len(bs)
            assert f[:20].sum() == n
            feats[side].append(f)
    g, r = np.array(feats[0]), np.array(feats[1])
    se = np.sqrt(g.var(0, ddof=1) / n_el + r.var(0, ddof=1) / n_el)
    live = se > 0
    z = (g.mean(0)[live] - r.mean(0)[live]) / se[live]
    assert np.abs(z).max() < 4.5, (np.abs(z).max(), np.argmax(np.abs(z)))
    assert stats.chi2.sf(float((z * z).sum()), int(live.sum())) > 1e-4, float((z * z).sum())
    # the same elections through sample_posterior: win frequencies vs the oracle (a close 20-way race)
    w = t.sample_posterior_counts(4000, n, replace=True, seed="C4LAW") / 4000.0
    freq, _ = o.sample_posterior(4000, n, 1, True, 4, "C4LAW")
    pooled = (w + freq) / 2
    tol = 4.5 * np.sqrt(np.maximum(pooled * (1 - pooled), 1.0 / 4000) * 2 / 4000)
    assert np.all(np.abs(w - freq) <= tol), (w.tolist(), freq.tolist())
