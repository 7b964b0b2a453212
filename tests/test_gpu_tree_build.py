"""GPU parity tests for the device-side tree build (csrc/tree_build.cu), the replacement of
IRVNode::update (irv_node.cpp:176-221) + DirichletTree::update (dirichlet_tree.h:182-193) for the device mirror.

Integer work, so everything is bit-exact:
 * the device-built arrays, read back and rewritten as a node dump, equal the oracle's dump of the reference
   tree after the same updates (prefix, slot order, counts, which children exist);
 * they equal, array for array, the mirror the host trie produces (tuning host_build=1), including the packed
   observed multiset the without-replacement posterior starts from;
 * sampling from either mirror gives identical win counts.
"""
import warnings

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(pkg):
    lib = pkg._native.lib()
    if lib.dtree_device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests cannot run (there is no CPU fallback)")
    return lib


def ragged_ballots(rng, nc, n, p_empty=0.05):
    """Random ballots of every length 0..nc, skewed so that prefixes repeat."""
    w = np.exp(-0.4 * np.arange(nc))
    out = []
    for _ in range(n):
        if rng.random() < p_empty:
            out.append(())
            continue
        ln = int(rng.integers(1, nc + 1))
        g = rng.gumbel(size=nc) + np.log(w)
        out.append(tuple(int(c) for c in np.argsort(-g)[:ln]))
    return out


def image_to_node_dump(img, nc):
    """The dtree_export_nodes stream (depth, prefix, K, candidates, counts, has-child flags) from device arrays."""
    nb, cnt, child, cand = img["node_base"], img["slot_count"], img["slot_child"], img["slot_cand"]
    n = len(nb) - 1
    prefix = {0: []}
    out = []
    for i in range(n):
        base, K = int(nb[i]), int(nb[i + 1] - nb[i] - 1)
        p = prefix.pop(i)
        assert len(p) == nc - K
        out += [len(p)] + p + [K] + [int(c) for c in cand[base:base + K]] + [int(c) for c in cnt[base:base + K + 1]]
        for j in range(K):
            c = int(child[base + j])
            out.append(1 if c >= 0 else 0)
            if c >= 0:
                assert c > i
                prefix[c] = p + [int(cand[base + j])]
        assert cand[base + K] == 0xFF and child[base + K] == -1
    return np.array(out, dtype=np.uint32)


def build_both(pkg, nc, batches, mn=0, mx=None, a0=1.0, vd=False):
    mx = nc if mx is None else mx
    names = pkg.workloads.candidate_names(nc)
    dev = pkg.dirichlet_tree(names, mn, mx, a0, vd, device=0)
    host = pkg.dirichlet_tree(names, mn, mx, a0, vd, device=0).tune(host_build=1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for prefs, offsets in batches:
            dev.update_indices(prefs, offsets)
            host.update_indices(prefs, offsets)
    return dev, host


def assert_same_image(a, b):
    assert a.keys() == b.keys()
    for k in a:
        assert a[k].shape == b[k].shape, k
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("nc,n,seed", [(2, 50, 1), (3, 200, 2), (4, 1000, 3), (7, 5000, 4), (10, 20000, 5), (13, 3000, 6),
                                       (14, 3000, 7), (20, 4000, 8), (26, 2000, 9), (32, 1500, 10)])
def test_device_built_tree_equals_reference_tree(pkg, L, port, nc, n, seed):
    rng = np.random.default_rng(seed)
    ballots = ragged_ballots(rng, nc, n)
    prefs, offsets = oracle.ballots_to_csr(ballots)
    dev, host = build_both(pkg, nc, [(prefs, offsets)])
    o = port.tree(nc, 0, nc, 1.0, False)
    o.update((prefs, offsets))
    img = dev.device_image()
    assert oracle.parse_node_dump(image_to_node_dump(img, nc)) == oracle.parse_node_dump(o.dump_nodes())
    assert_same_image(img, host.device_image())
    # the packed observed multiset: counts add up, first-preference tallies match the ballots
    assert int(img["obs_cnt"].sum()) == n
    firsts = np.bincount([b[0] for b in ballots if b], minlength=32)
    assert np.array_equal(img["obs_tally0"], firsts)
    assert dev.last_stats()["build_launches"] == 1  # the whole build is one cooperative kernel
    # the same tree through the other two device build paths: a launch per phase, and exact allocation with the
    # host reading every level's size (what trees too large for a-priori-bound allocation take on their own)
    for mode, least in ((1, 4), (2, 4)):
        dev.tune(build_mode=mode)
        assert_same_image(img, dev.device_image())
        assert dev.last_stats()["build_launches"] >= least
    dev.tune(build_mode=0)
    o.close()


def test_incremental_updates_reset_and_empty_tree(pkg, L, port):
    nc = 6
    rng = np.random.default_rng(11)
    dev, host = build_both(pkg, nc, [])
    o = port.tree(nc, 0, nc, 1.0, False)
    assert_same_image(dev.device_image(), host.device_image())  # root only
    assert oracle.parse_node_dump(image_to_node_dump(dev.device_image(), nc)) == oracle.parse_node_dump(o.dump_nodes())
    for step, n in enumerate([1, 7, 300, 2, 1500]):
        prefs, offsets = oracle.ballots_to_csr(ragged_ballots(rng, nc, n))
        dev.update_indices(prefs, offsets)
        host.update_indices(prefs, offsets)
        o.update((prefs, offsets))
        img = dev.device_image()
        assert oracle.parse_node_dump(image_to_node_dump(img, nc)) == oracle.parse_node_dump(o.dump_nodes()), step
        assert_same_image(img, host.device_image())
        # the host-side exports (lazily synchronised trie) still agree
        assert oracle.parse_node_dump(dev.export_nodes()) == oracle.parse_node_dump(o.dump_nodes())
    dev.reset()
    host.reset()
    o.reset()
    assert_same_image(dev.device_image(), host.device_image())
    assert len(dev.device_image()["node_base"]) == 2
    prefs, offsets = oracle.ballots_to_csr(ragged_ballots(rng, nc, 40))
    dev.update_indices(prefs, offsets)
    o.update((prefs, offsets))
    assert oracle.parse_node_dump(image_to_node_dump(dev.device_image(), nc)) == oracle.parse_node_dump(o.dump_nodes())
    o.close()


def test_log_of_equal_length_ballots_needs_no_offsets_upload(pkg, L, port):
    """While every ballot of the log has the same number of preferences the device writes the offsets itself
    (8 bytes per ballot less to upload); the first ballot of another length switches back to uploading them. Same tree
    either way, also after dtree_invalidate_device (the whole log from host memory again) and after a reset."""
    from elections_dtree_b200 import _native as N
    nc = 7
    rng = np.random.default_rng(23)
    dev, host = build_both(pkg, nc, [])
    o = port.tree(nc, 0, nc, 1.0, False)

    def full(n):
        w = np.exp(-0.5 * np.arange(nc))
        return [tuple(int(c) for c in np.argsort(-(rng.gumbel(size=nc) + np.log(w)))) for _ in range(n)]

    for step, ballots in enumerate([full(500), full(300), ragged_ballots(rng, nc, 200), full(100)]):
        prefs, offsets = oracle.ballots_to_csr(ballots)
        dev.update_indices(prefs, offsets)
        host.update_indices(prefs, offsets)
        o.update((prefs, offsets))
        img = dev.device_image()
        assert oracle.parse_node_dump(image_to_node_dump(img, nc)) == oracle.parse_node_dump(o.dump_nodes()), step
        assert_same_image(img, host.device_image())
        N.check(L.dtree_invalidate_device(dev._h))
        assert_same_image(dev.device_image(), img)
    dev.reset()
    host.reset()
    prefs, offsets = oracle.ballots_to_csr(full(64))
    dev.update_indices(prefs, offsets)
    host.update_indices(prefs, offsets)
    assert_same_image(dev.device_image(), host.device_image())
    o.close()


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4", "C5"])
def test_benchmark_trees_match_the_host_trie_and_sample_identically(pkg, L, name):
    W = pkg.workloads
    cfg = W.CONFIGS[name]
    prefs, offsets = W.observed_ballots(name)
    nc = cfg["n_candidates"]
    dev, host = build_both(pkg, nc, [(prefs, offsets)], cfg["min_depth"], cfg["max_depth"], cfg["a0"], cfg["vd"])
    img = dev.device_image()
    assert_same_image(img, host.device_image())
    assert np.array_equal(image_to_node_dump(img, nc), dev.export_nodes())
    n_el = 64 if name in ("C3", "C4") else 512
    for replace in (False, True):
        a = dev.sample_posterior_counts(n_el, cfg["n_ballots"], replace=replace, seed="BUILDSEED0")
        b = host.sample_posterior_counts(n_el, cfg["n_ballots"], replace=replace, seed="BUILDSEED0")
        assert np.array_equal(a, b) and int(np.sum(a)) == n_el


def test_invalidate_rebuilds_from_host_memory(pkg, L):
    nc = 5
    rng = np.random.default_rng(12)
    prefs, offsets = oracle.ballots_to_csr(ragged_ballots(rng, nc, 500))
    dev, host = build_both(pkg, nc, [(prefs, offsets)])
    first = dev.device_image()
    h2d0 = dev.last_stats()["h2d_bytes"]
    L.dtree_invalidate_device(dev._h)
    assert_same_image(first, dev.device_image())
    assert dev.last_stats()["h2d_bytes"] > h2d0  # the log travelled again
    assert_same_image(first, host.device_image())
