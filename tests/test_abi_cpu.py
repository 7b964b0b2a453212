"""CPU-side checks of the product library: it loads, exports every symbol the headers declare, builds
node counts bit-exactly like the reference's IRVNode::update, and refuses to sample without a GPU.
No compute kernels run here."""
import ctypes as C
import os
import warnings

import numpy as np
import pytest

from oracle import oracle


def test_library_exports_every_declared_symbol(pkg):
    L = pkg._native.lib()
    declared = pkg._native.declared_symbols()
    assert len(declared) >= 35
    assert [s for s in declared if not hasattr(L, s)] == []
    assert set(declared) == set(pkg._native._SIGS)
    assert b"sm_100a" in L.dtree_version()


def test_library_is_sm100a_only(pkg):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "--list-elf", pkg._native.LIB_PATH], capture_output=True, text=True).stdout
    archs = {tok for line in out.splitlines() for tok in line.replace(".", " ").split() if tok.startswith("sm_")}
    assert archs == {"sm_100a"}, archs


def test_philox_known_answers_on_host(pkg):
    """Random123 kat_vectors for philox4x32-10."""
    L = pkg._native.lib()
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kats:
        out = (C.c_uint32 * 4)()
        assert L.dtree_debug_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out, 0) == 0
        assert list(out) == want


def test_seed_hash_is_a_function_of_the_string(pkg):
    L = pkg._native.lib()
    a = L.dtree_debug_hash_seed(b"ABCDEFGHIJ", 10)
    assert a == L.dtree_debug_hash_seed(b"ABCDEFGHIJ", 10)
    assert a != L.dtree_debug_hash_seed(b"ABCDEFGHIK", 10)
    assert a != L.dtree_debug_hash_seed(b"ABCDEFGHI", 9)


def _trees(pkg, port, nc, mn, mx, a0=1.0, vd=False):
    return pkg.dirichlet_tree(pkg.workloads.candidate_names(nc), mn, mx, a0, vd), port.tree(nc, mn, mx, a0, vd)


@pytest.mark.parametrize("nc,mn,mx", [(2, 0, 1), (3, 0, 3), (4, 0, 3), (6, 0, 5), (7, 2, 7), (10, 0, 10), (14, 0, 6),
                                      (26, 0, 26), (32, 3, 32)])
def test_update_counts_bit_exact_vs_oracle(pkg, port, nc, mn, mx):
    """IRVNode::update (irv_node.cpp:176-221): same nodes, same slot order, same counts."""
    rng = np.random.default_rng(nc)
    pmf = rng.random(nc + 1) + 0.02
    t, o = _trees(pkg, port, nc, mn, mx)
    total = 0
    for batch in range(3):
        n = [300, 1, 700][batch]
        prefs, offsets = pkg.workloads.plackett_luce_ballots(nc, 0.3, n, 100 * nc + batch, pmf)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t.update_indices(prefs, offsets)
        o.update((prefs, offsets))
        total += n
        assert oracle.parse_node_dump(t.export_nodes()) == oracle.parse_node_dump(o.dump_nodes())
        assert t.n_observed == o.n_observed() == total
    # the observed multiset derived from the counts == the reference's std::map, up to the
    # (n candidates) == (n-1 candidates) identification
    def canon(prefs, offsets, counts):
        d = {}
        for b, c in zip(oracle.csr_to_ballots(prefs, offsets), counts):
            b = b[:nc - 1]
            d[b] = d.get(b, 0) + int(c)
        return d
    assert canon(*t.observed_indices()) == canon(*o.observed())
    t.reset()
    o.reset()
    assert oracle.parse_node_dump(t.export_nodes()) == oracle.parse_node_dump(o.dump_nodes())
    assert t.n_observed == 0


def test_update_log_accepts_offsets_that_do_not_start_at_zero_and_rejects_atomically(pkg, port):
    """dtree_update appends to the ballot log: a CSR window into a larger array (offsets[0] != 0) is rebased, several
    updates accumulate, a rejected batch leaves no trace, reset empties the log (R_tree.cpp:125-153, :155-...)."""
    nc = 6
    prefs, offsets = pkg.workloads.plackett_luce_ballots(nc, 0.3, 400, 9, np.ones(nc + 1))
    t, o = _trees(pkg, port, nc, 0, nc)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t.update_indices(prefs, offsets[:101])    # ballots 0..99
        t.update_indices(prefs, offsets[100:251])  # ballots 100..249 as a window: offsets[0] > 0
        bad = prefs.copy()
        bad[int(offsets[300])] = nc  # unknown candidate in ballot 300
        with pytest.raises(pkg._native.DtreeError, match="Unknown candidate"):
            t.update_indices(bad, offsets[250:])
        assert t.n_observed == 250
        t.update_indices(prefs, offsets[250:])
    o.update((prefs, offsets))
    assert t.n_observed == o.n_observed() == 400
    assert oracle.parse_node_dump(t.export_nodes()) == oracle.parse_node_dump(o.dump_nodes())
    t.reset()
    o.reset()
    assert t.n_observed == 0
    assert oracle.parse_node_dump(t.export_nodes()) == oracle.parse_node_dump(o.dump_nodes())
    o.close()


def test_update_matches_recorded_reference_dumps(pkg, ref_vectors):
    """tests/golden/ref_vectors.json was written by the unmodified reference."""
    for case in ref_vectors:
        nc = case["n_candidates"]
        t = pkg.dirichlet_tree(pkg.workloads.candidate_names(nc), case["min_depth"], case["max_depth"], case["a0"],
                               case["vd"])
        names = t.candidates
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t.update([[names[c] for c in b] for b in case["observed"]])
        assert oracle.parse_node_dump(t.export_nodes()) == oracle.parse_node_dump(case["node_dump"])
        assert t.depth_factors().tolist() == case["depth_factors"]


def test_depth_factors_and_stale_quirk(pkg):
    """irv_node.cpp:13-25 and irv_node.h:127 (set_min_depth leaves the factors stale)."""
    mk = lambda mn, mx: pkg.dirichlet_tree(list("ABCD"), mn, mx)
    assert mk(0, 3).depth_factors().tolist() == [12, 3, 1]
    assert mk(2, 3).depth_factors().tolist() == [9, 3, 1]
    assert mk(3, 3).depth_factors().tolist() == [6, 2, 1]
    assert mk(0, 4).depth_factors().tolist() == [24, 6, 2, 1]
    t = mk(0, 3)
    t.min_depth = 3
    assert t.depth_factors().tolist() == [12, 3, 1]
    t.max_depth = 3
    assert t.depth_factors().tolist() == [6, 2, 1]


def test_constructor_validation(pkg):
    """tests/testthat/test-dtree_constructor.R and test-minmaxdepth.R:61-76"""
    with pytest.raises(ValueError):
        pkg.dirtree(list("ABCDEFGHIJKLMNOPQRSTUVWXYZ"), min_depth=4, max_depth=3)
    with pytest.raises(ValueError):
        pkg.dirtree(["A", "A", "B"])
    with pytest.raises(TypeError):
        pkg.dirtree([1, 2, 3])
    with pytest.raises(ValueError):
        pkg.dirtree(list("ABC"), a0=-1)
    with pytest.raises(TypeError):
        pkg.dirtree(list("ABC"), vd="yes")
    with pytest.raises(ValueError):
        pkg.dirtree(list("ABC"), max_depth=4)
    t = pkg.dirtree(list("ABCDEFGH"), min_depth=3)
    with pytest.raises(ValueError):
        t.max_depth = 2
    t = pkg.dirtree(list("ABCDEFGH"), max_depth=3)
    with pytest.raises(ValueError):
        t.min_depth = 4
    assert pkg.dirtree(list("ABCD")).max_depth == 4 and pkg.dirichlet_tree(list("ABCD")).max_depth == 3


def test_changing_parameters(pkg):
    """tests/testthat/test-changing_parameters.R"""
    t = pkg.dirtree(list("ABCDE"), a0=1.0, min_depth=1, max_depth=4, vd=False)
    t.a0, t.min_depth, t.max_depth, t.vd = 2.5, 2, 3, True
    assert (t.a0, t.min_depth, t.max_depth, t.vd, t.n_candidates) == (2.5, 2, 3, True, 5)
    assert t.candidates == list("ABCDE")
    with pytest.raises(ValueError):
        t.a0 = -0.1
    with pytest.raises(TypeError):
        t.vd = 1


def test_update_validation_and_warnings(pkg):
    """tests/testthat/test-update_and_reset.R:32-83; R_tree.cpp:30-31,139-147,95-105"""
    t = pkg.dirtree(list("ABCD"), min_depth=2)
    with pytest.warns(UserWarning, match="fewer than `minDepth`"):
        pkg.update(t, [["A"]])
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        pkg.update(t, [[], ["A", "B"], ["A", "B", "C", "D"]])  # blank and long-enough ballots do not warn
    with pytest.raises(ValueError, match="Unknown candidate"):
        pkg.update(t, [["A", "Z"]])
    assert t.n_observed == 4
    with pytest.raises(pkg._native.DtreeError):  # repeated candidate: rejected, nothing changes
        t.update_indices(np.array([0, 0], dtype=np.uint8), np.array([0, 2], dtype=np.uint64))
    assert t.n_observed == 4
    t2 = pkg.dirtree(list("ABCD"))
    pkg.update(t2, [["A"]])
    with pytest.warns(UserWarning, match="have been observed"):
        t2.min_depth = 2
    t2.update([["A", "B"]], counts=[5])
    assert t2.n_observed == 6


def test_sampling_argument_validation(pkg):
    """tests/testthat/test-posterior.R:31-69"""
    t = pkg.dirtree(list("ABCDEFGHIJ"))
    pkg.update(t, [list("ABCDEFGHIJ"), list("JIHGFEDCBA")])
    with pytest.raises(ValueError, match="n_ballots"):
        pkg.sample_posterior(t, 1, 1)
    with pytest.raises(ValueError, match="n_elections"):
        pkg.sample_posterior(pkg.dirtree(list("ABC")), 0, 1)
    with pytest.raises(ValueError, match="n_threads"):
        pkg.sample_posterior(t, 1, 10, n_threads=0)
    with pytest.raises(ValueError):
        pkg.sample_predictive(t, 0)
    with pytest.raises(NotImplementedError):
        pkg.social_choice([["A"]], "stv")
    assert pkg.social_choice([["A", "B"], ["B"], ["A"]], "plurality") == ["A"]


def test_no_cpu_fallback(pkg):
    """Without a usable GPU every sampling / social-choice entry point must fail, not fall back."""
    if pkg._native.lib().dtree_device_count() > 0:
        pytest.skip("a GPU is visible")
    t = pkg.dirtree(list("ABCD"))
    for call in (lambda: pkg.sample_posterior(t, 10, 10), lambda: pkg.sample_predictive(t, 10),
                 lambda: pkg.social_choice([["A", "B"], ["B"]], "irv"),
                 lambda: t.posterior_set_indices(0, 10)):
        with pytest.raises(pkg._native.DtreeError) as ei:
            call()
        assert ei.value.code == pkg._native.ERR_CUDA and "no CPU path" in str(ei.value)


def test_product_sources_never_touch_the_oracle(pkg):
    """The checker is test infrastructure: nothing under the package may import, link or open it."""
    root = os.path.dirname(pkg._native.LIB_PATH)
    pkgdir = os.path.dirname(root)
    for dp, _, files in os.walk(pkgdir):
        if "build" in dp.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in src.replace("checked against the CPU oracle", ""), os.path.join(dp, f)


def build_shim_driver(out_path):
    """Compiles tests/shim/shim_driver.cpp — the Rcpp shim itself plus the replay of the reference's testthat
    scenarios — against tests/stubs/Rcpp.h and links it with lib/libdtree_b200.so. Returns the g++ result."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    lib_dir = os.path.join(root, "elections.dtree_b200", "lib")
    return subprocess.run([gxx, "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(root, "tests", "stubs"),
                           "-I" + os.path.join(root, "include"), os.path.join(root, "tests", "shim", "shim_driver.cpp"),
                           "-o", out_path, "-L" + lib_dir, "-ldtree_b200", "-Wl,-rpath," + lib_dir],
                          capture_output=True, text=True)


def test_rcpp_shim_compiles_and_links_against_the_c_abi(tmp_path):
    """INTEGRATION.md's Rcpp layer (elections.dtree_b200/host/r/R_tree_b200.cpp) compiles against
    include/dtree_b200.h with the functional Rcpp stand-in of tests/stubs/ and links with the library (R is not
    installed in this image). Without a GPU the driver must refuse to sample: there is no CPU path behind the shim
    either. tests/test_gpu_shim.py runs it on the GPU box."""
    import subprocess
    exe = str(tmp_path / "shim_driver")
    r = build_shim_driver(exe)
    assert r.returncode == 0, r.stderr
    import elections_dtree_b200._native as N
    if N.lib().dtree_device_count() == 0:
        run = subprocess.run([exe], capture_output=True, text=True)
        assert run.returncode != 0 and "no CPU path" in run.stderr


def test_update_property_random_ballot_streams(pkg, port):
    """Property test (hypothesis): any stream of valid ballots — blank, partial, full, repeated, in any batching —
    leaves the host tree with exactly the reference's nodes, slot order and counts (irv_node.cpp:176-221)."""
    from hypothesis import given, settings, strategies as st

    @st.composite
    def streams(draw):
        nc = draw(st.integers(2, 9))
        ballot = st.lists(st.integers(0, nc - 1), unique=True, max_size=nc)
        batches = draw(st.lists(st.lists(ballot, max_size=12), min_size=1, max_size=4))
        return nc, draw(st.integers(0, nc)), batches

    @settings(max_examples=60, deadline=None)
    @given(streams())
    def check(args):
        nc, mx, batches = args
        t = pkg.dirichlet_tree(pkg.workloads.candidate_names(nc), 0, mx, 1.0, False)
        o = port.tree(nc, 0, mx, 1.0, False)
        for batch in batches:
            if not batch:
                continue
            prefs, offsets = oracle.ballots_to_csr(batch)
            t.update_indices(prefs, offsets)
            o.update((prefs, offsets))
        assert oracle.parse_node_dump(t.export_nodes()) == oracle.parse_node_dump(o.dump_nodes())
        assert t.n_observed == o.n_observed()
        o.close()

    check()


def test_lane_kernel_launch_shape(pkg):
    """csrc/launch_shape.h: a thread-per-election job is launched as ONE block per SM of as many warps as it needs to run
    in a single pass, never more warps than an SM holds or than have scratch, and the grid always covers the job or
    fills the machine (the warps then claim batch after batch)."""
    import ctypes as C
    L = pkg._native.lib()

    def shape(batches, sms=148, wps=24, fit=10**9, forced=0):
        w, g = C.c_uint32(0), C.c_uint32(0)
        assert L.dtree_debug_lane_shape(batches, sms, wps, fit, forced, C.byref(w), C.byref(g)) == 0
        return w.value, g.value

    # the benchmark job: 100 000 elections = 3 125 batches -> 22 warps per block, 143 blocks (one per SM, one pass)
    assert shape(3125) == (22, 143)
    # 12 500 elections: 3 warps per block on 131 SMs; a tiny job: single-warp blocks
    assert shape(391) == (3, 131)
    assert shape(5) == (1, 5)
    # larger than one pass: full blocks on every SM
    assert shape(31250) == (24, 148)
    rng = np.random.default_rng(5)
    for _ in range(2000):
        batches = int(rng.integers(1, 200000))
        sms = int(rng.integers(1, 200))
        wps = int(rng.integers(1, 25))
        fit = int(rng.integers(1, 6000))
        forced = int(rng.choice([0, 0, 0, int(rng.integers(1, 25))]))
        w, g = shape(batches, sms, wps, fit, forced)
        assert 1 <= w <= wps and w <= fit and g >= 1
        assert w * g <= max(fit, w)                      # never more warps than have scratch
        assert g <= sms * (wps // w)                     # never more blocks than are resident
        covers = w * g >= batches
        full = g == min(sms * (wps // w), max(1, fit // w))
        assert covers or full                            # one pass, or the machine (or the memory budget) is full
        if forced == 0 and covers and fit >= sms * wps:
            assert g <= sms                              # one block per SM when the job fits one pass
