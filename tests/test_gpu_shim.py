"""Runs the Rcpp shim itself on the GPU (SURVEY 8(f)-3 prep).

R and Rcpp are not in this image, so elections.dtree_b200/host/r/R_tree_b200.cpp — the file a maintainer drops into
the R package in place of src/R_tree.cpp, R_social_choice.cpp and RcppModules.cpp — is compiled against the functional
Rcpp stand-in of tests/stubs/Rcpp.h and driven by tests/shim/shim_driver.cpp, which replays the scenarios of
tests/testthat/test-posterior.R, test-update_and_reset.R, test-minmaxdepth.R, test-determinism.R and the Wakehurst
golden vector of test-socialchoice.R with the arguments R/dtree.R would pass. What is exercised is the shim class
RDirichletTree and social_choice_irv, not a Python restatement of them.
"""
import subprocess

import pytest

from test_abi_cpu import build_shim_driver

pytestmark = pytest.mark.gpu


def test_shim_replays_the_reference_testthat_scenarios(tmp_path, wakehurst):
    exe = str(tmp_path / "shim_driver")
    r = build_shim_driver(exe)
    assert r.returncode == 0, r.stderr
    names = [n.replace(" ", "_") for n in wakehurst["candidates"]]  # the driver's file format is blank-separated
    data = tmp_path / "wakehurst.txt"
    with open(data, "w") as f:
        f.write(" ".join(names) + "\n")
        for b, c in zip(wakehurst["ballots"], wakehurst["counts"]):
            f.write("%d %s\n" % (c, " ".join(names[i] for i in b)))
    expected = [n.replace(" ", "_") for n in wakehurst["irv_elimination_order"] + [wakehurst["irv_winner"]]]
    run = subprocess.run([exe, str(data)] + expected,
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
    assert run.stdout.strip().startswith("shim ok:"), run.stdout
    assert int(run.stdout.split()[2]) > 1100  # every replayed assertion ran
