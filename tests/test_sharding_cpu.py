"""world_size-2 (and 3) gloo runs of the N>1 host logic on CPU: the rank partition covers every election
exactly once, the seed is agreed on, and the all-reduced win counts equal the single-process totals.
The per-rank compute is injected (a deterministic stand-in keyed on the global election id), because
the product's compute exists only on the GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_winner(election, nc):
    """Deterministic stand-in for "who wins global election e" (same in every process, no hashing)."""
    return ((int(election) * 2654435761) >> 7) % nc


def _worker(rank, world, port, n_elections, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, ROOT)
    from conftest import load_pkg
    pkg = load_pkg()
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    t = pkg.dirtree(list("ABCDE"))
    seen = []

    def simulate(first, n):
        seen.append((first, n))
        w = np.zeros(5, dtype=np.uint64)
        for e in range(first, first + n):
            w[_fake_winner(e, 5)] += 1
        return w

    probs = pkg.sharding.sample_posterior_sharded(t, n_elections, 10, seed=None if rank else "FIXED", simulate=simulate)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array([probs[c] for c in "ABCDE"] + list(seen[0])))
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n_elections", [(2, 101), (2, 1), (3, 50)])
def test_sharded_totals_equal_single_process(tmp_path, world, n_elections):
    mp.spawn(_worker, args=(world, _free_port(), n_elections, str(tmp_path)), nprocs=world, join=True)
    rows = [np.load(tmp_path / ("r%d.npy" % r)) for r in range(world)]
    want = np.zeros(5)
    for e in range(n_elections):
        want[_fake_winner(e, 5)] += 1
    want /= n_elections
    covered = []
    for r in rows:
        assert np.allclose(r[:5], want)
        covered += list(range(int(r[5]), int(r[5] + r[6])))
    assert sorted(covered) == list(range(n_elections))


def test_election_range_partition():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_pkg
    sh = load_pkg().sharding
    for n in (0, 1, 7, 8, 100_000, 12_345):
        for g in (1, 2, 4, 8):
            ranges = [sh.election_range(n, r, g) for r in range(g)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert max(h - l for l, h in ranges) == -(-n // g)
