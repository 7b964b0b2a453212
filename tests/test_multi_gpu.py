"""NCCL path on real GPUs (skipped with fewer than two): one process per GPU, contiguous election ranges, one
all-reduce of the win counts — the totals must equal the single-GPU totals bit for bit (north star, part 3)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, ROOT)
    from conftest import load_pkg
    pkg = load_pkg()
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    W = pkg.workloads
    prefs, offsets = W.plackett_luce_ballots(9, 0.1, 500, 33, [1, 1, 1, 1, 1, 1, 1, 1, 1, 4])
    t = pkg.dirichlet_tree(W.candidate_names(9), 0, 9, 1.0, False, device=rank)
    t.update_indices(prefs, offsets)
    probs = pkg.sharding.sample_posterior_sharded(t, 3001, 4000, seed=None if rank else "NCCLSEED")
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array([probs[c] for c in t.candidates]))
    dist.destroy_process_group()


def test_sharded_nccl_totals_equal_single_gpu(pkg, tmp_path):
    import torch
    import torch.multiprocessing as mp
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = min(n, 4)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    W = pkg.workloads
    prefs, offsets = W.plackett_luce_ballots(9, 0.1, 500, 33, [1, 1, 1, 1, 1, 1, 1, 1, 1, 4])
    t = pkg.dirichlet_tree(W.candidate_names(9), 0, 9, 1.0, False, device=0)
    t.update_indices(prefs, offsets)
    want = t.sample_posterior_counts(3001, 4000, seed="NCCLSEED") / 3001.0
    assert want.max() < 1.0  # a race with more than one possible winner
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("r%d.npy" % r)), want)
