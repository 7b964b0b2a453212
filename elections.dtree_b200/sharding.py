"""One process per GPU: contiguous election ranges per rank and the only collective on the path.

Elections are independent given the tree (the reference already splits them over threads without
communication, R_tree.cpp:200-242), and every draw is a function of (seed, GLOBAL election id), so
rank r of G simulates ids [r*ceil(n/G), (r+1)*ceil(n/G)) on its own GPU with the tree replicated,
and the per-candidate win counts (<= 32 uint64) are summed with one all-reduce (NCCL over
NVLink/NVSwitch; 256 bytes, latency-bound). Totals are bit-identical for any G.
"""
import numpy as np


def election_range(n_elections, rank, world_size):
    """[lo, hi) of global election ids owned by `rank`."""
    per = -(-int(n_elections) // int(world_size))
    lo = min(int(n_elections), rank * per)
    return lo, min(int(n_elections), lo + per)


def sample_posterior_sharded(dtree, n_elections, n_ballots, n_winners=1, replace=False, seed=None, group=None,
                             simulate=None):
    """sample_posterior over all ranks of `group`; every rank returns the same {candidate: probability}.

    simulate(first_election, n) -> uint64[n_candidates] replaces the CUDA call; it exists so that the
    partition + reduction logic can be exercised on CPU-only hosts (gloo tests), never as a fallback.
    """
    import torch
    import torch.distributed as dist

    from . import _native as N
    from .api import gseed

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [gseed() if seed is None else seed]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    seed = box[0]
    lo, hi = election_range(n_elections, rank, world)
    nc = dtree.n_candidates
    if simulate is not None:
        wins = torch.from_numpy(np.asarray(simulate(lo, hi - lo), dtype=np.uint64).astype(np.int64))
    else:
        dev = torch.device("cuda", torch.cuda.current_device())
        wins = torch.zeros(nc, dtype=torch.int64, device=dev)
        if hi > lo:
            s = seed.encode()
            stream = torch.cuda.current_stream(dev).cuda_stream
            N.check(N.lib().dtree_sample_posterior_device(dtree._h, lo, hi - lo, int(n_ballots), int(n_winners),
                                                          int(replace), s, len(s), wins.data_ptr(), stream))
            N.check(N.lib().dtree_sample_posterior_finish(dtree._h, stream))
    dist.all_reduce(wins, op=dist.ReduceOp.SUM, group=group)
    total = wins.cpu().numpy().astype(np.float64)
    return {c: total[i] / float(n_elections) for i, c in enumerate(dtree.candidates)}
