"""The algorithm of the device-side tree build (csrc/tree_build.cu), restated with numpy and checked on the
CPU against the host trie (which tests/test_abi_cpu.py pins to the reference's IRVNode::update, irv_node.cpp:176-221):

  1. per ballot, the slot each preference takes under the swap rule (irv_node.cpp:203-220) — a function of the ballot alone;
  2. per depth, one count per ballot still walking (stop slot if the ballot ends there, irv_node.cpp:191-194), then the
     slots that received a count and may have children (irv_node.cpp:210-215) numbered by a prefix sum in slot order.

Numbering in slot order is breadth-first order, so the arrays come out exactly as HostTree::flatten lays them out. The GPU
kernels themselves are compared array for array in tests/test_gpu_tree_build.py; this test keeps the algorithm honest
where there is no GPU.
"""
import numpy as np
import pytest

from oracle import oracle


def slot_sequences(ballots, nc):
    out = []
    for b in ballots:
        path = list(range(nc))
        s = []
        for d, c in enumerate(b[:nc - 1]):
            i = path.index(c, d)
            s.append(i - d)
            path[i], path[d] = path[d], path[i]
        out.append(s)
    return out


def build_level_synchronous(ballots, nc):
    """Returns the dtree_export_nodes stream (depth, prefix, K, candidates, counts, has-child flags) in level order."""
    seqs = slot_sequences(ballots, nc)
    node_base, slot_count, slot_child, slot_cand, prefix = [0, nc + 1], [0] * (nc + 1), [-1] * (nc + 1), list(range(nc)) + [255], [[]]
    level_first, level_nodes = 0, 1
    cursor = [0] * len(ballots)  # slot index taken at the previous depth; None = done
    for depth in range(nc):
        K = nc - depth
        for b, (ballot, s) in enumerate(zip(ballots, seqs)):
            if cursor[b] is None:
                continue
            node = 0 if depth == 0 else slot_child[cursor[b]]
            base = node_base[node]
            if len(ballot) == depth:
                slot_count[base + K] += 1
                cursor[b] = None
            else:
                idx = base + s[depth]
                slot_count[idx] += 1
                cursor[b] = None if K == 2 else idx
        if K <= 2:
            break
        slot_lo = node_base[level_first]
        n_slots_level = level_nodes * (K + 1)
        flags = np.array([(i % (K + 1)) != K and slot_count[slot_lo + i] > 0 for i in range(n_slots_level)], dtype=np.int64)
        ranks = np.cumsum(flags) - flags  # exclusive prefix sum
        n_new = int(flags.sum())
        if n_new == 0:
            break
        first_new, new_slot_base = len(node_base) - 1, len(slot_count)
        node_base.pop()  # the sentinel moves to the end again below
        node_base += [new_slot_base + r * K for r in range(n_new)] + [new_slot_base + n_new * K]
        slot_count += [0] * (n_new * K)
        slot_child += [-1] * (n_new * K)
        slot_cand += [0] * (n_new * K)
        prefix += [None] * n_new
        for i in np.flatnonzero(flags):
            pn, j = divmod(int(i), K + 1)
            pbase = slot_lo + pn * (K + 1)
            child = first_new + int(ranks[i])
            cbase = new_slot_base + int(ranks[i]) * K
            slot_child[slot_lo + int(i)] = child
            for q in range(K - 1):
                slot_cand[cbase + q] = slot_cand[pbase] if q + 1 == j else slot_cand[pbase + q + 1]
            slot_cand[cbase + K - 1] = 255
            prefix[child] = prefix[level_first + pn] + [slot_cand[pbase + j]]
        level_first, level_nodes = first_new, n_new
    out = []
    n_nodes = len(prefix)
    assert len(node_base) == n_nodes + 1 and node_base[-1] == len(slot_count)
    for i in range(n_nodes):
        base, K = node_base[i], node_base[i + 1] - node_base[i] - 1
        p = prefix[i]
        out += [len(p)] + p + [K] + slot_cand[base:base + K] + slot_count[base:base + K + 1]
        out += [1 if slot_child[base + j] >= 0 else 0 for j in range(K)]
    return np.array(out, dtype=np.uint32)


@pytest.mark.parametrize("nc,n,seed", [(2, 30, 1), (3, 100, 2), (4, 400, 3), (6, 1500, 4), (9, 800, 5), (13, 300, 6)])
def test_level_synchronous_build_equals_the_trie(pkg, nc, n, seed):
    rng = np.random.default_rng(seed)
    w = np.exp(-0.4 * np.arange(nc))
    ballots = []
    for _ in range(n):
        ln = int(rng.integers(0, nc + 1))
        g = rng.gumbel(size=nc) + np.log(w)
        ballots.append(tuple(int(c) for c in np.argsort(-g)[:ln]))
    prefs, offsets = oracle.ballots_to_csr(ballots)
    t = pkg.dirichlet_tree(pkg.workloads.candidate_names(nc), 0, nc)
    t.update_indices(prefs, offsets)
    assert oracle.parse_node_dump(build_level_synchronous(ballots, nc)) == oracle.parse_node_dump(t.export_nodes())
