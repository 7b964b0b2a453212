"""The decision rule of the deferred kernels, checked on a small CPU model: instant-runoff rounds may be
decided while nodes of the ballot tree are still unsplit, as long as the ballots inside them cannot change which
candidate is weakest. The model keeps what makes the rule exact on the GPU — a node's split is a pure function of
(seed, node prefix, count) — and compares three evaluations of the same elections:

  eager      split everything an elimination releases before deciding the next round (irv_ballot.cpp:86-133 on a
             lazily expanded tree; what deferred_kernel did before, and what simulate_kernel's materialisation implies)
  heaviest   lane_kernel's order: split the waiting node with the most ballots until the round is implied
  threshold  deferred_kernel's round-1 order: split every waiting node holding at least need / (2 n_waiting) + 1 ballots
  histogram  deferred_kernel's order now: waiting ballots are counted per weight class (2^b <= ballots < 2^(b+1));
             every node of weight >= 2^b is split for the largest b whose lighter classes together hold at most 3/4 of
             the margin the round is missing (with nothing to spare, everything); single-ballot nodes wait apart and
             are touched only then

All of them must give the same elimination order in every election, including ties (broken by a per-round draw that is
a pure function of the seed), near-ties and elections with a handful of ballots. No GPU, no library: host logic only.
"""
import hashlib

import numpy as np
import pytest


def node_rng(seed, prefix, tag=b"split"):
    h = hashlib.blake2b(tag + bytes([255]) + bytes(prefix), key=seed.to_bytes(8, "little"), digest_size=8).digest()
    return np.random.default_rng(int.from_bytes(h, "little"))


def split(seed, nc, max_depth, prefix, count, skew):
    """Children of a node: {next candidate: count}, the rest stop here. Pure function of its arguments."""
    rest = [c for c in range(nc) if c not in prefix]
    if len(prefix) >= max_depth or not rest or count == 0:
        return {}
    rng = node_rng(seed, prefix)
    alpha = np.array([skew ** c + 0.05 for c in rest] + [0.3])
    p = rng.dirichlet(alpha)
    k = rng.multinomial(count, p)
    return {c: int(n) for c, n in zip(rest, k[:-1]) if n}


class Election:
    def __init__(self, seed, nc, max_depth, n_ballots, skew, policy):
        self.seed, self.nc, self.max_depth, self.skew, self.policy = seed, nc, max_depth, skew, policy
        self.tally = [0] * nc
        self.eliminated = set()
        self.parked = {c: [] for c in range(nc)}  # candidate -> [(prefix, count)]
        self.waiting = []  # [(prefix, count)] whose last candidate is out
        self.splits = 0
        self._split((), n_ballots)

    def _split(self, prefix, count):
        self.splits += 1
        for c, n in split(self.seed, self.nc, self.max_depth, prefix, count, self.skew).items():
            node = (prefix + (c,), n)
            if c in self.eliminated:
                self.waiting.append(node)
            else:
                self.tally[c] += n
                self.parked[c].append(node)

    def _resolve(self, need):
        if self.policy == "eager" or need is None:
            todo, self.waiting = self.waiting, []
        elif self.policy == "heaviest":
            i = max(range(len(self.waiting)), key=lambda j: self.waiting[j][1])
            todo = [self.waiting.pop(i)]
        elif self.policy == "threshold":
            tau = need // (2 * max(len(self.waiting), 1)) + 1
            todo = [w for w in self.waiting if w[1] >= tau]
            self.waiting = [w for w in self.waiting if w[1] < tau]
            assert todo
        else:  # histogram (deferred.cuh: hist[], tau_quarters = 3)
            hist = [0] * 32
            for _, n in self.waiting:
                hist[n.bit_length() - 1] += n
            allow, below, cls = (need * 3) >> 2, 0, 0
            for b in range(32):
                if b == 0 or below <= allow:
                    cls = b
                below += hist[b]
            tau = 1 if cls == 0 else 1 << cls
            todo = [w for w in self.waiting if w[1] >= tau]
            self.waiting = [w for w in self.waiting if w[1] < tau]
            assert todo
        for prefix, count in todo:
            self._split(prefix, count)

    def run(self):
        order = []
        for rnd in range(self.nc - 1):
            while True:
                standing = [c for c in range(self.nc) if c not in self.eliminated]
                u = sum(n for _, n in self.waiting)
                t = sorted(self.tally[c] for c in standing)
                mn, second = t[0], t[1]
                if self.policy == "eager":
                    if not self.waiting:
                        break
                    self._resolve(None)
                    continue
                if u == 0 or second - mn > u:
                    break
                self._resolve(second - mn)
            tied = [c for c in standing if self.tally[c] == mn]
            pick = tied[int(node_rng(self.seed, (rnd,), b"tie").integers(len(tied)))] if len(tied) > 1 else tied[0]
            order.append(pick)
            self.eliminated.add(pick)
            self.waiting += self.parked[pick]
            self.parked[pick] = []
        order.append(next(c for c in range(self.nc) if c not in self.eliminated))
        return order


@pytest.mark.parametrize("nc,max_depth,n_ballots,skew", [(4, 4, 40, 1.0), (5, 5, 12, 1.0), (6, 6, 100_000, 0.6),
                                                         (6, 3, 500, 1.0), (8, 8, 1_000_000, 0.8), (7, 7, 300, 0.9),
                                                         (10, 10, 1_000_000, 0.5), (9, 9, 5_000, 1.0)])
def test_lazy_rounds_give_the_eager_elimination_order(nc, max_depth, n_ballots, skew):
    saved = {"heaviest": 0, "threshold": 0, "histogram": 0}
    eager_splits = 0
    for seed in range(60):
        want = Election(seed, nc, max_depth, n_ballots, skew, "eager")
        order = want.run()
        assert sorted(order) == list(range(nc))
        eager_splits += want.splits
        for policy in ("heaviest", "threshold", "histogram"):
            e = Election(seed, nc, max_depth, n_ballots, skew, policy)
            assert e.run() == order, (seed, policy)
            assert e.splits <= want.splits
            saved[policy] += want.splits - e.splits
    if n_ballots >= 100_000 and skew < 1.0:  # lopsided elections: most of the tree never has to be looked at
        assert saved["heaviest"] > 0.5 * eager_splits and saved["threshold"] > 0.3 * eager_splits, (saved, eager_splits)
        assert saved["histogram"] >= saved["threshold"], saved  # the histogram rule never splits more than the old one here
