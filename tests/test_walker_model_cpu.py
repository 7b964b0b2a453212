"""CPU model of the walkers' stacks (csrc/simulate.cuh, walk_small_items): frames wait in three LIFO stacks by ballots
held (1 / 2-3 / 4-8); the single-ballot stack holds B frames, the other two share B / 2 slots. The kernel claims that
they cannot overflow because (a) a frame holds at least 1 / 2 / 4 ballots, (b) children never hold more ballots than
their parent and (c) frames are drawn from the queue only while the ballots held stay within the budget B. This test
replays the kernel's scheduling rule on random subtrees and checks those bounds at every step, and that every frame of
the queue is eventually walked to its end."""
import numpy as np
import pytest


def bucket(c):
    return 0 if c <= 1 else 1 if c <= 3 else 2


def split(rng, count, depth, max_depth):
    """Children of a frame: its ballots spread over random outcomes; some stop here. Returns child counts."""
    if depth + 1 >= max_depth:
        return []  # leaf children: finished ballots
    stay = rng.binomial(count, 0.85)  # the rest stop at this node
    if stay == 0:
        return []
    k = int(rng.integers(1, 6))
    parts = rng.multinomial(stay, np.ones(k) / k)
    return [int(p) for p in parts if p]


@pytest.mark.parametrize("B,urn_max,seed", [(128, 8, 1), (80, 8, 2), (168, 8, 3), (128, 3, 4), (32, 8, 5)])
def test_walker_stacks_never_overflow(B, urn_max, seed):
    rng = np.random.default_rng(seed)
    queue = [(int(rng.integers(1, urn_max + 1)), int(rng.integers(0, 6))) for _ in range(4000)]
    n_queue_ballots = sum(c for c, _ in queue)
    stacks = [[], [], []]
    ballots = 0
    pend = []          # claimed from the queue, not yet taken in (at most 32)
    q_at = 0
    finished = 0       # ballots that reached the end of their walk
    steps = 0

    def check():
        assert len(stacks[0]) <= B
        assert len(stacks[1]) + len(stacks[2]) <= B // 2
        assert ballots == sum(c for s in stacks for c, _ in s) <= B

    while True:
        steps += 1
        assert steps < 10 ** 6
        b = 2 if len(stacks[2]) >= 32 else 1 if len(stacks[1]) >= 32 else 0 if len(stacks[0]) >= 32 else -1
        if b < 0:
            if not pend and q_at < len(queue):
                pend = queue[q_at:q_at + 32]
                q_at += len(pend)
            took = 0
            if pend:
                total = sum(c for c, _ in pend)
                if ballots + total <= B:
                    took = len(pend)
                else:  # the longest prefix that fits
                    acc = 0
                    for c, _ in pend:
                        if ballots + acc + c > B:
                            break
                        acc += c
                        took += 1
                for c, d in pend[:took]:
                    stacks[bucket(c)].append((c, d))
                    ballots += c
                pend = pend[took:]
                check()
                if took:
                    continue
            b = 2 if stacks[2] else 1 if stacks[1] else 0 if stacks[0] else -1
            if b < 0:
                if not pend and q_at >= len(queue):
                    break
                pytest.fail("empty stacks must have room for a frame")
        take = min(32, len(stacks[b]))
        frames = [stacks[b].pop() for _ in range(take)]
        ballots -= sum(c for c, _ in frames)
        for c, d in frames:
            kids = split(rng, c, d, 9)
            assert sum(kids) <= c
            finished += c - sum(kids)
            for k in kids:
                stacks[bucket(k)].append((k, d + 1))
                ballots += k
        check()
    assert finished == n_queue_ballots and ballots == 0
