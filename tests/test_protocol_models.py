"""Model checks of the two cross-GPU synchronisation protocols: the push ("flag-in-data") gradient exchange of csrc/sgd.cu
(`allreduce_sgd_push_kernel`, first part) and the one-slot flag barrier of csrc/common.cuh (second part).

The kernel's safety argument (SURVEY §7.4 hard part #1: no reset races across back-to-back steps) is: lines carry
epoch = step + 1, inboxes are double-buffered by step parity, and a peer can only overwrite parity p two steps later --
which needs my push of the step in between, which I issue only after I finished reading parity p.  Here that argument is
checked mechanically on an abstract machine: every store of an 8-byte half line and every poll is one atomic action,
ranks interleave arbitrarily (exhaustively for the small configuration, randomly for larger ones), 16-byte lines may
tear between their halves, and a reader accepts a line only when both halves carry the expected epoch -- exactly the
kernel's rule.  Checked: every value a rank consumes is the sender's value of THAT step, nobody deadlocks.
The same machine with a single-buffered inbox is shown to fail, so the model can see the bug the parity prevents.
"""
import random

import pytest


class Rank:
    def __init__(self, r, world, steps):
        self.r, self.world, self.steps = r, world, steps
        self.step = 0
        self.todo = []           # pending atomic actions of the current step
        self.reading = None      # index of the source being polled
        self.begin_step()

    def begin_step(self):
        s, w, r = self.step, self.world, self.r
        # push: two half-line stores per peer (a 16-byte line = two independently visible 8-byte halves)
        self.todo = [("store", (r + i) % w, half) for i in range(1, w) for half in (0, 1)]
        self.polls = [q for q in range(w) if q != r]      # fixed rank order, own contribution comes from registers

    def done(self):
        return self.step >= self.steps


def value(rank, step, half):
    return (rank, step, half)


def run(world, steps, choose, double_buffered=True, max_ticks=100000):
    """Runs the abstract machine under scheduler `choose(list_of_runnable_ranks) -> rank`.  Returns None when every rank
    finished all steps with correct data, else a string describing the violation."""
    nbuf = 2 if double_buffered else 1
    # inbox[dst][parity][src][half] = (value, epoch); epoch 0 = freshly zeroed
    inbox = [[[[(None, 0), (None, 0)] for _ in range(world)] for _ in range(nbuf)] for _ in range(world)]
    ranks = [Rank(r, world, steps) for r in range(world)]
    for _ in range(max_ticks):
        runnable = []
        for k in ranks:
            if k.done():
                continue
            if k.todo:
                runnable.append(k)
            else:                                    # polling: runnable only when the awaited line is complete
                q = k.polls[0]
                line = inbox[k.r][k.step % nbuf][q]
                if line[0][1] == k.step + 1 and line[1][1] == k.step + 1:
                    runnable.append(k)
        if all(k.done() for k in ranks):
            return None
        if not runnable:
            return "deadlock at steps %s" % [k.step for k in ranks]
        k = choose(runnable)
        if k.todo:
            _, dst, half = k.todo.pop(0)
            inbox[dst][k.step % nbuf][k.r][half] = (value(k.r, k.step, half), k.step + 1)
        else:
            q = k.polls.pop(0)
            line = inbox[k.r][k.step % nbuf][q]
            for half in (0, 1):
                if line[half][0] != value(q, k.step, half):
                    return f"rank {k.r} consumed {line[half][0]} for (rank {q}, step {k.step}, half {half})"
            if not k.polls:                          # all sources summed: SGD update, next step
                k.step += 1
                if not k.done():
                    k.begin_step()
    return "did not terminate"


def exhaustive(world, steps, double_buffered=True, limit=400000):
    """DFS over every interleaving (scheduler choice sequences), with replay."""
    stack, explored = [[]], 0
    while stack:
        prefix = stack.pop()
        trace = []

        def choose(runnable, prefix=prefix, trace=trace):
            i = len(trace)
            c = prefix[i] if i < len(prefix) else 0
            trace.append((c, len(runnable)))
            return runnable[c]

        bad = run(world, steps, choose, double_buffered)
        explored += 1
        if bad:
            return bad, explored
        if explored > limit:
            pytest.skip("state space larger than the exploration limit")
        # branch: for every choice point beyond the prefix, schedule the alternatives
        for i in range(len(prefix), len(trace)):
            for alt in range(1, trace[i][1]):
                stack.append([c for c, _ in trace[:i]] + [alt])
    return None, explored


def test_two_ranks_two_steps_every_interleaving_is_safe():
    bad, n = exhaustive(world=2, steps=2)
    assert bad is None, bad
    assert n > 100                                   # the search really branched


@pytest.mark.parametrize("world,steps", [(2, 6), (3, 5), (4, 4), (8, 3)])
def test_random_interleavings_are_safe(world, steps):
    for seed in range(300):
        rng = random.Random(seed * 7919 + world)
        # skewed schedulers (one rank much faster / slower than the others) are the interesting ones
        weights = [rng.choice([1, 1, 5, 25]) for _ in range(world)]
        bad = run(world, steps, lambda rs: rng.choices(rs, weights=[weights[k.r] for k in rs])[0])
        assert bad is None, (seed, bad)


def test_single_buffered_inbox_is_caught_by_the_model():
    """Without the parity double-buffer a fast rank overwrites a line its peer has not consumed: the peer then waits for
    an epoch that is gone (deadlock) -- the model must find such a schedule."""
    found = None
    for seed in range(400):
        rng = random.Random(seed)
        weights = [25, 1, 1]
        found = run(3, 4, lambda rs: rng.choices(rs, weights=[weights[k.r] for k in rs])[0], double_buffered=False)
        if found:
            break
    assert found is not None and ("deadlock" in found or "consumed" in found)


# ---------------------------------------------------------------------------------------------------------------------
# The flag barrier of csrc/common.cuh (`block_barrier_all_ranks`): ONE slot per (block, source rank), monotonically
# growing epochs, ">= epoch" wait.  Claim: one slot suffices because a peer can be at most one barrier ahead, and a rank
# that passes barrier k knows every peer has ARRIVED at barrier k (so its writes before the barrier are visible).
def run_barrier(world, rounds, choose, compare=lambda have, want: have >= want):
    flags = [[0] * world for _ in range(world)]          # flags[dst][src]
    arrived = [0] * world                                # highest barrier each rank has entered
    epoch = [0] * world
    todo = [[] for _ in range(world)]                    # pending stores of the barrier being executed
    passed = [0] * world
    for _ in range(200000):
        if all(p == rounds for p in passed):
            return None
        runnable = []
        for r in range(world):
            if passed[r] == rounds:
                continue
            if epoch[r] == passed[r]:                    # not inside a barrier: enter the next one
                runnable.append(r)
            elif todo[r]:
                runnable.append(r)
            elif all(compare(flags[r][q], epoch[r]) for q in range(world)):
                runnable.append(r)
        if not runnable:
            return "deadlock: passed=%s" % passed
        r = choose(runnable)
        if epoch[r] == passed[r]:
            epoch[r] += 1
            arrived[r] = epoch[r]
            todo[r] = list(range(world))                 # signal every peer (and itself)
        elif todo[r]:
            flags[todo[r].pop(0)][r] = epoch[r]
        else:
            k = epoch[r]
            if any(arrived[q] < k for q in range(world)):
                return f"rank {r} passed barrier {k} before everybody arrived: {arrived}"
            passed[r] = k
    return "did not terminate"


@pytest.mark.parametrize("world,rounds", [(2, 6), (3, 5), (8, 4)])
def test_flag_barrier_with_one_slot_per_source_is_safe(world, rounds):
    for seed in range(300):
        rng = random.Random(seed * 104729 + world)
        weights = [rng.choice([1, 1, 5, 25]) for _ in range(world)]
        bad = run_barrier(world, rounds, lambda rs: rng.choices(rs, weights=[weights[r] for r in rs])[0])
        assert bad is None, (seed, bad)


def test_flag_barrier_needs_the_monotonic_compare():
    """With an equality wait a fast peer that is already one barrier ahead overwrites the slot and the slow rank never
    sees 'its' epoch: the model must find the hang.  (The kernel waits for `>= epoch`, wrap-safe.)"""
    found = None
    for seed in range(400):
        rng = random.Random(seed)
        weights = [25, 1, 1]
        found = run_barrier(3, 5, lambda rs: rng.choices(rs, weights=[weights[r] for r in rs])[0], compare=lambda have, want: have == want)
        if found:
            break
    assert found is not None and "deadlock" in found
