"""Timing model of the fused bitmap exchange (gpu-pruner_b200/csrc/gpr_kernels.cuh: fold_words,
exchange_bitmaps_ll, k_fold; gpr_api.cu: decide_impl's scratch sets and exchange buffer sets).

The tagged-slot exchange has no acknowledgements: rank A overwrites the slots it filled D steps ago, and the safety
of that rests on a chain of happens-before edges through the scratch-set guard of the reduce kernels.  This test
restates those edges as a max-plus recurrence over random (and adversarial) kernel durations and checks, for every
pair of ranks and every step, that a slot is never overwritten while its receiver may still be polling it:

    push_A(n + D)  >  collected_B(n)            for all A != B

Edges (S scratch sets, D exchange buffer sets, step n on one rank):
    R_n may start once R_{n-1} has started            (programmatic dependent launch: nothing orders it behind F_{n-1})
    R_n ends after its first publish, which waits for F_{n-S} done          (wait_scratch_free)
    F_n starts when R_n has ended                                           (griddepcontrol.wait)
    in-order protocol : F_n pushes after F_{n-1} is done
    pipelined protocol: F_n pushes right away; only the assembling CTA waits for F_{n-1}
    F_n has collected step n when every peer's push of step n has landed
    F_n is done after that and after F_{n-1} is done

The recurrence is evaluated step by step, which is itself the proof that no wait cycle exists: every quantity of
step n depends on quantities of earlier steps, or on pushes of step n, which depend on earlier steps only."""
import random

import pytest


def simulate(world, steps, S, D, pipelined, rng, slow_rank=None, zero=False):
    """Returns (push, collected): per rank, per step times."""
    def dur(lo, hi, rank):
        if zero:
            return 0.0
        x = rng.uniform(lo, hi)
        if slow_rank is not None and rank == slow_rank:
            x *= rng.choice((1.0, 1.0, 5.0, 40.0))   # a rank that falls steps behind now and then
        return x

    r_start = [[0.0] * steps for _ in range(world)]
    r_end = [[0.0] * steps for _ in range(world)]
    push = [[0.0] * steps for _ in range(world)]
    collected = [[0.0] * steps for _ in range(world)]
    done = [[0.0] * steps for _ in range(world)]
    for n in range(steps):
        for a in range(world):
            start = (r_start[a][n - 1] if n else 0.0) + dur(0.0, 1.0, a)   # host launch cadence only
            r_start[a][n] = start
            first_publish = start + dur(0.0, 2.0, a)
            if n >= S:
                first_publish = max(first_publish, done[a][n - S])
            r_end[a][n] = first_publish + dur(0.0, 40.0, a)
            f_start = r_end[a][n]
            p = f_start + dur(0.0, 20.0, a)
            if not pipelined and n:
                p = max(p, done[a][n - 1])
            push[a][n] = p
        for a in range(world):
            arrive = max(push[b][n] + dur(0.0, 3.0, b) for b in range(world) if b != a) if world > 1 else 0.0
            collected[a][n] = max(push[a][n], arrive) + dur(0.0, 1.0, a)
            d = collected[a][n] + dur(0.0, 4.0, a)
            if n:
                d = max(d, done[a][n - 1])
            done[a][n] = d
    return push, collected


def violations(push, collected, D):
    world, steps = len(push), len(push[0])
    bad = 0
    for a in range(world):
        for b in range(world):
            if a == b:
                continue
            for n in range(steps - D):
                if not push[a][n + D] >= collected[b][n]:
                    bad += 1
    return bad


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("pipelined", [False, True])
def test_slots_are_never_overwritten_early(world, pipelined):
    # what the library ships: two scratch sets, four exchange buffer sets
    for seed in range(60):
        rng = random.Random(seed * 7919 + world)
        slow = None if seed % 3 == 0 else rng.randrange(world)
        push, collected = simulate(world, 64, S=2, D=4, pipelined=pipelined, rng=rng, slow_rank=slow)
        assert violations(push, collected, 4) == 0, (seed, world, pipelined)
    push, collected = simulate(world, 32, S=2, D=4, pipelined=pipelined, rng=random.Random(1), zero=True)
    assert violations(push, collected, 4) == 0


def test_in_order_protocol_needs_only_two_buffer_sets():
    # round 2's first form: a fold pushes after its predecessor is done, so two buffer sets were enough
    for seed in range(60):
        rng = random.Random(seed)
        push, collected = simulate(4, 64, S=2, D=2, pipelined=False, rng=rng, slow_rank=seed % 4)
        assert violations(push, collected, 2) == 0


def test_pipelined_protocol_would_break_with_two_buffer_sets():
    # the reason for the deeper buffers: with the early push, depth 2 lets a fast rank overwrite step n's slots
    # while a peer that is waiting for a third, slow rank has not read them yet
    bad = 0
    for seed in range(60):
        rng = random.Random(seed)
        push, collected = simulate(4, 64, S=2, D=2, pipelined=True, rng=rng, slow_rank=seed % 4)
        bad += violations(push, collected, 2)
    assert bad > 0
