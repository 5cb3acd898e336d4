"""Host-side weighted choice (reference demo1/kmc.py:21-54), used only where the reference uses it
on the host: the 'approx' initial-state generator.  The per-event class choice runs on the device
(csrc/replica_kernel.cuh:pick_class) with the same arithmetic."""
import bisect
import random

__all__ = ["weighted_choice"]


def weighted_choice(choices, howmany=None, rng=random):
    """Weighted random choice over (value, weight) pairs; weights need not sum to 1.

    Semantics follow the reference: zero weights are dropped, the rest stably sorted by weight,
    cumulative sums taken sequentially, one ``rng.uniform(0, total)`` per draw, ``bisect_left``."""
    live = [(v, w) for (v, w) in choices if w != 0.0]
    if not live:
        raise ValueError("Cannot choose from total weight of zero!")
    live.sort(key=lambda vw: vw[1])
    if live[0][1] < 0.0:
        raise ValueError("Negative weight {!s} for {!r}".format(live[0][1], live[0][0]))
    cumulative, acc = [], 0
    for _, w in live:
        acc = acc + w
        cumulative.append(acc)
    total = cumulative[-1]

    def one():
        return live[bisect.bisect_left(cumulative, rng.uniform(0.0, total))][0]

    return one() if howmany is None else [one() for _ in range(howmany)]
