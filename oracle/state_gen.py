"""CPU restatement of the reference's initial-state generators under a Philox stream
(TEST INFRASTRUCTURE ONLY -- the checker for kmc_generate_state; never imported by the product).

Follows /root/reference/demo1/state.py:342-356 ('approx': weighted_choice per node, then a layer per
monovacancy in node order), :360-378 ('exact': shuffle, runs of the shuffled list by rounded cumulative
fractions sorted ascending), kmc.py:21-54 (weighted_choice: drop zero weights, stable sort, sequential
sums, uniform(0, total), bisect_left) and CPython's random.Random.shuffle / choice / _randbelow.
Pinned against the reference itself in tests/golden/state_gen_philox.npz (oracle/make_golden.py).
Ties between equal fractions: canonical key order (divacancy, monovacancy[, remainder]); the reference
iterates a set there."""
import bisect
import math

import numpy as np

import philox

KEYS = ["divacancy", "monovacancy"]


class _Stream:
    """The draw recipe of philox.PhiloxRandom, written out without random.Random."""

    def __init__(self, seed, replica):
        self.seed, self.replica, self.q = int(seed), int(replica), 0

    def block(self):
        q = self.q
        self.q += 1
        x = philox.philox4x32_10(q & 0xFFFFFFFF, q >> 32, self.seed >> 32, 1, self.seed & 0xFFFFFFFF,
                                 self.replica & 0xFFFFFFFF)
        return int(x[0]), int(x[1])

    def random(self):
        x0, x1 = self.block()
        return float(((x0 << 32) | x1) >> 11) * 2.0 ** -53

    def randbelow(self, n):
        k = n.bit_length()
        while True:
            r = self.block()[0] >> (32 - k)
            if r < n:
                return r


def generate(dim, mode, params, seed, replica=0):
    """status plane (uint8[dim0*dim1]) of gen_random_state(dim, mode, params, rng=PhiloxRandom(seed, replica))."""
    n = dim[0] * dim[1]
    p = {k: float(params.get(k, 0.0)) for k in KEYS}
    g = _Stream(seed, replica)
    st = np.zeros(n, dtype=np.uint8)
    if mode == "exact":
        nodes = list(range(n))
        for i in range(n - 1, 0, -1):                 # random.shuffle
            j = g.randbelow(i + 1)
            nodes[i], nodes[j] = nodes[j], nodes[i]
        keys = sorted(KEYS, key=lambda k: p[k])
        cumul, acc = [0], 0
        for k in keys:
            acc = acc + p[k]
            cumul.append(acc)
        idx = [round(r * n) for r in cumul]
        for k, (a, b) in zip(keys, zip(idx[:-1], idx[1:])):
            for node in nodes[a:b]:
                st[node] = 3 if k == "divacancy" else 1 + g.randbelow(2)      # rng.choice([1, 2])
    elif mode == "approx":
        pairs = [(k, p[k]) for k in KEYS]
        rest = 1.0 - math.fsum(p.values())
        if rest > 0:
            pairs.append(("remainder", rest))
        live = sorted([(k, w) for k, w in pairs if w != 0.0], key=lambda kw: kw[1])
        cumul, acc = [], 0
        for _, w in live:
            acc = acc + w
            cumul.append(acc)
        total = cumul[-1]
        chosen = [live[bisect.bisect_left(cumul, 0.0 + (total - 0.0) * g.random())][0] for _ in range(n)]
        for node, k in enumerate(chosen):
            if k == "divacancy":
                st[node] = 3
            elif k == "monovacancy":
                st[node] = 1 + g.randbelow(2)
    else:
        raise AssertionError("complete switch")
    return st
