"""Shared helpers for the parity tests: seeded lattices, rule tables, oracle access."""
import math
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import kmc_oracle as ko  # noqa: E402  (the checker)
import philox  # noqa: E402

# the 13 classes of the rules the reference implements itself (classes 13..16 = MoveMonovacancy, an extension)
ALL_ORDER = list(range(13))
TEST_RATES = [10.0, 20.0, 12.0, 3.0, 7.0, 7.0, 5000.0, 400.0, 10.0, 300.0, 100.0, 100.0, 55.0]

KB_EV = 1.38064852e-23 * 6.24150913e18        # reference demo1/sim.py:27-29
WSE2_BARRIERS = {2: 2.25, 3: 2.11199155, 4: 2.52009312, 5: 2.25, 6: 3.12293677, 7: 4.43051273,
                 8: 6.72584072, 9: 6.72584072, 12: 3.41217502}
WSE2_ORDER = [8, 9, 12, 2, 3, 4, 5, 6, 7]     # config/WSe2.yaml rule order x kinds() order


def wse2_rates(temperature):
    r = [0.0] * 13
    for c, e in WSE2_BARRIERS.items():
        r[c] = math.exp(-e / (temperature * KB_EV))     # sim.py:179-181
    return r


def exact_state(dim, frac_div, frac_mono, seed):
    """Same recipe as oracle/make_golden.py:exact_state."""
    rng = random.Random(seed)
    n = dim[0] * dim[1]
    nodes = list(range(n))
    rng.shuffle(nodes)
    nd = round(frac_div * n)
    nm = round((frac_div + frac_mono) * n)
    st = np.zeros(n, dtype=np.uint8)
    for i in nodes[:nd]:
        st[i] = 3
    for i in nodes[nd:nm]:
        st[i] = rng.choice([1, 2])
    return st


def evolved_state(dim, seed, nsteps=400, rates=None, order=None):
    """A lattice with every status code present (trefoils included): run the oracle for a while."""
    o = ko.Oracle(dim, rates or TEST_RATES, order or ALL_ORDER, exact_state(dim, 0.25, 0.15, seed))
    o.run_philox(seed, 0, 0, nsteps, log=False)
    return o.status()


def iid_state(dim, seed, p=(0.90, 0.025, 0.025, 0.05)):
    """SURVEY section 8(d) config-4 recipe: iid status per site, no trefoils."""
    rng = np.random.Generator(np.random.Philox(seed))
    return rng.choice(np.arange(4, dtype=np.uint8), size=dim[0] * dim[1], p=p).astype(np.uint8)


# all nine rules: the reference's eight + MoveMonovacancy (17 classes)
MM_ORDER = list(range(17))
MM_RATES = TEST_RATES + [30.0, 60.0, 20.0, 40.0]
WSE2_FULL_BARRIERS = {**WSE2_BARRIERS, 13: 2.56175687, 14: 3.03662851, 15: 2.33673358, 16: 2.56175687}
WSE2_FULL_ORDER = [8, 9, 12, 13, 14, 15, 16, 2, 3, 4, 5, 6, 7]      # config/WSe2.yaml rule order, every rule


def wse2_full_rates(temperature):
    r = [0.0] * 17
    for c, e in WSE2_FULL_BARRIERS.items():
        r[c] = math.exp(-e / (temperature * KB_EV))
    return r


def reference_rules_only(slots=None, counts=None, rows=None):
    """The oracle always enumerates all nine rules; an engine without MoveMonovacancy rates reports only the
    reference's own 13 classes / 17 slots.  Blank the rest of an oracle table before comparing."""
    out = []
    if slots is not None:
        s2 = np.array(slots, copy=True)
        s2[..., 17:] = 0xFF
        out.append(s2)
    if counts is not None:
        c2 = np.array(counts, copy=True)
        c2[..., 13:] = 0
        out.append(c2)
    if rows is not None:
        r2 = np.array(rows, copy=True)
        r2[..., 13:] = 0
        out.append(r2)
    return out[0] if len(out) == 1 else out
