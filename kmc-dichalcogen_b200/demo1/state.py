"""Lattice state on the host: a flat uint8 status plane (the device layout) with the reference's
entity view on top (reference demo1/state.py:43-317) for (de)serialisation and state generation.

status[a * dim[1] + b]: 0 pristine, 1/2 monovacancy (layer), 3 divacancy, 4 + 3*orientation + role for
trefoil members.  The reference keeps entity sets + a per-node dict; here the plane is authoritative
and the entity lists are derived on demand.
"""
import random

import numpy as np

from . import tables as T
from .kmc import weighted_choice

RANDOM_MODES = ["exact", "approx"]                 # reference state.py:319
RANDOM_PARAMS = ["divacancy", "monovacancy"]       # canonical order (the reference's is a set's)
LAYERS = [1, 2]


class State:
    def __init__(self, dim, status=None):
        dim = (int(dim[0]), int(dim[1]))
        if len(dim) != 2 or dim[0] <= 0 or dim[1] <= 0:
            raise ValueError("bad dimensions: %r" % (dim,))
        self.dim = dim
        n = dim[0] * dim[1]
        if status is None:
            self.status = np.zeros(n, dtype=np.uint8)
        else:
            self.status = np.array(status, dtype=np.uint8).reshape(n).copy()

    # -- reference-style accessors -------------------------------------------------------
    def clone(self):
        return State(self.dim, self.status)

    def _site(self, node):
        return (node[0] % self.dim[0]) * self.dim[1] + (node[1] % self.dim[1])

    def vacant_layerset_at(self, node):
        s = int(self.status[self._site(node)])
        if s > 3:
            raise KeyError("monovacancy layers not defined at node")
        return s

    def is_pristine(self, node):
        return self.status[self._site(node)] == T.PRISTINE

    def is_divacancy(self, node):
        return self.status[self._site(node)] == T.DIVACANCY

    def is_monovacancy(self, node):
        return self.status[self._site(node)] in (T.MONO1, T.MONO2)

    def is_trefoil(self, node):
        return self.status[self._site(node)] >= T.TREFOIL_BASE

    def new_vacancy(self, node, layers):
        i = self._site(node)
        assert self.status[i] == 0, "node is not pristine"
        assert 1 <= layers <= 3
        self.status[i] = layers

    def new_divacancy(self, node):
        self.new_vacancy(node, T.DIVACANCY)

    def new_trefoil(self, nodes):
        nodes = [tuple(n) for n in nodes]
        anchor, orient = self._anchor_of(nodes)
        for role, off in enumerate(T.TREFOIL_MEMBERS[orient]):
            i = self._site(T.add(anchor, off, self.dim))
            assert self.status[i] == 0, "node is not pristine"
            self.status[i] = T.TREFOIL_BASE + 3 * orient + role

    def _anchor_of(self, nodes):
        want = set(T.add(n, (0, 0), self.dim) for n in nodes)
        for p in want:
            for orient in (0, 1):
                if set(T.add(p, off, self.dim) for off in T.TREFOIL_MEMBERS[orient]) == want:
                    return p, orient
        raise ValueError("nodes do not form a trefoil: %r" % (nodes,))

    # -- entity views / serialisation (reference state.py:75-92) ------------------------
    def vacancies(self):
        d1 = self.dim[1]
        return [((int(i) // d1, int(i) % d1), int(self.status[i]))
                for i in np.nonzero((self.status >= 1) & (self.status <= 3))[0]]

    def trefoils(self):
        out = []
        for i in np.nonzero((self.status == 4) | (self.status == 7))[0]:
            orient = (int(self.status[i]) - T.TREFOIL_BASE) // 3
            p = T.node_of(int(i), self.dim)
            out.append([T.add(p, off, self.dim) for off in T.TREFOIL_MEMBERS[orient]])
        return out

    def to_dict(self):
        return {
            "dim": list(self.dim),
            "vacancies": [[list(node), layers] for (node, layers) in self.vacancies()],
            "trefoils": [[[list(n) for n in nodes]] for nodes in self.trefoils()],
        }

    @classmethod
    def from_dict(cls, d):
        dim = tuple(d["dim"])
        if len(dim) != 2 or not all(isinstance(x, int) and x > 0 for x in dim):
            raise ValueError("bad 'dim'")
        self = cls(dim)
        for node, layers in d["vacancies"]:
            self.new_vacancy(tuple(node), layers)
        for (nodes,) in d["trefoils"]:
            self.new_trefoil(nodes)
        return self

    def validate(self):
        """Entity invariants (reference state.py:102-149): known codes and complete trefoils."""
        s = self.status
        if (s > 9).any():
            raise AssertionError("invalid status byte")
        for i in np.nonzero(s >= T.TREFOIL_BASE)[0]:
            orient, role = divmod(int(s[i]) - T.TREFOIL_BASE, 3)
            off = T.TREFOIL_MEMBERS[orient][role]
            anchor = T.add(T.node_of(int(i), self.dim), (-off[0], -off[1]), self.dim)
            for r, o in enumerate(T.TREFOIL_MEMBERS[orient]):
                if s[self._site(T.add(anchor, o, self.dim))] != T.TREFOIL_BASE + 3 * orient + r:
                    raise AssertionError("incomplete trefoil at %r" % (anchor,))
        return True


# ---- random initial states (reference state.py:321-378) ------------------------------------------
def _assign(state, key, node, rng):
    if key == "divacancy":
        state.new_divacancy(node)
    elif key == "monovacancy":
        state.new_vacancy(node, rng.choice(LAYERS))


class PhiloxRandom(random.Random):
    """A ``random.Random`` whose bits come from the Philox stream the device generator uses
    (csrc/stategen_kernel.cuh): call q of replica r is Philox4x32-10(counter = (q lo, q hi, seed hi, 1),
    key = (seed lo, r)); ``random()`` = 53 high bits of (x0, x1), ``getrandbits(k <= 32)`` = top k bits of
    x0.  shuffle / choice / uniform are CPython's own on top of those two."""

    def __init__(self, seed, replica=0):
        self._pseed, self._replica, self._q = int(seed) & (2 ** 64 - 1), int(replica) & 0xFFFFFFFF, 0
        super().__init__(0)

    def seed(self, *a, **k):              # the stream is fixed by (seed, replica); random.Random.__init__ calls this
        pass

    def _block(self):
        q, s = self._q, self._pseed
        self._q += 1
        c = [q & 0xFFFFFFFF, (q >> 32) & 0xFFFFFFFF, s >> 32, 1]
        k0, k1 = s & 0xFFFFFFFF, self._replica
        for _ in range(10):
            p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k0, p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k1, p0 & 0xFFFFFFFF]
            k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
        return c

    def random(self):
        c = self._block()
        return float(((c[0] << 32) | c[1]) >> 11) * 2.0 ** -53

    def getrandbits(self, k):
        if not 0 < k <= 32:
            raise ValueError("PhiloxRandom.getrandbits supports 1..32 bits")
        return self._block()[0] >> (32 - k)


def generator_plan(n_sites, mode, params):
    """What the reference derives before it starts drawing (state.py:342-378), as the arguments of
    kmc_generate_state: (mode id, status code per key, run bounds [exact], cumulative weights [approx])."""
    import math
    params = {k: float(params.get(k, 0.0)) for k in RANDOM_PARAMS}
    code_of = {"divacancy": 3, "monovacancy": 1, "remainder": 0}
    if mode == "exact":
        keys = sorted(RANDOM_PARAMS, key=lambda k: params[k])
        cumul, acc = [0], 0
        for k in keys:
            acc = acc + params[k]
            cumul.append(acc)
        return 0, [code_of[k] for k in keys], [round(r * n_sites) for r in cumul], []
    if mode == "approx":
        pairs = [(k, params[k]) for k in RANDOM_PARAMS]
        rest = 1.0 - math.fsum(params.values())
        if rest > 0:
            pairs.append(("remainder", rest))
        live = [(k, w) for (k, w) in pairs if w != 0.0]          # weighted_choice (kmc.py:30-41)
        if not live:
            raise ValueError("Cannot choose from total weight of zero!")
        live.sort(key=lambda kw: kw[1])
        cumul, acc = [], 0
        for _, w in live:
            acc = acc + w
            cumul.append(acc)
        return 1, [code_of[k] for k, _ in live], [], cumul
    raise AssertionError("complete switch")


def gen_random_state(dim, mode, params, rng=random, **_kw):
    """Same procedure and the same draws from ``rng`` as the reference's generators, so a seeded
    ``random.Random`` gives the reference's lattice.  Ties between equal fractions are broken by
    RANDOM_PARAMS order (the reference iterates a set there, i.e. hash order)."""
    state = State(dim)
    params = {k: float(params.get(k, 0.0)) for k in RANDOM_PARAMS}
    nodes = [(a, b) for a in range(state.dim[0]) for b in range(state.dim[1])]
    if mode == "exact":
        rng.shuffle(nodes)
        keys = sorted(RANDOM_PARAMS, key=lambda k: params[k])           # stable: ties keep list order
        cumul, acc = [0], 0
        for k in keys:
            acc = acc + params[k]
            cumul.append(acc)
        idx = [round(r * len(nodes)) for r in cumul]
        for k, (start, end) in zip(keys, zip(idx[:-1], idx[1:])):
            for node in nodes[start:end]:
                _assign(state, k, node, rng)
    elif mode == "approx":
        import math
        pairs = [(k, params[k]) for k in RANDOM_PARAMS]
        rest = 1.0 - math.fsum(params.values())
        if rest > 0:
            pairs.append(("remainder", rest))
        chosen = weighted_choice(pairs, howmany=len(nodes), rng=rng)
        for node, key in zip(nodes, chosen):
            _assign(state, key, node, rng)
    else:
        raise AssertionError("complete switch")
    return state
