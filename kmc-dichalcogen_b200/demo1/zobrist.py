"""Zobrist key of the lattice, maintained per logged event on the host (reference ZobristKey,
demo1/incremental.py:335-380, toggled by State's five mutators, state.py:199-248).

The reference draws one random value per entity on demand from the global ``random`` module, so its key
values depend on history and process; here the table is fixed (same as csrc/zobrist_kernel.cuh):

    value(site, code) = top `bits` bits of (x0 << 32 | x1),
    (x0, x1, ..) = Philox4x32-10(counter = (site, code, seed >> 32, 2), key = (seed & 0xffffffff, 0))

code = status byte of the entity's site: 1, 2, 3 for Vacancy(node, layers), 4 / 7 for an Up / Down trefoil
at its anchor.  Equal keys <=> equal lattices (up to 2^-bits collisions), across runs and processes.
"""
from . import tables as T

_MASK = 0xFFFFFFFF


def value(site, code, bits, seed):
    c = [site & _MASK, code & _MASK, (seed >> 32) & _MASK, 2]
    k0, k1 = seed & _MASK, 0
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k0, p1 & _MASK, (p0 >> 32) ^ c[3] ^ k1, p0 & _MASK]
        k0, k1 = (k0 + 0x9E3779B9) & _MASK, (k1 + 0xBB67AE85) & _MASK
    return ((c[0] << 32) | c[1]) >> (64 - bits)


def _carries(code):
    return 1 <= code <= 4 or code == 7


class ZobristTracker:
    """Incremental key: ``apply(event)`` replays a logged event (tables.replay_event) on a private copy of the
    status plane and toggles the values of the entities that disappeared and appeared."""

    def __init__(self, status, dim, bits=64, seed=0):
        if not 1 <= int(bits) <= 64:
            raise ValueError("zobrist bits must be 1..64")
        self.bits, self.seed, self.dim = int(bits), int(seed) & (2 ** 64 - 1), tuple(dim)
        self.status = [int(s) for s in status]
        self._memo = {}
        self.key = 0
        for site, code in enumerate(self.status):
            if _carries(code):
                self.key ^= self._value(site, code)

    def _value(self, site, code):
        k = (site, code)
        v = self._memo.get(k)
        if v is None:
            v = self._memo[k] = value(site, code, self.bits, self.seed)
        return v

    def apply(self, event):
        sites = T.event_sites(event, self.dim)
        old = [self.status[s] for s in sites]
        T.replay_event(self.status, event, self.dim)
        for s, o in zip(sites, old):
            n = self.status[s]
            if o != n:
                if _carries(o):
                    self.key ^= self._value(s, o)
                if _carries(n):
                    self.key ^= self._value(s, n)
        return self.key

    @staticmethod
    def from_scratch(status, bits=64, seed=0):
        key = 0
        for site, code in enumerate(status):
            if _carries(int(code)):
                key ^= value(site, int(code), int(bits), int(seed) & (2 ** 64 - 1))
        return key
