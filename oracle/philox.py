"""Philox4x32-10 counter-based RNG, restated in numpy (TEST INFRASTRUCTURE ONLY).

The reference never seeds its RNGs; parity is defined under *shared* uniform
draws.  This is the draw recipe every engine (oracle, CUDA) uses:

    key     = (seed & 0xffffffff, replica)
    counter = (step & 0xffffffff, step >> 32, seed >> 32, 0)
    x0..x3  = Philox4x32-10(counter, key)        (Salmon et al., SC'11, Random123 constants)
    u1      = (((x0 << 32) | x1) >> 11) * 2**-53      class draw  (replaces random.uniform, kmc.py:49)
    u2      = (((x2 << 32) | x3) >> 11) * 2**-53      in-class draw (replaces rng.choice, incremental.py:285)
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint64 arrays holding 32-bit values."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) & MASK for x in (c0, c1, c2, c3))
    k0 = np.asarray(k0, dtype=np.uint64) & MASK
    k1 = np.asarray(k1, dtype=np.uint64) & MASK
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return c0, c1, c2, c3


def draws(seed, replica, step0, nsteps):
    """(nsteps, 2) float64 array of (u1, u2) for steps step0 .. step0+nsteps-1."""
    seed = int(seed)
    steps = np.arange(step0, step0 + nsteps, dtype=np.uint64)
    x0, x1, x2, x3 = philox4x32_10(
        steps & MASK, steps >> np.uint64(32),
        np.full(nsteps, seed >> 32, dtype=np.uint64), np.zeros(nsteps, dtype=np.uint64),
        np.full(nsteps, seed & 0xFFFFFFFF, dtype=np.uint64),
        np.full(nsteps, int(replica) & 0xFFFFFFFF, dtype=np.uint64))
    u1 = (((x0 << np.uint64(32)) | x1) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    u2 = (((x2 << np.uint64(32)) | x3) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    return np.stack([u1, u2], axis=1)


def draws3(seed, replica, step0, nsteps):
    """u3 in (0, 1] for steps step0 .. step0+nsteps-1: the clock draw of the stochastic KMC clock,
    counter domain 3 of the same stream: ((((x0 << 32) | x1) >> 11) + 1) * 2**-53."""
    seed = int(seed)
    steps = np.arange(step0, step0 + nsteps, dtype=np.uint64)
    x0, x1, _, _ = philox4x32_10(
        steps & MASK, steps >> np.uint64(32),
        np.full(nsteps, seed >> 32, dtype=np.uint64), np.full(nsteps, 3, dtype=np.uint64),
        np.full(nsteps, seed & 0xFFFFFFFF, dtype=np.uint64),
        np.full(nsteps, int(replica) & 0xFFFFFFFF, dtype=np.uint64))
    return ((((x0 << np.uint64(32)) | x1) >> np.uint64(11)) + np.uint64(1)).astype(np.float64) * 2.0 ** -53


import random as _random


class PhiloxRandom(_random.Random):
    """random.Random over a Philox stream, used to drive the reference's gen_random_state (state.py:329-378)
    with reproducible bits (TEST INFRASTRUCTURE ONLY; the device generator csrc/stategen_kernel.cuh consumes
    the same stream):

        block(q) = Philox4x32-10(counter = (q & 0xffffffff, q >> 32, seed >> 32, 1), key = (seed & 0xffffffff, replica))
        random()       = (((x0 << 32) | x1) >> 11) * 2**-53          q += 1
        getrandbits(k) = x0 >> (32 - k),  1 <= k <= 32               q += 1

    Everything else (uniform, shuffle, choice, _randbelow with k = n.bit_length() and rejection) is
    CPython's random.Random."""

    def __init__(self, seed, replica=0):
        self.pseed, self.replica, self.q = int(seed), int(replica), 0
        super().__init__(0)

    def seed(self, *args, **kwargs):
        pass

    def _x(self):
        q = self.q
        self.q += 1
        x = philox4x32_10(q & 0xFFFFFFFF, q >> 32, self.pseed >> 32, 1, self.pseed & 0xFFFFFFFF,
                          self.replica & 0xFFFFFFFF)
        return [int(v) for v in x]

    def random(self):
        x = self._x()
        return float(((x[0] << 32) | x[1]) >> 11) * 2.0 ** -53

    def getrandbits(self, k):
        assert 0 < k <= 32
        return self._x()[0] >> (32 - k)
