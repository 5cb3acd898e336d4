"""Zobrist key of a lattice from scratch, numpy (TEST INFRASTRUCTURE ONLY -- the checker for kmc_zobrist
and demo1.zobrist; never imported by the product).

The reference (demo1/incremental.py:335-380, state.py:199-248) XORs one random value per entity on the
lattice -- Vacancy(node, layers) and Trefoil(frozenset of 3 nodes) -- drawing the values on demand from the
global `random` module, so its keys are history- and process-dependent and cannot be reproduced.  What is
reproducible, and what tests pin against the reference's own keys (tests/golden/zobrist_ref.npz), is the
structure: the key is a function of the entity set only, so two steps of a trajectory have equal keys iff
they have equal lattices.  Our value table is fixed:

    value(site, code) = top `bits` bits of (x0 << 32 | x1),
    (x0, x1, ..) = Philox4x32-10(counter = (site, code, seed >> 32, 2), key = (seed & 0xffffffff, 0))

with code = the status byte of the entity's site: 1, 2, 3 for a vacancy with that layer set, 4 / 7 for an
Up / Down trefoil at its anchor site (non-anchor members 5, 6, 8, 9 carry nothing).
"""
import numpy as np

import philox


def values(sites, codes, bits, seed):
    sites = np.asarray(sites, dtype=np.uint64)
    codes = np.asarray(codes, dtype=np.uint64)
    n = sites.size
    x0, x1, _, _ = philox.philox4x32_10(sites, codes, np.full(n, int(seed) >> 32, dtype=np.uint64),
                                        np.full(n, 2, dtype=np.uint64),
                                        np.full(n, int(seed) & 0xFFFFFFFF, dtype=np.uint64), np.zeros(n, dtype=np.uint64))
    v = (x0 << np.uint64(32)) | x1
    return v >> np.uint64(64 - bits)


def key(status, bits, seed):
    status = np.asarray(status, dtype=np.uint8).ravel()
    sel = np.flatnonzero((status == 1) | (status == 2) | (status == 3) | (status == 4) | (status == 7))
    if sel.size == 0:
        return 0
    return int(np.bitwise_xor.reduce(values(sel, status[sel], bits, seed)))
