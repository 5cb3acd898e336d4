"""CPU checks of the host-callable pieces of the CUDA sources (no GPU needed): the byte-parallel sweep
arithmetic and the per-site evaluator of the replica kernel, against the oracle's full enumeration."""
import ctypes as C

import numpy as np
import pytest

from util_cases import ko, philox, evolved_state, iid_state
from demo1 import _native as nat


@pytest.fixture(scope="module")
def lib():
    L = nat.load()
    L.kmc_hostcheck_sweep_words.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.kmc_hostcheck_sweep_words16.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.kmc_hostcheck_sweep_nibbles.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.kmc_hostcheck_eval_sites.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.kmc_hostcheck_eval_sites_mm.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.kmc_hostcheck_draw.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.POINTER(C.c_double),
                                     C.POINTER(C.c_double)]
    return L


CASES = [((8, 8), 1), ((5, 8), 2), ((7, 12), 3), ((16, 16), 4), ((40, 72), 5), ((64, 64), 6), ((9, 20), 7)]


@pytest.mark.parametrize("dim,seed", CASES)
def test_sweep_words_match_oracle(lib, dim, seed):
    for st in (evolved_state(dim, seed), iid_state(dim, seed), np.full(dim[0] * dim[1], 3, np.uint8)):
        n = st.size
        slots = np.empty((17, n), dtype=np.uint8)
        counts = np.zeros(13, dtype=np.int64)
        assert lib.kmc_hostcheck_sweep_words(st.ctypes.data, dim[0], dim[1], slots.ctypes.data,
                                             counts.ctypes.data) == 0
        # (these kernels enumerate the 13 classes / 17 slots of the reference's own rules)
        want_slots, want_counts, _ = ko.sweep(st, dim)
        assert (slots.T == want_slots[:, :17]).all()
        assert (counts == want_counts[:13]).all()


@pytest.mark.parametrize("dim,seed", [((8, 16), 1), ((16, 16), 2), ((5, 32), 3), ((64, 64), 4), ((7, 48), 5),
                                      ((40, 80), 6)])
@pytest.mark.parametrize("fn", ["kmc_hostcheck_sweep_words16", "kmc_hostcheck_sweep_nibbles"])
def test_sweep_words16_match_oracle(lib, dim, seed, fn):
    for st in (evolved_state(dim, seed), iid_state(dim, seed), np.full(dim[0] * dim[1], 3, np.uint8)):
        n = st.size
        slots = np.empty((17, n), dtype=np.uint8)
        counts = np.zeros(13, dtype=np.int64)
        assert getattr(lib, fn)(st.ctypes.data, dim[0], dim[1], slots.ctypes.data, counts.ctypes.data) == 0
        # (these kernels enumerate the 13 classes / 17 slots of the reference's own rules)
        want_slots, want_counts, _ = ko.sweep(st, dim)
        assert (slots.T == want_slots[:, :17]).all()
        assert (counts == want_counts[:13]).all()


@pytest.mark.parametrize("dim,seed", CASES + [((5, 5), 8), ((7, 9), 9), ((11, 5), 10)])
def test_eval_site_matches_oracle(lib, dim, seed):
    for st in (evolved_state(dim, seed), iid_state(dim, seed)):
        n = st.size
        out = np.empty((n, 13), dtype=np.uint8)
        lib.kmc_hostcheck_eval_sites(st.ctypes.data, dim[0], dim[1], out.ctypes.data)
        want_slots, _, _ = ko.sweep(st, dim)
        want = np.stack([(want_slots == c).sum(axis=1) for c in range(13)], axis=1)
        assert (out == want).all()
        # the MoveMonovacancy-aware evaluator: all 17 classes
        out17 = np.empty((n, 17), dtype=np.uint8)
        lib.kmc_hostcheck_eval_sites_mm(st.ctypes.data, dim[0], dim[1], out17.ctypes.data)
        want17 = np.stack([(want_slots == c).sum(axis=1) for c in range(17)], axis=1)
        assert (out17 == want17).all()


def test_device_draw_recipe_matches_oracle(lib):
    for seed, rep, step in [(0, 0, 0), (20261019, 7, 123), ((1 << 45) + 3, 4000000000, (1 << 35) + 9)]:
        u1, u2 = C.c_double(), C.c_double()
        lib.kmc_hostcheck_draw(seed, rep, step, C.byref(u1), C.byref(u2))
        assert (u1.value, u2.value) == ko.draw(seed, rep, step)
        assert (u1.value, u2.value) == tuple(philox.draws(seed, rep, step, 1)[0])


def test_row_count_field_map_is_a_bijection(lib):
    lib.kmc_hostcheck_rc_field.argtypes = [C.c_int]
    fields = [lib.kmc_hostcheck_rc_field(c) for c in range(13)]
    assert len(set(fields)) == 13 and max(fields) < 14 and min(fields) >= 0


def test_clock_pieces_match_the_oracle_bit_for_bit(lib):
    """kmc_common.cuh:kmc_log / kmc_draw3 / kmc_time_step as compiled for the host == oracle/kmc_oracle.c."""
    import math
    lib.kmc_hostcheck_log.restype = C.c_double
    lib.kmc_hostcheck_log.argtypes = [C.c_double]
    lib.kmc_hostcheck_draw3.restype = C.c_double
    lib.kmc_hostcheck_draw3.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64]
    lib.kmc_hostcheck_time_step.restype = C.c_double
    lib.kmc_hostcheck_time_step.argtypes = [C.c_int, C.c_double, C.c_uint64, C.c_uint32, C.c_uint64]
    rng = np.random.default_rng(9)
    for x in np.concatenate([rng.random(5000), [1.0, 2.0 ** -53, 0.5, 1.0 - 2.0 ** -53, 0.7071067811865476]]):
        if x > 0:
            assert lib.kmc_hostcheck_log(float(x)) == ko.log(float(x))
    for seed, rep, step in [(0, 0, 0), (20261019, 7, 123), ((1 << 45) + 3, 4000000000, (1 << 35) + 9)]:
        u3 = lib.kmc_hostcheck_draw3(seed, rep, step)
        assert u3 == ko.draw3(seed, rep, step) == philox.draws3(seed, rep, step, 1)[0]
        assert lib.kmc_hostcheck_time_step(2, 3.5e-5, seed, rep, step) == (0.0 - ko.log(u3)) / 3.5e-5
        assert lib.kmc_hostcheck_time_step(1, 3.5e-5, seed, rep, step) == 1.0 / 3.5e-5
