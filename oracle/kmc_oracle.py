"""ctypes front-end of the CPU oracle ``oracle/kmc_oracle.c`` (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module,
and only as the checker.  The product never does.  See the header of kmc_oracle.c for what is
restated from the reference (file:line) and how it is pinned.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libkmc_oracle.so")
NC, NS, NONE = 17, 29, 0xFF

EVENT_DTYPE = np.dtype([
    ("site", "<u4"), ("cls", "u1"), ("slot", "u1"), ("flags", "<u2"),
    ("total_rate", "<f8"), ("total_naive", "<f8"), ("counts", "<u4", (NC,)), ("pad", "<u4"),
])
assert EVENT_DTYPE.itemsize == 96


def build(force=False):
    """Compile the oracle with gcc (also done by ``__graft_entry__.build()``)."""
    src = os.path.join(HERE, "kmc_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "libkmc_oracle.so"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        P = C.POINTER
        L.kmco_create.restype = C.c_void_p
        L.kmco_create.argtypes = [C.c_int, C.c_int, P(C.c_double), P(C.c_int), C.c_int]
        L.kmco_destroy.argtypes = [C.c_void_p]
        L.kmco_set_status.argtypes = [C.c_void_p, C.c_void_p]
        L.kmco_get_status.argtypes = [C.c_void_p, C.c_void_p]
        L.kmco_get_slots.argtypes = [C.c_void_p, C.c_void_p]
        L.kmco_get_counts.argtypes = [C.c_void_p, C.c_void_p]
        L.kmco_get_hist.argtypes = [C.c_void_p, C.c_void_p]
        L.kmco_get_ties.restype = C.c_int64
        L.kmco_get_ties.argtypes = [C.c_void_p]
        L.kmco_get_steps.restype = C.c_int64
        L.kmco_get_steps.argtypes = [C.c_void_p]
        L.kmco_reset_stats.argtypes = [C.c_void_p]
        L.kmco_sweep.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.kmco_step.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        L.kmco_rank_to_move_scan.argtypes = [C.c_void_p, C.c_int, C.c_int64, P(C.c_int), P(C.c_int)]
        L.kmco_perform_given.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.kmco_validate.argtypes = [C.c_void_p]
        L.kmco_draw.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, P(C.c_double), P(C.c_double)]
        L.kmco_run_draws.restype = C.c_int64
        L.kmco_run_draws.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.kmco_run_philox.restype = C.c_int64
        L.kmco_run_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int64, C.c_void_p]
        L.kmco_run_replicas.restype = C.c_int64
        L.kmco_run_replicas.argtypes = [C.c_int, C.c_int, P(C.c_double), P(C.c_int), C.c_int, C.c_void_p,
                                        C.c_int, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int64,
                                        C.c_void_p, C.c_int]
        L.kmco_log.restype = C.c_double
        L.kmco_log.argtypes = [C.c_double]
        L.kmco_draw3.restype = C.c_double
        L.kmco_draw3.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64]
        L.kmco_set_clock.argtypes = [C.c_void_p, C.c_int]
        L.kmco_get_clock.restype = C.c_double
        L.kmco_get_clock.argtypes = [C.c_void_p]
        L.kmco_set_clock_key.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        L.kmco_enable_tracers.argtypes = [C.c_void_p]
        L.kmco_get_tracers.argtypes = [C.c_void_p, C.c_void_p]
        L.kmco_get_msd.argtypes = [C.c_void_p, P(C.c_int64), P(C.c_int64)]
        assert L.kmco_event_size() == EVENT_DTYPE.itemsize
        _lib = L
    return _lib


def _rates(rates):
    rates = [float(x) for x in rates]
    r = (C.c_double * NC)(*(rates + [0.0] * (NC - len(rates))))      # 13-entry tables predate MoveMonovacancy
    return r


def _order(order):
    return (C.c_int * max(1, len(order)))(*[int(x) for x in order])


def sweep(status, dim):
    """Stateless full enumeration.  Returns (slots[n,17] uint8, counts[13] int64, rowcnt[d0,13] int32)."""
    status = np.ascontiguousarray(status, dtype=np.uint8).ravel()
    d0, d1 = dim
    assert status.size == d0 * d1
    slots = np.empty((d0 * d1, NS), dtype=np.uint8)
    counts = np.zeros(NC, dtype=np.int64)
    rowcnt = np.zeros((d0, NC), dtype=np.int32)
    lib().kmco_sweep(status.ctypes.data, d0, d1, slots.ctypes.data, counts.ctypes.data, rowcnt.ctypes.data)
    return slots, counts, rowcnt


def log(x):
    """The engine's deterministic fp64 logarithm (kmc_common.cuh:kmc_log restated in kmc_oracle.c)."""
    return float(lib().kmco_log(float(x)))


def draw3(seed, replica, step):
    return float(lib().kmco_draw3(int(seed), int(replica), int(step)))


def draw(seed, replica, step):
    u1, u2 = C.c_double(), C.c_double()
    lib().kmco_draw(seed, replica, step, C.byref(u1), C.byref(u2))
    return u1.value, u2.value


class Oracle:
    """One trajectory on the CPU oracle."""

    def __init__(self, dim, rates, class_order, status=None):
        self.dim = (int(dim[0]), int(dim[1]))
        self.n = self.dim[0] * self.dim[1]
        self._h = lib().kmco_create(self.dim[0], self.dim[1], _rates(rates), _order(class_order),
                                    len(class_order))
        if not self._h:
            raise ValueError("bad oracle parameters")
        if status is not None:
            self.set_status(status)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().kmco_destroy(self._h)
            self._h = None

    def set_status(self, status):
        status = np.ascontiguousarray(status, dtype=np.uint8).ravel()
        assert status.size == self.n
        lib().kmco_set_status(self._h, status.ctypes.data)

    def status(self):
        out = np.empty(self.n, dtype=np.uint8)
        lib().kmco_get_status(self._h, out.ctypes.data)
        return out

    def slots(self):
        out = np.empty((self.n, NS), dtype=np.uint8)
        lib().kmco_get_slots(self._h, out.ctypes.data)
        return out

    def counts(self):
        out = np.empty(NC, dtype=np.int64)
        lib().kmco_get_counts(self._h, out.ctypes.data)
        return out

    def hist(self):
        out = np.empty(NC, dtype=np.int64)
        lib().kmco_get_hist(self._h, out.ctypes.data)
        return out

    def ties(self):
        return int(lib().kmco_get_ties(self._h))

    def step(self, u1, u2):
        ev = np.zeros(1, dtype=EVENT_DTYPE)
        rc = lib().kmco_step(self._h, float(u1), float(u2), ev.ctypes.data)
        return rc, ev[0]

    def run_draws(self, u, log=True):
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1, 2)
        ev = np.zeros(len(u), dtype=EVENT_DTYPE) if log else None
        done = lib().kmco_run_draws(self._h, u.ctypes.data, len(u), ev.ctypes.data if log else None)
        return int(done), ev

    def run_philox(self, seed, replica, step0, nsteps, log=True):
        ev = np.zeros(nsteps, dtype=EVENT_DTYPE) if log else None
        done = lib().kmco_run_philox(self._h, seed, replica, step0, nsteps, ev.ctypes.data if log else None)
        return int(done), ev

    def perform_given(self, site, slot):
        return lib().kmco_perform_given(self._h, int(site), int(slot))

    def rank_to_move_scan(self, cls, rank):
        s, j = C.c_int(), C.c_int()
        rc = lib().kmco_rank_to_move_scan(self._h, int(cls), int(rank), C.byref(s), C.byref(j))
        return rc, s.value, j.value

    # -- extensions: KMC clock and divacancy tracers (same definitions as the CUDA engine) -----------
    def set_clock(self, mode):
        """1: t += 1/total_rate; 2: t += -ln(u3)/total_rate (u3 = third Philox draw of the event)."""
        lib().kmco_set_clock(self._h, int(mode))

    def set_clock_key(self, seed, replica):
        """Philox key of u3 when the class / in-class draws are injected (run_draws)."""
        lib().kmco_set_clock_key(self._h, int(seed), int(replica))

    def clock(self):
        return float(lib().kmco_get_clock(self._h))

    def enable_tracers(self):
        lib().kmco_enable_tracers(self._h)

    def tracers(self):
        out = np.empty((self.n, 2), dtype=np.int32)
        lib().kmco_get_tracers(self._h, out.ctypes.data)
        return out

    def msd(self):
        sq, cnt = C.c_int64(), C.c_int64()
        lib().kmco_get_msd(self._h, C.byref(sq), C.byref(cnt))
        return sq.value, cnt.value

    def validate(self):
        """0 when the incremental tables equal a from-scratch enumeration (sim.py:159-168)."""
        return lib().kmco_validate(self._h)


def run_replicas(dim, rates, class_order, status, seed, nsteps, replica0=0, step0=0, threads=1):
    """Batch of independent trajectories; ``status`` [R, n] is updated in place.
    Returns (events_done, hist[13])."""
    status = np.ascontiguousarray(status, dtype=np.uint8)
    assert status.ndim == 2 and status.shape[1] == dim[0] * dim[1]
    hist = np.zeros(NC, dtype=np.int64)
    done = lib().kmco_run_replicas(dim[0], dim[1], _rates(rates), _order(class_order), len(class_order),
                                   status.ctypes.data, status.shape[0], replica0, seed, step0, nsteps,
                                   hist.ctypes.data, threads)
    return int(done), hist, status
