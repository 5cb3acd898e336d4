"""Thin object wrapper over the C ABI: one `Engine` = `n_replicas` independent trajectories on one GPU.

This is the device-side half of `KmcSim` (reference demo1/sim.py:31-168); `demo1/sim.py` in this
package is the reference-facing half (rule specs, rates, log dicts).
"""
import ctypes as C

import numpy as np

from . import _native as nat


class Engine:
    def __init__(self, dim, class_rates, class_order, n_replicas=1, device=0):
        self.dim = (int(dim[0]), int(dim[1]))
        self.n = self.dim[0] * self.dim[1]
        self.n_replicas = int(n_replicas)
        self.device = int(device)
        self._lib = nat.load()
        self._h = C.c_void_p()
        class_rates = [float(x) for x in class_rates]
        class_rates += [0.0] * (nat.N_CLASSES - len(class_rates))       # 13-entry tables predate MoveMonovacancy
        rates = (C.c_double * nat.N_CLASSES)(*class_rates)
        order = (C.c_int * max(1, len(class_order)))(*[int(c) for c in class_order])
        nat.check(self._lib.kmc_create(C.byref(self._h), self.device, self.dim[0], self.dim[1],
                                       self.n_replicas, rates, order, len(class_order)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.kmc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- lattice ------------------------------------------------------------------------
    def upload_state(self, status):
        status = np.ascontiguousarray(status, dtype=np.uint8).reshape(self.n_replicas, self.n)
        nat.check(self._lib.kmc_upload_state(self._h, status.ctypes.data))

    def download_state(self):
        out = np.empty((self.n_replicas, self.n), dtype=np.uint8)
        nat.check(self._lib.kmc_download_state(self._h, out.ctypes.data))
        return out

    def download_state_into(self, out):
        assert out.dtype == np.uint8 and out.size == self.n_replicas * self.n and out.flags.c_contiguous
        nat.check(self._lib.kmc_download_state(self._h, out.ctypes.data))
        return out

    def generate_state(self, mode, params, seed, replica0=0):
        """gen_random_state (reference state.py:329-378) for every replica on the device; replica r gets what
        ``state.gen_random_state(dim, mode, params, rng=state.PhiloxRandom(seed, replica0 + r))`` gives."""
        from .state import generator_plan
        mode_id, codes, bounds, cumul = generator_plan(self.n, mode, params)
        codes = np.asarray(codes, dtype=np.int32)
        bounds = np.asarray(bounds, dtype=np.int64)
        cumul = np.asarray(cumul, dtype=np.float64)
        nat.check(self._lib.kmc_generate_state(self._h, mode_id, len(codes), codes.ctypes.data,
                                               bounds.ctypes.data if bounds.size else None,
                                               cumul.ctypes.data if cumul.size else None,
                                               int(seed) & (2 ** 64 - 1), int(replica0)))

    def set_state_device(self, dev_ptr):
        nat.check(self._lib.kmc_set_state_device(self._h, C.c_void_p(int(dev_ptr))))

    # -- full enumeration ---------------------------------------------------------------
    def sweep(self):
        nat.check(self._lib.kmc_sweep(self._h))

    def counts(self):
        out = np.empty((self.n_replicas, nat.N_CLASSES), dtype=np.int64)
        nat.check(self._lib.kmc_get_counts(self._h, out.ctypes.data))
        return out

    def slots(self):
        out = np.empty((self.n_replicas, self.n, nat.N_SLOTS), dtype=np.uint8)
        nat.check(self._lib.kmc_get_slots(self._h, out.ctypes.data))
        return out

    def row_counts(self):
        out = np.empty((self.n_replicas, self.dim[0], nat.N_CLASSES), dtype=np.uint32)
        nat.check(self._lib.kmc_get_row_counts(self._h, out.ctypes.data))
        return out

    # -- events -------------------------------------------------------------------------
    def _events_buffer(self, nsteps, log_mode):
        if log_mode == nat.LOG_NONE:
            return None
        dt = nat.EVENT_FULL if log_mode == nat.LOG_FULL else nat.EVENT_COMPACT
        return np.zeros((self.n_replicas, int(nsteps)), dtype=dt)

    def run(self, nsteps, seed, replica0=0, log_mode=nat.LOG_NONE, validate=False):
        ev = self._events_buffer(nsteps, log_mode)
        rc = self._lib.kmc_run(self._h, int(nsteps), int(seed), int(replica0), log_mode,
                               ev.ctypes.data if ev is not None else None, int(bool(validate)))
        self._last_events = ev
        nat.check(rc)
        return ev

    def step_draws(self, draws, log_mode=nat.LOG_FULL, validate=False):
        draws = np.ascontiguousarray(draws, dtype=np.float64)
        draws = draws.reshape(self.n_replicas, -1, 2)
        nsteps = draws.shape[1]
        ev = self._events_buffer(nsteps, log_mode)
        rc = self._lib.kmc_step_draws(self._h, draws.ctypes.data, nsteps, log_mode,
                                      ev.ctypes.data if ev is not None else None, int(bool(validate)))
        self._last_events = ev
        nat.check(rc)
        return ev

    def run_noinc(self, nsteps, seed, replica0=0, log_mode=nat.LOG_NONE):
        ev = self._events_buffer(nsteps, log_mode)
        rc = self._lib.kmc_run_noinc(self._h, int(nsteps), int(seed), int(replica0), log_mode,
                                     ev.ctypes.data if ev is not None else None)
        self._last_events = ev
        nat.check(rc)
        return ev

    def step_draws_noinc(self, draws, log_mode=nat.LOG_FULL):
        draws = np.ascontiguousarray(draws, dtype=np.float64).reshape(self.n_replicas, -1, 2)
        nsteps = draws.shape[1]
        ev = self._events_buffer(nsteps, log_mode)
        rc = self._lib.kmc_step_draws_noinc(self._h, draws.ctypes.data, nsteps, log_mode,
                                            ev.ctypes.data if ev is not None else None)
        self._last_events = ev
        nat.check(rc)
        return ev

    def check_status(self):
        """Wait for the queued launches and raise what they raised: ValueError when a replica stopped on a total
        rate of zero (reference kmc.py:31-32), AssertionError on a validation mismatch.  The synchronous
        counterpart of a statistics-only ``run`` (which returns as soon as the launch is queued)."""
        nat.check(self._lib.kmc_check_status(self._h))

    def perform_given(self, replica, site, slot):
        nat.check(self._lib.kmc_perform_given(self._h, int(replica), int(site), int(slot)))

    # -- statistics ---------------------------------------------------------------------
    def histogram(self):
        out = np.empty(nat.N_CLASSES, dtype=np.int64)
        nat.check(self._lib.kmc_get_histogram(self._h, out.ctypes.data))
        return out

    def histogram_per_replica(self):
        out = np.empty((self.n_replicas, nat.N_CLASSES), dtype=np.int64)
        nat.check(self._lib.kmc_get_histogram_per_replica(self._h, out.ctypes.data))
        return out

    def zobrist(self, bits=64, seed=0):
        """Zobrist key of each replica's current lattice, computed from scratch on the device (uint64[R])."""
        out = np.empty(self.n_replicas, dtype=np.uint64)
        nat.check(self._lib.kmc_zobrist(self._h, int(bits), int(seed) & (2 ** 64 - 1), out.ctypes.data))
        return out

    def hops(self):
        """MoveDivacancy events per replica and direction k (slot 7 + k), shape (n_replicas, 6)."""
        out = np.empty((self.n_replicas, 6), dtype=np.int64)
        nat.check(self._lib.kmc_get_hops(self._h, out.ctypes.data))
        return out

    def displacement(self):
        """Net axial displacement (da, db) of the divacancy population per replica: sum_k hops[k] * nbr_k with
        the neighbour order of the reference's neighbors_and_mutuals (state.py:443-459)."""
        from .tables import MOVE_TARGET as NBR
        return self.hops() @ np.asarray(NBR, dtype=np.int64)

    def steps_done(self):
        out = np.empty(self.n_replicas, dtype=np.int64)
        nat.check(self._lib.kmc_get_steps_done(self._h, out.ctypes.data))
        return out

    def ties(self):
        v = C.c_int64()
        nat.check(self._lib.kmc_get_ties(self._h, C.byref(v)))
        return v.value

    def enable_clock(self, mode=1):
        """0 / False off, 1 / True mean residence time (sum of 1 / total_rate), 2 or "stochastic": the BKL clock
        t += -ln(u3) / total_rate with a third Philox draw per event."""
        mode = 2 if mode == "stochastic" else int(mode)
        nat.check(self._lib.kmc_enable_clock(self._h, mode))

    def enable_tracers(self, on=True):
        """Track the unwrapped displacement of every divacancy (see include/kmc_b200.h)."""
        nat.check(self._lib.kmc_enable_tracers(self._h, int(bool(on))))

    def msd(self):
        """(sum of squared tracer displacements, number of tracers) per replica, int64 each; the squared length
        of an axial displacement (da, db) is da^2 + da*db + db^2 lattice constants^2."""
        sq = np.empty(self.n_replicas, dtype=np.int64)
        cnt = np.empty(self.n_replicas, dtype=np.int64)
        nat.check(self._lib.kmc_get_msd(self._h, sq.ctypes.data, cnt.ctypes.data))
        return sq, cnt

    def tracers(self, replica=0):
        """Raw tracer table of one replica as an (n, 2) int16 array of axial displacements (da, db); only the rows
        of sites that currently hold a divacancy are meaningful."""
        out = np.empty(self.n, dtype=np.uint32)
        nat.check(self._lib.kmc_get_tracers(self._h, int(replica), out.ctypes.data))
        return out.view(np.int16).reshape(self.n, 2)

    def clock(self):
        """Per-replica KMC time accumulated since the last reset_stats (see enable_clock)."""
        out = np.empty(self.n_replicas, dtype=np.float64)
        nat.check(self._lib.kmc_get_clock(self._h, out.ctypes.data))
        return out

    def reset_stats(self):
        nat.check(self._lib.kmc_reset_stats(self._h))

    # -- measurement --------------------------------------------------------------------
    def sync(self):
        nat.check(self._lib.kmc_sync(self._h))

    def flush_l2(self, nbytes=0):
        nat.check(self._lib.kmc_flush_l2(self._h, int(nbytes)))

    def probe_store_pattern(self):
        nat.check(self._lib.kmc_probe_store_pattern(self._h))

    def timer_start(self):
        nat.check(self._lib.kmc_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        nat.check(self._lib.kmc_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        v = C.c_int64()
        nat.check(self._lib.kmc_get_launch_count(self._h, C.byref(v)))
        return v.value

    def device_pointers(self):
        ptrs = [C.c_void_p() for _ in range(4)]
        nat.check(self._lib.kmc_get_device_pointers(self._h, *[C.byref(p) for p in ptrs]))
        return {k: p.value for k, p in zip(("status", "slots", "row_counts", "histogram"), ptrs)}

    def set_warps_per_block(self, warps):
        nat.check(self._lib.kmc_set_warps_per_block(self._h, int(warps)))

    def set_replica_kernel(self, variant):
        """0 auto, 1 byte-lattice kernel, 2 bit-sliced kernel (rows of <= 64 sites)."""
        nat.check(self._lib.kmc_set_replica_kernel(self._h, int(variant)))

    def set_sweep_double_buffer(self, on=True):
        """Measurement helper: successive sweeps alternate between two slot-plane buffers."""
        nat.check(self._lib.kmc_set_sweep_double_buffer(self._h, int(bool(on))))

    def set_sweep_kernel(self, variant):
        """Aligned-shape sweep kernel: 0 auto (rolling window for 4096-site rows), 1 nibble/PRMT, 2 byte-parallel."""
        nat.check(self._lib.kmc_set_sweep_kernel(self._h, int(variant)))
