"""KmcSim: the reference's engine object (reference demo1/sim.py:31-198) over the CUDA engine.

Same constructor, methods and log dictionaries; the work happens on the GPU through the C ABI
(demo1/engine.py -> libkmc_b200.so).  Events are produced on the device in chunks and handed out
one by one, so `perform_random_move()` keeps its one-event-per-call contract.

What the host keeps: YAML/rule bookkeeping, Arrhenius rates (math.exp exactly as the reference, so
the rate bits are equal), exact `total_rate = math.fsum(weights)` from the logged class counts, and
formatting.  There is no CPU fallback: without the library or a GPU, construction raises.
"""
import math
import os
import warnings

import numpy as np

from . import tables as T

# reference demo1/sim.py:27-29
eV_PER_J = 6.24150913e18
BOLTZMANN__J_PER_K = 1.38064852e-23
BOLTZMANN__eV_PER_K = BOLTZMANN__J_PER_K * eV_PER_J
DEFAULT_KIND = T.DEFAULT_KIND


class RuleSpec:
    """A rule name and its rates or barriers (reference sim.py:170-198)."""

    def __init__(self, rule_name, rates, rate_is_barrier):
        if rule_name not in T.RULE_KINDS:
            raise RuntimeError("config: unknown rule: %s" % rule_name)
        self.name = rule_name
        self._rates = dict(rates)
        self._is_barrier = bool(rate_is_barrier)

    def kinds(self):
        return list(T.RULE_KINDS[self.name])

    def rates(self, temperature=None):
        if self._is_barrier:
            if temperature is None:
                raise ValueError("must specify temperature to compute rates from barrier")
            return {k: math.exp(-v / (temperature * BOLTZMANN__eV_PER_K)) for k, v in self._rates.items()}
        return dict(self._rates)

    def format(self, kind=None):
        return self.name if kind is None else "%s:%s" % (self.name, kind)


def class_table(rule_specs, temperature):
    """Rule specs -> (rates_by_rule, class_rate[13], class_order): the per-class rate vector and the
    weighted_choice input order (rules in the given order x Rule.kinds() order) the engine needs.
    Kind validation as in the reference (sim.py:71-78)."""
    rules_and_rates, rates13, order = {}, [0.0] * T.N_CLASSES, []
    for spec in rule_specs:
        rates = spec.rates(temperature)
        for kind in spec.kinds():
            if kind not in rates:
                raise RuntimeError("config: %s: Missing required rate for kind: %s" % (spec.format(), kind))
        for kind in rates:
            if kind not in spec.kinds():
                warnings.warn("config: %s: Unexpected kind: %s" % (spec.format(), kind))
        rules_and_rates[spec.name] = rates
        for kind in spec.kinds():
            cid = T.CLASS_ID[(spec.name, kind)]
            rates13[cid] = rates[kind]
            order.append(cid)
    return rules_and_rates, rates13, order


def np_first_stop(recs):
    """Index of the first 'nothing can happen' record (class 0xFF) in a chunk, or None."""
    import numpy as np
    hit = np.flatnonzero(recs["cls"] == 0xFF)
    return int(hit[0]) if hit.size else None


class KmcSim:
    """Runs the KMC simulation of one trajectory on the GPU.

    ``seed``/``replica`` select the Philox stream (the reference never seeds; by default a fresh seed
    is drawn from the OS).  ``chunk`` events are generated per kernel launch."""

    def __init__(self, initial_state, rule_specs, temperature=None, incremental=True, seed=None,
                 replica=0, device=0, chunk=4096, validate_launches=False, zobrist_bits=None, zobrist_seed=None):
        from .engine import Engine
        self.rule_specs = list(rule_specs)
        self.temperature = temperature
        self.incremental = incremental
        self.seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed)
        self.replica = int(replica)
        self.chunk = int(chunk)
        self.validate_launches = bool(validate_launches)
        # Zobrist key per logged event (reference state.py:52-54, sim.py:141-142): off unless bits are given
        self.zobrist_bits = None if zobrist_bits is None else int(zobrist_bits)
        self.zobrist_seed = self.seed if zobrist_seed is None else int(zobrist_seed)
        self._zobrist = None
        self._Engine = Engine
        self._device = device
        self._engine = None
        self._buffer = None
        self._cursor = 0
        self._decoded_for = None
        self.state = None
        self.initialize(initial_state)

    # -- setup -----------------------------------------------------------------------------
    def initialize(self, state=None):
        """(Re)build everything; a given state replaces the current one (it is not modified)."""
        if state is not None:
            self.state = state.clone()
        elif self._engine is not None:
            self._sync_state()
        assert self.state is not None
        self.rules_and_rates, rates13, order = class_table(self.rule_specs, self.temperature)
        self._rates13 = rates13
        self._order = order
        if self._engine is None:
            self._engine = self._Engine(self.state.dim, rates13, order, n_replicas=1, device=self._device)
        self.state.validate()
        self._engine.upload_state(self.state.status)
        self._buffer, self._cursor = None, 0
        if self.zobrist_bits is not None:
            from .zobrist import ZobristTracker
            self._zobrist = ZobristTracker(self.state.status, self.state.dim, self.zobrist_bits, self.zobrist_seed)

    def rate(self, rule, kind):
        return self.rules_and_rates[rule][kind]

    def rates_info(self):
        return {"%s:%s" % (rule, kind): rate for rule, rates in self.rules_and_rates.items()
                for kind, rate in rates.items()}

    # -- events ----------------------------------------------------------------------------
    def _refill(self, nsteps):
        from . import _native as nat
        if self.incremental:
            ev = self._engine.run(nsteps, self.seed, replica0=self.replica, log_mode=nat.LOG_FULL,
                                  validate=self.validate_launches)
        else:
            ev = self._engine.run_noinc(nsteps, self.seed, replica0=self.replica, log_mode=nat.LOG_FULL)
        return ev[0]

    def perform_random_move(self):
        """Select and perform one random event; returns the reference's summary dict
        (reference sim.py:106-143)."""
        if self._buffer is None or self._cursor >= len(self._buffer):
            try:
                self._buffer = self._refill(self.chunk)
            except ValueError:
                # total rate hit zero inside the chunk: hand out what happened before it
                self._buffer = self._engine._last_events[0]
                self._cursor = 0
                if self._buffer[0]["cls"] == 0xFF:
                    raise
            else:
                self._cursor = 0
        if self._decoded_for is not self._buffer:
            self._decode_buffer()
        i = self._cursor
        cls = self._dec_cls[i]
        if cls == 0xFF:
            raise ValueError("Cannot choose from total weight of zero!")
        self._cursor += 1
        rule, kind = T.CLASSES[cls]
        d = {
            "rule": rule,
            "move": T.move_info(self._dec_site[i], self._dec_slot[i], self.state.dim),
            "kind": kind,
            "rate": self._rates13[cls],
            "total_rate": self._dec_total[i],           # reference sim.py:138
        }
        if self._zobrist is not None:
            d["zobrist"] = self._zobrist.apply(d)       # reference sim.py:141-142
        return d

    def _decode_buffer(self):
        """Per-launch decoding of the event records for perform_random_move: class / site / slot as Python ints
        and the exactly rounded total rate (math.fsum of count * rate, sim.py:138) of every event, taken for the
        whole chunk at once -- the same values `_format` computes record by record."""
        import numpy as np
        recs = self._buffer
        order = list(self._order)
        rates = np.array([self._rates13[c] for c in order], dtype=np.float64)
        weights = (recs["counts"][:, order].astype(np.float64) * rates[None, :]).tolist()
        self._dec_total = [math.fsum(w) for w in weights]
        self._dec_cls, self._dec_site, self._dec_slot = recs["cls"].tolist(), recs["site"].tolist(), recs["slot"].tolist()
        self._decoded_for = recs

    def _format(self, rec):
        rule, kind = T.CLASSES[int(rec["cls"])]
        weights = [float(int(rec["counts"][c])) * self._rates13[c] for c in self._order]
        d = {
            "rule": rule,
            "move": T.move_info(int(rec["site"]), int(rec["slot"]), self.state.dim),
            "kind": kind,
            "rate": self._rates13[int(rec["cls"])],
            "total_rate": math.fsum(weights),           # reference sim.py:138
        }
        if self._zobrist is not None:
            d["zobrist"] = self._zobrist.apply(d)       # reference sim.py:141-142
        return d

    # -- bulk log writer --------------------------------------------------------------------
    def format_records_json(self, recs):
        """``[json.dumps(self._format(r), sort_keys=True) for r in recs]`` without building a dict per event:
        the products count*rate are taken for the whole chunk at once (same single rounding as the Python
        floats of the slow path), ``math.fsum`` runs per event in C, the move payloads are formatted from
        precomputed coordinates.  Character-identical to the slow path (tests/test_host.py)."""
        import numpy as np
        n = len(recs)
        if n == 0:
            return []
        d0, d1 = self.state.dim
        order = list(self._order)
        rates = np.array([self._rates13[c] for c in order], dtype=np.float64)
        weights = (recs["counts"][:, order].astype(np.float64) * rates[None, :]).tolist()
        totals = [repr(math.fsum(w)) for w in weights]
        site = recs["site"].astype(np.int64)
        a, b = (site // d1), (site % d1)
        cls, slot = recs["cls"].tolist(), recs["slot"].tolist()
        a, b = a.tolist(), b.tolist()
        head = ['{"kind": "%s", "move": ' % T.CLASSES[c][1] for c in range(T.N_CLASSES)]
        import json
        tail = [', "rate": %s, "rule": "%s", "total_rate": ' % (json.dumps(self._rates13[c]), T.CLASSES[c][0])
                for c in range(T.N_CLASSES)]
        out = []
        for i in range(n):
            sl, x, y = slot[i], a[i], b[i]
            if sl in (0, 1, 6):
                mv = '{"node": [%d, %d]}' % (x, y)
            elif sl <= 5:
                mv = '{"layer": %d, "node": [%d, %d]}' % (sl - 1 if sl <= 3 else sl - 3, x, y)
            elif sl <= 12:
                da, db = T.MOVE_TARGET[sl - 7]
                mv = '{"now": [%d, %d], "was": [%d, %d]}' % ((x + da) % d0, (y + db) % d1, x, y)
            elif sl >= T.MM_SLOT0:
                da, db = T.MOVE_TARGET[(sl - T.MM_SLOT0) >> 1]
                mv = '{"layer": %d, "now": [%d, %d], "was": [%d, %d]}' % (
                    ((sl - T.MM_SLOT0) & 1) + 1, (x + da) % d0, (y + db) % d1, x, y)
            else:
                m = T.TREFOIL_MEMBERS[(sl - 13) & 1]
                mv = '{"nodes": [[%d, %d], [%d, %d], [%d, %d]]}' % (
                    x, y, (x + m[1][0]) % d0, (y + m[1][1]) % d1, (x + m[2][0]) % d0, (y + m[2][1]) % d1)
            c = cls[i]
            out.append(head[c] + mv + tail[c] + totals[i] + "}")
        return out

    def iter_json_lines(self, nsteps):
        """The next ``nsteps`` events as JSON strings, chunk by chunk (what the CLI writes; same content and
        order as calling perform_random_move ``nsteps`` times).  Raises ValueError like perform_random_move
        when the total rate reaches zero, after yielding the events before it.  While one chunk is being
        formatted the kernel launch for the next one runs on a helper thread (ctypes drops the GIL), but only
        if the caller still needs events from it, so the lattice never runs ahead of ``nsteps`` by more than
        the usual look-ahead of one chunk."""
        import json
        from concurrent.futures import ThreadPoolExecutor
        done = 0
        pool, pending = None, None

        def fetch():
            try:
                return self._refill(self.chunk), None
            except ValueError:
                return self._engine._last_events[0], None
            except BaseException as e:                     # AssertionError (validate), RuntimeError (CUDA)
                return None, e

        try:
            while done < nsteps:
                if self._zobrist is not None:
                    yield [json.dumps(self.perform_random_move(), sort_keys=True)]
                    done += 1
                    continue
                if self._buffer is None or self._cursor >= len(self._buffer):
                    buf, err = pending.result() if pending is not None else fetch()
                    pending = None
                    if err is not None:
                        raise err
                    self._buffer, self._cursor = buf, 0
                take = min(nsteps - done, len(self._buffer) - self._cursor)
                recs = self._buffer[self._cursor:self._cursor + take]
                stop = np_first_stop(recs)
                if stop is None and self._cursor + take >= len(self._buffer) and done + take < nsteps:
                    if pool is None:
                        pool = ThreadPoolExecutor(max_workers=1)
                    pending = pool.submit(fetch)           # next launch overlaps the formatting below
                if stop is not None:
                    if stop:
                        yield self.format_records_json(recs[:stop])
                    self._cursor += stop
                    raise ValueError("Cannot choose from total weight of zero!")
                self._cursor += take
                done += take
                yield self.format_records_json(recs)
        finally:
            if pending is not None:                        # generator abandoned mid-way: keep the look-ahead
                buf, err = pending.result()
                if err is None:
                    self._buffer, self._cursor = buf, 0
            if pool is not None:
                pool.shutdown(wait=True)

    def perform_random_move_prefetch(self):
        """Fill the look-ahead buffer (one kernel launch of ``chunk`` events)."""
        try:
            self._buffer = self._refill(self.chunk)
        except ValueError:
            self._buffer = self._engine._last_events[0]
        self._cursor = 0

    def is_zobrist_enabled(self):
        return self._zobrist is not None

    def zobrist_key(self):
        """Key of the lattice after the events handed out so far (reference state.py:306-317)."""
        assert self._zobrist is not None, "shouldn't be called unless zobrist bits are set"
        return self._zobrist.key

    def perform_given_move(self, rule, move):
        """Apply an explicit move (reference sim.py:145-157).  ``move`` is either the reference's own
        key for that rule (node / (node, layer, new) / (old, new) / three nodes) or a (site, slot) pair."""
        self._discard_lookahead()
        site, slot = T.slot_of_move(rule, move, self.state.dim)
        if T.SLOT_RULE[slot] != rule:
            raise ValueError("slot %d does not belong to rule %s" % (slot, rule))
        self._engine.perform_given(0, site, slot)
        if self._zobrist is not None:
            self._zobrist.apply({"rule": rule, "move": T.move_info(site, slot, self.state.dim)})

    def _discard_lookahead(self):
        if self._buffer is not None and self._cursor < len(self._buffer):
            raise RuntimeError("events were generated ahead of the caller (chunk=%d): construct KmcSim "
                               "with chunk=1 to interleave explicit moves" % self.chunk)
        self._buffer, self._cursor = None, 0

    def _sync_state(self):
        self.state.status[:] = self._engine.download_state()[0]

    def current_state(self):
        """Lattice after every event generated so far (including look-ahead events in the buffer)."""
        self._sync_state()
        return self.state

    def validate(self):
        """Expensive self-check (reference sim.py:159-168).

        The incrementally maintained move tables only live inside a kernel launch, so their check
        against a from-scratch enumeration runs on the device at the end of every launch when the
        sim was built with ``validate_launches=True`` (what ``--validate-every N`` does, with
        launches of N events) and raises AssertionError from ``perform_random_move``.  Here the
        lattice invariants are checked, and one launch of zero events re-enumerates the lattice."""
        self._sync_state()
        self.state.validate()
        self._engine.run(0, self.seed, replica0=self.replica, validate=True)
        if self._zobrist is not None and (self._buffer is None or self._cursor >= len(self._buffer)):
            # incrementally toggled key against a from-scratch pass over the device lattice
            # (reference state.py:136-141); only meaningful when no look-ahead events are pending
            fresh = int(self._engine.zobrist(self.zobrist_bits, self.zobrist_seed)[0])
            assert fresh == self._zobrist.key, "zobrist key mismatch: %x != %x" % (fresh, self._zobrist.key)
        return True

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None
