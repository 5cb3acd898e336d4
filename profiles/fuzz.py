"""Randomised parity run: random shapes, rule subsets, rates, kernel variants and batch sizes; every case is compared with
the CPU oracle (events, final lattice, sweep tables, no-incremental events).  FUZZ_SECONDS bounds the run."""
import os, sys, time
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from util_cases import ko, iid_state, reference_rules_only
from demo1 import _native as nat
from demo1.engine import Engine

rng = np.random.default_rng(int(os.environ.get("FUZZ_SEED", "1")))
deadline = time.time() + float(os.environ.get("FUZZ_SECONDS", "300"))
ncase = 0
stats = {"variant_runs": 0, "declined": 0, "zero_rate_cases": 0, "events": 0, "mm_cases": 0}
while time.time() < deadline:
    ncase += 1
    d0 = int(rng.choice([5, 7, 8, 9, 13, 16, 31, 32, 33, 64, 65, 100, 128]))
    d1 = int(rng.choice([5, 7, 8, 15, 16, 17, 32, 48, 63, 64, 65, 96, 128, 130, 200, 256] +
                        ([320, 512, 1000, 1024, 2048] if os.environ.get("FUZZ_WIDE") else [])))
    if d0 == 6 or d1 == 6:
        continue
    dim = (d0, d1)
    mm = bool(rng.integers(0, 2))
    ncls = 17 if mm else 13
    live = [c for c in range(ncls) if rng.random() < 0.75]
    if mm and not any(c >= 13 for c in live):
        live.append(13)
    if not live:
        live = [2]
    order = [int(c) for c in rng.permutation(live)]
    rates = [0.0] * ncls
    for c in live:
        rates[c] = float(rng.choice([1.0, 2.0, 5.0, 10.0, 100.0, 0.5])) * (1.0 + (rng.random() if rng.random() < 0.5 else 0.0))
    R = int(rng.choice([1, 2, 3, 5, 8]))
    nsteps = int(rng.choice([20, 60, 150]))
    p = rng.dirichlet([6, 1, 1, 2])
    sts = np.stack([iid_state(dim, int(rng.integers(1, 1 << 30)), p=tuple(p)) for _ in range(R)])
    seed = int(rng.integers(1, 1 << 40))
    tag = "case %d dim %s mm %d order %s R %d steps %d" % (ncase, dim, mm, order, R, nsteps)
    want, wfinal = [], []
    for r in range(R):
        o = ko.Oracle(dim, rates, order, sts[r])
        done, ev = o.run_philox(seed, r, 0, nsteps)
        want.append((done, ev)); wfinal.append(o.status())
    zero = any(d < nsteps for d, _ in want)
    stats["zero_rate_cases"] += int(zero); stats["mm_cases"] += int(mm)
    # incremental kernels
    variants = [0, 1, 3] + ([2, 4] if d1 <= 64 else []) + ([5] if 64 < d1 <= 2048 else [])
    for v in variants:
        eng = Engine(dim, rates, order, n_replicas=R)
        try:
            eng.set_replica_kernel(v)
            eng.upload_state(sts)
            try:
                ev = eng.run(nsteps, seed, log_mode=nat.LOG_COMPACT, validate=True)
            except ValueError:
                assert zero, tag + " variant %d: unexpected zero rate" % v
                ev = eng._last_events
            except RuntimeError as e:
                if any(k in str(e) for k in ("MoveMonovacancy is enabled", "needs", "fit", "variant")):
                    stats["declined"] += 1
                    continue                      # this variant declines the rule set / shape
                raise
            fin = eng.download_state()
            stats["variant_runs"] += 1; stats["events"] += R * nsteps
            for r in range(R):
                d, w = want[r]
                assert (ev[r]["cls"][:d] == w["cls"][:d]).all() and (ev[r]["site"][:d] == w["site"][:d]).all() and \
                       (ev[r]["slot"][:d] == w["slot"][:d]).all(), tag + " variant %d replica %d" % (v, r)
                assert (fin[r] == wfinal[r]).all(), tag + " variant %d replica %d final" % (v, r)
        finally:
            eng.close()
    # sweep + both no-incremental paths
    eng = Engine(dim, rates, order, n_replicas=R)
    eng.upload_state(sts)
    eng.sweep()
    slots, counts, rows = eng.slots(), eng.counts(), eng.row_counts()
    for r in range(R):
        ws, wc, wr = ko.sweep(sts[r], dim)
        if not mm:
            ws, wc, wr = reference_rules_only(ws, wc, wr)
        assert (slots[r] == ws).all() and (counts[r] == wc).all() and (rows[r] == wr).all(), tag + " sweep replica %d" % r
    eng.close()
    for planes in ("0", "1"):
        os.environ["KMC_NOINC_PLANES"] = planes
        eng = Engine(dim, rates, order, n_replicas=R)
        eng.upload_state(sts)
        n2 = min(nsteps, 40)
        try:
            ev = eng.run_noinc(n2, seed, log_mode=nat.LOG_COMPACT)
        except ValueError:
            assert zero or any(d < n2 for d, _ in want), tag
            ev = eng._last_events
        for r in range(R):
            d = min(want[r][0], n2)
            w = want[r][1]
            assert (ev[r]["cls"][:d] == w["cls"][:d]).all() and (ev[r]["site"][:d] == w["site"][:d]).all(), tag + " noinc planes=%s replica %d" % (planes, r)
        eng.close()
    os.environ.pop("KMC_NOINC_PLANES", None)
    if ncase % 10 == 0:
        print("fuzz: %d cases ok" % ncase, flush=True)
print("fuzz done: %d cases, all equal to the oracle; %s" % (ncase, stats))
