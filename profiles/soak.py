"""Soak run: long trajectories with all 13 classes live, every launch validated (incremental tables against a fresh
enumeration on every replica), a sample of replicas compared with the oracle at the end."""
import sys, time
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import os
from util_cases import ko, ALL_ORDER, TEST_RATES, exact_state, MM_RATES, MM_ORDER, wse2_full_rates, WSE2_FULL_ORDER
from demo1.engine import Engine

RATES, ORDER = TEST_RATES, ALL_ORDER


def soak(dim, R, variant, launches, per_launch, picks, rates=None, order=None, env=None, noinc=False):
    global TEST_RATES, ALL_ORDER
    TEST_RATES, ALL_ORDER = rates or RATES, order or ORDER
    for k, v in (env or {}).items():
        os.environ[k] = v
    eng = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=R)
    eng.set_replica_kernel(variant)
    eng.generate_state("exact", {"divacancy": 0.15, "monovacancy": 0.10}, 99)
    st0 = eng.download_state()[picks].copy()
    t0 = time.time()
    for _ in range(launches):
        if noinc:
            eng.run_noinc(per_launch, 4242)
        else:
            eng.run(per_launch, 4242, validate=True)
    final, hist = eng.download_state(), eng.histogram_per_replica()
    assert (eng.steps_done() == launches * per_launch).all()
    for i, r in enumerate(picks):
        o = ko.Oracle(dim, TEST_RATES, ALL_ORDER, st0[i]); o.run_philox(4242, r, 0, launches * per_launch, log=False)
        assert (final[r] == o.status()).all(), (dim, variant, r)
        assert (hist[r] == o.hist()).all(), (dim, variant, r)
    print("soak ok: %s x %d, variant %d%s%s, %d events per replica, %d live classes, %d launches validated, %.1f s; class mix %s"
          % (dim, R, variant, " no-incremental" if noinc else "", " " + str(env) if env else "", launches * per_launch,
             int((hist.sum(axis=0) > 0).sum()), launches, time.time() - t0,
             (hist.sum(axis=0) / hist.sum()).round(3).tolist()), flush=True)
    eng.close()

soak((64, 64), 4096, 4, 10, 20000, [0, 1, 2047, 4095])
soak((64, 64), 512, 2, 10, 20000, [0, 511])
soak((200, 200), 296, 5, 10, 10000, [0, 295])
soak((200, 200), 16, 1, 5, 10000, [0, 15])
soak((1024, 1024), 32, 5, 4, 10000, [0, 31])
soak((37, 53), 2048, 4, 10, 10000, [0, 2047])
# round 2: MoveMonovacancy on the half-warp kernel (16 live classes) and the byte kernel (17), the full WSe2 rule set,
# the CTA-per-replica wide kernel, the no-incremental event on planes
MM16 = [c for c in MM_ORDER if c != 1]
soak((64, 64), 4096, 4, 10, 20000, [0, 1, 2047, 4095], rates=MM_RATES, order=MM16)
soak((64, 64), 2048, 0, 10, 20000, [0, 1, 2047], rates=wse2_full_rates(1500.0), order=WSE2_FULL_ORDER)
soak((37, 53), 1024, 4, 10, 10000, [0, 1023], rates=MM_RATES, order=MM16)
soak((64, 64), 296, 1, 5, 10000, [0, 295], rates=MM_RATES, order=MM_ORDER)
soak((200, 200), 148, 5, 10, 10000, [0, 147], env={"KMC_WIDE_CTA": "1"})
soak((1024, 1024), 32, 5, 4, 10000, [0, 31], env={"KMC_WIDE_CTA": "1"})
soak((300, 260), 4, 0, 4, 2000, [0, 3], env={"KMC_NOINC_PLANES": "1"}, noinc=True)
# the one-kernel no-incremental event: cooperative launch per call, per-event launches, several replicas, every rule
soak((512, 512), 3, 0, 3, 3000, [0, 2], env={"KMC_NOINC_PLANES": "1"}, noinc=True)
soak((512, 512), 3, 0, 2, 2000, [0, 2], env={"KMC_NOINC_PLANES": "1", "KMC_NOINC_NO_PERSIST": "1"}, noinc=True)
os.environ.pop("KMC_NOINC_NO_PERSIST", None)
soak((300, 260), 5, 0, 3, 2000, [0, 4], rates=MM_RATES, order=MM_ORDER, env={"KMC_NOINC_PLANES": "1"}, noinc=True)
soak((1024, 1100), 2, 0, 2, 1500, [0, 1], rates=wse2_full_rates(1500.0), order=WSE2_FULL_ORDER, noinc=True)
# MoveMonovacancy on the warp-per-replica bit-plane kernel and on the CTA-per-replica wide kernel
soak((64, 64), 1024, 2, 10, 10000, [0, 1023], rates=MM_RATES, order=MM_ORDER)
soak((200, 200), 64, 5, 10, 10000, [0, 63], rates=MM_RATES, order=MM_ORDER)
