"""Soak run: long trajectories with all 13 classes live, every launch validated (incremental tables against a fresh
enumeration on every replica), a sample of replicas compared with the oracle at the end."""
import sys, time
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from util_cases import ko, ALL_ORDER, TEST_RATES, exact_state
from demo1.engine import Engine

def soak(dim, R, variant, launches, per_launch, picks):
    eng = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=R)
    eng.set_replica_kernel(variant)
    eng.generate_state("exact", {"divacancy": 0.15, "monovacancy": 0.10}, 99)
    st0 = eng.download_state()[picks].copy()
    t0 = time.time()
    for _ in range(launches):
        eng.run(per_launch, 4242, validate=True)
    final, hist = eng.download_state(), eng.histogram_per_replica()
    assert (eng.steps_done() == launches * per_launch).all()
    for i, r in enumerate(picks):
        o = ko.Oracle(dim, TEST_RATES, ALL_ORDER, st0[i]); o.run_philox(4242, r, 0, launches * per_launch, log=False)
        assert (final[r] == o.status()).all(), (dim, variant, r)
        assert (hist[r] == o.hist()).all(), (dim, variant, r)
    print("soak ok: %s x %d, variant %d, %d events per replica, all classes live, %d launches validated, %.1f s; class mix %s"
          % (dim, R, variant, launches * per_launch, launches, time.time() - t0, (hist.sum(axis=0) / hist.sum()).round(3).tolist()), flush=True)
    eng.close()

soak((64, 64), 4096, 4, 10, 20000, [0, 1, 2047, 4095])
soak((64, 64), 512, 2, 10, 20000, [0, 511])
soak((200, 200), 296, 5, 10, 10000, [0, 295])
soak((200, 200), 16, 1, 5, 10000, [0, 15])
soak((1024, 1024), 32, 5, 4, 10000, [0, 31])
soak((37, 53), 2048, 4, 10, 10000, [0, 2047])
