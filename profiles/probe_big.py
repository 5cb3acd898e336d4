"""One-off check of 64-bit indexing: BASELINE config 5 at its full size on ONE GPU (1024x1024 x 1024 replicas: 1 GiB
of status bytes, 6.6 GB of planes), first / middle / last replica against the oracle."""
import sys, time
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from util_cases import ko, WSE2_ORDER, wse2_rates
from demo1.engine import Engine
dim, R, seed, n = (1024, 1024), 1024, 20261019, 3000
rates = wse2_rates(1500.0)
eng = Engine(dim, rates, WSE2_ORDER, n_replicas=R)
t0 = time.time(); eng.generate_state("approx", {"divacancy": 0.05, "monovacancy": 0.05}, seed); print("generate %.2f s" % (time.time() - t0))
st0 = eng.download_state()
picks = [0, 511, 1023]
keep = st0[picks].copy(); del st0
eng.run(200, seed, validate=True)
eng.timer_start(); eng.run(n - 200, seed, validate=False); ms = eng.timer_stop()
print("1024^2 x %d: %.3e events/s" % (R, R * (n - 200) / ms * 1e3))
eng.run(0, seed, validate=True)
final = eng.download_state()
for i, r in enumerate(picks):
    o = ko.Oracle(dim, rates, WSE2_ORDER, keep[i]); o.run_philox(seed, r, 0, n, log=False)
    assert (final[r] == o.status()).all(), r
    assert (eng.histogram_per_replica()[r] == o.hist()).all()
print("ok: replicas", picks, "bit-equal to the oracle; zobrist", [hex(int(x)) for x in eng.zobrist(64, 1)[picks]])
