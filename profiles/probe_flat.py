"""A/B of a library build (KMC_LIB=...) on the three latency-bound shapes: one trajectory, 2048 replicas, config 5."""
import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import numpy as np
import bench
if os.environ.get("KMC_LIB"):
    import demo1._native as _n
    _n.LIB_PATH = os.path.abspath(os.environ["KMC_LIB"])
from demo1.engine import Engine
def timed(eng, ev, warm):
    eng.run(warm, bench.SEED); eng.sync()
    eng.timer_start(); eng.run(ev, bench.SEED); return eng.timer_stop()
one = Engine(bench.DIM, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=1)
one.upload_state(bench.synthetic_lattices(1, 0))
ms = timed(one, 200000, 2000); print("single trajectory: %.3e events/s" % (200000 / (ms * 1e-3))); one.close()
b = Engine(bench.DIM, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=2048)
b.upload_state(bench.synthetic_lattices(2048, 0))
ms = timed(b, 10000, 1000); print("2048 replicas: %.3e events/s" % (2048 * 10000 / (ms * 1e-3))); b.close()
lat = Engine((1024, 1024), bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=128)
rng = np.random.Generator(np.random.Philox(bench.SEED))
base = rng.choice(np.arange(4, dtype=np.uint8), size=1024 * 1024, p=(0.90, 0.025, 0.025, 0.05)).astype(np.uint8)
lat.upload_state(np.tile(base, (128, 1)))
ms = timed(lat, 4000, 200); print("1024^2 x 128: %.3e events/s" % (128 * 4000 / (ms * 1e-3))); lat.close()
