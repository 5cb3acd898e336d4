"""Profiling target: 1024x1024 x 128 replicas, incremental (lattice in HBM/L2)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "kmc-dichalcogen_b200"))
import numpy as np
import bench
if os.environ.get("KMC_LIB"):
    import demo1._native as _n
    _n.LIB_PATH = os.path.abspath(os.environ["KMC_LIB"])
from demo1.engine import Engine
nrep, ev = 128, 1000
lat = Engine((1024, 1024), bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=nrep)
rng = np.random.Generator(np.random.Philox(bench.SEED))
base = rng.choice(np.arange(4, dtype=np.uint8), size=1024 * 1024, p=(0.90, 0.025, 0.025, 0.05)).astype(np.uint8)
lat.upload_state(np.tile(base, (nrep, 1)))
lat.run(100, bench.SEED); lat.sync()
lat.timer_start(); lat.run(ev, bench.SEED); ms = lat.timer_stop()
print("1024^2 x %d: %.3e events/s, %.2f us per event per replica" % (nrep, nrep * ev / (ms * 1e-3), ms * 1e3 / ev))
