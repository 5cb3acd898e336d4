"""BASELINE config 5 at one GPU's share: 1024x1024 x 128 replicas; one warp per replica vs a CTA per replica."""
import os, sys
import numpy as np
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import bench
from demo1.engine import Engine
rng = np.random.Generator(np.random.Philox(bench.SEED))
base = rng.choice(np.arange(4, dtype=np.uint8), size=1024 * 1024, p=(0.90, 0.025, 0.025, 0.05)).astype(np.uint8)
for nrep in (128, 296, 1024):
    for cta in ("0", "1"):
        os.environ["KMC_WIDE_CTA"] = cta
        lat = Engine((1024, 1024), bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=nrep)
        lat.upload_state(np.tile(base, (nrep, 1)))
        lat.run(200, bench.SEED); lat.sync()
        lat.timer_start(); lat.run(4000, bench.SEED); ms = lat.timer_stop()
        lat.check_status()
        lat.close()
        print("replicas %4d  cta_per_replica=%s: %.3e events/s  (%.2f us per event per replica)" % (
            nrep, cta, nrep * 4000 / (ms * 1e-3), ms * 1e3 / 4000), flush=True)
