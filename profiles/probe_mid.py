"""Mid-size lattices, one trajectory and small batches: byte lattice in shared memory (variant 1) vs wide bit planes
in HBM/L2 (variant 5)."""
import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import bench
from util_cases import TEST_RATES, ALL_ORDER
from demo1.engine import Engine
E = 20000
for dim in ((192, 192), (128, 128), (256, 256), (448, 448)):
    for R in (1, 148, 1184):
        for rates, order, tag in ((bench.wse2_rates(), bench.WSE2_ORDER, "wse2"), (TEST_RATES, ALL_ORDER, "mixed")):
            lat = np.tile(bench.iid_lattice(dim, 3), (R, 1))
            res = []
            for variant in (1, 5):
                try:
                    eng = Engine(dim, rates, order, n_replicas=R)
                    eng.set_replica_kernel(variant)
                    eng.upload_state(lat); eng.run(1000, 5); eng.sync()
                    eng.timer_start(); eng.run(E, 5); ms = eng.timer_stop()
                    res.append("%.2f us/event" % (ms * 1e3 / E))
                    eng.close()
                except Exception as e:
                    res.append("n/a (%s)" % str(e)[:40])
            print("%s R=%4d %-5s: byte %s | wide %s" % (dim, R, tag, res[0], res[1]), flush=True)
