"""No-incremental event on small and mid-size lattices: byte path vs bit-plane path (where should the automatic choice switch?)."""
import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import bench
from demo1.engine import Engine
from util_cases import TEST_RATES, ALL_ORDER
for dim in ((64, 64), (128, 128), (200, 200), (256, 256), (512, 512), (1024, 1024)):
    out = []
    for planes in ("0", "1"):
        os.environ["KMC_NOINC_PLANES"] = planes
        eng = Engine(dim, TEST_RATES, ALL_ORDER, n_replicas=1)
        eng.upload_state(np.zeros((1, dim[0] * dim[1]), np.uint8))
        eng.run_noinc(200, 5); eng.sync()
        eng.timer_start(); eng.run_noinc(2000, 5); ms = eng.timer_stop()
        eng.close()
        out.append("planes=%s %.2f us" % (planes, 1e3 * ms / 2000))
    print(dim, "   ".join(out), flush=True)
