"""Ceiling probe: the sweep's store pattern alone vs the sweep kernels (4096x4096)."""
import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import numpy as np
import bench
from demo1.engine import Engine
sw = Engine(bench.SWEEP_DIM, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=1)
sw.upload_state(bench.iid_lattice(bench.SWEEP_DIM, bench.SEED))
def timeit(fn, label):
    for _ in range(3): fn()
    ts = []
    for _ in range(30):
        sw.flush_l2(); sw.timer_start(); fn(); ts.append(sw.timer_stop())
    print("%-44s min %.1f us  median %.1f us  avg %.1f us" % (label, min(ts) * 1e3, np.median(ts) * 1e3, np.mean(ts) * 1e3))
timeit(sw.probe_store_pattern, "store pattern only (285 MB, 17 planes)")
timeit(lambda: sw.flush_l2(4096 * 4096 * 17), "cudaMemsetAsync 285 MB")
for variant, rpc in ((1, 0), (0, 2), (0, 5)):
    os.environ["KMC_SWEEP_ROWS_PER_CTA"] = str(rpc)
    sw.set_sweep_kernel(variant)
    timeit(sw.sweep, "sweep variant %d rows/CTA %d" % (variant, rpc))
