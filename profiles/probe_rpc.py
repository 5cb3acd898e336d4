"""Sweep kernel tuning probe: rows per CTA of the rolling-window kernel (env KMC_SWEEP_ROWS_PER_CTA)."""
import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import numpy as np
import bench
from demo1.engine import Engine
sw = Engine(bench.SWEEP_DIM, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=1)
sw.upload_state(bench.iid_lattice(bench.SWEEP_DIM, bench.SEED))
for rpc in (1, 2, 3, 4, 5, 6, 8, 10, 16):
    os.environ["KMC_SWEEP_ROWS_PER_CTA"] = str(rpc)
    sw.set_sweep_kernel(0)
    for _ in range(3): sw.sweep()
    ts = []
    for _ in range(20):
        sw.flush_l2(); sw.timer_start(); sw.sweep(); ts.append(sw.timer_stop())
    print("rolling rows/CTA %2d: min %.1f us avg %.1f us -> %.0f GB/s" % (rpc, min(ts) * 1e3, np.mean(ts) * 1e3, 4096 * 4096 * 18 / (np.mean(ts) * 1e-3) / 1e9))
