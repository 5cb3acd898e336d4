import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import numpy as np
which = sys.argv[1]
from demo1 import _native
if which != "3":
    _native.LIB_PATH = os.path.join(_native.PKG_ROOT, "libkmc_b200_occ%s.so" % which)
import bench
from demo1.engine import Engine
sw = Engine(bench.SWEEP_DIM, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=1)
sw.upload_state(bench.iid_lattice(bench.SWEEP_DIM, bench.SEED))
for rpc in (0, 2, 3, 4, 5, 6, 7):
    if rpc: os.environ["KMC_SWEEP_ROWS_PER_CTA"] = str(rpc)
    for _ in range(3): sw.sweep()
    ts = []
    for _ in range(30):
        sw.flush_l2(); sw.timer_start(); sw.sweep(); ts.append(sw.timer_stop())
    print("occ %s rows/CTA %s: min %.1f median %.1f avg %.1f us" % (which, rpc or "auto", min(ts) * 1e3, np.median(ts) * 1e3, np.mean(ts) * 1e3))
