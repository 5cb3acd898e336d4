import sys, os
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import bench
from demo1.engine import Engine
import numpy as np
e = Engine((64, 64), bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=1)
nbytes = 4096 * 4096 * 17
for nb in (nbytes, 256 << 20, 1 << 30):
    e.flush_l2(nb); e.sync()
    ts = []
    for _ in range(10):
        e.timer_start(); e.flush_l2(nb); ts.append(e.timer_stop())
    print("memset %d MB: min %.1f us avg %.1f us -> %.0f GB/s" % (nb >> 20, min(ts) * 1e3, np.mean(ts) * 1e3, nb / (min(ts) * 1e-3) / 1e9))
# same, sweep variants
sw = Engine(bench.SWEEP_DIM, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=1)
sw.upload_state(bench.iid_lattice(bench.SWEEP_DIM, bench.SEED))
for v in (0, 1, 2):
    sw.set_sweep_kernel(v)
    for _ in range(3): sw.sweep()
    ts = []
    for _ in range(20):
        sw.flush_l2(); sw.timer_start(); sw.sweep(); ts.append(sw.timer_stop())
    print("sweep variant %d: min %.1f us avg %.1f us -> %.0f GB/s (18 B/site)" % (v, min(ts) * 1e3, np.mean(ts) * 1e3, 4096 * 4096 * 18 / (np.mean(ts) * 1e-3) / 1e9))
