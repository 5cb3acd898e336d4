"""Full rate sweep of 16.8 M sites in different shapes (17 planes and all 29): which aligned kernel serves them and how
close each gets to the HBM roofline (18 resp. 30 algorithmic bytes per site, MEASURED_PEAKS hbm_gbs)."""
import os, sys, json
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import bench
from demo1.engine import Engine
from util_cases import WSE2_FULL_ORDER, wse2_full_rates
peak, _ = bench.measured_peaks()
for dim, R in (((4096, 4096), 1), ((1024, 1024), 16), ((256, 256), 256), ((64, 64), 4096), ((16, 64), 16384), ((128, 32), 4096), ((64, 16), 16384)):
    out = []
    for name, rates, order, bps in (("17 planes", bench.wse2_rates(), bench.WSE2_ORDER, 18), ("29 planes", wse2_full_rates(1500.0), WSE2_FULL_ORDER, 30)):
        n = dim[0] * dim[1]
        rng = np.random.Generator(np.random.Philox(7))
        st = rng.choice(np.arange(4, dtype=np.uint8), size=(R, n), p=(0.90, 0.025, 0.025, 0.05)).astype(np.uint8)
        sw = Engine(dim, rates, order, n_replicas=R)
        sw.upload_state(st)
        for _ in range(3): sw.sweep()
        sw.sync(); ms = []
        for _ in range(8):
            sw.flush_l2(); sw.timer_start(); sw.sweep(); ms.append(sw.timer_stop())
        sw.close()
        t = float(np.mean(ms))
        out.append("%s %.1f us (%.2f)" % (name, 1e3 * t, R * n * bps / (t * 1e-3) / 1e9 / peak))
    print("%-12s x %-5d %s" % (dim, R, "   ".join(out)), flush=True)
