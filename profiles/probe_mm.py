"""Where MoveMonovacancy (the full config/WSe2.yaml rule set) runs on general kernels: sweep at 4096^2 and config 5."""
import os, sys, math
import numpy as np
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import bench
if os.environ.get("KMC_LIB"):
    import demo1._native as _n
    _n.LIB_PATH = os.path.abspath(os.environ["KMC_LIB"])
from demo1.engine import Engine
bar = dict(bench.WSE2_BARRIERS); bar.update({13: 2.56175687, 14: 3.03662851, 15: 2.33673358, 16: 2.56175687})
full = [(math.exp(-bar[c] / (bench.TEMPERATURE * bench.KB_EV)) if c in bar else 0.0) for c in range(17)]
full_order = [8, 9, 12, 13, 14, 15, 16, 2, 3, 4, 5, 6, 7]
for name, rates, order in (("13 classes", bench.wse2_rates(), bench.WSE2_ORDER), ("full", full, full_order)):
    sw = Engine(bench.SWEEP_DIM, rates, order, n_replicas=1)
    sw.upload_state(bench.iid_lattice(bench.SWEEP_DIM, bench.SEED))
    for _ in range(3): sw.sweep()
    sw.sync(); ms = []
    for _ in range(10):
        sw.flush_l2(); sw.timer_start(); sw.sweep(); ms.append(sw.timer_stop())
    print("sweep 4096^2 %-10s: %.1f us" % (name, 1e3 * float(np.mean(ms))), flush=True)
    sw.close()
rng = np.random.Generator(np.random.Philox(bench.SEED))
base = rng.choice(np.arange(4, dtype=np.uint8), size=1024 * 1024, p=(0.90, 0.025, 0.025, 0.05)).astype(np.uint8)
for name, rates, order in (("13 classes", bench.wse2_rates(), bench.WSE2_ORDER), ("full", full, full_order)):
    lat = Engine((1024, 1024), rates, order, n_replicas=128)
    lat.upload_state(np.tile(base, (128, 1)))
    lat.run(100, bench.SEED); lat.sync()
    lat.timer_start(); lat.run(1000, bench.SEED); ms = lat.timer_stop()
    lat.check_status(); lat.close()
    print("1024^2 x 128 %-10s: %.3e events/s" % (name, 128 * 1000 / (ms * 1e-3)), flush=True)
# a replica batch of short rows: the 16-sites-per-thread nibble kernel (4096 replicas of 64x64 = 16.8 M sites as well)
for name, rates, order in (("13 classes", bench.wse2_rates(), bench.WSE2_ORDER), ("full", full, full_order)):
    sw = Engine(bench.DIM, rates, order, n_replicas=4096)
    sw.upload_state(bench.synthetic_lattices(4096, 0))
    for _ in range(3): sw.sweep()
    sw.sync(); ms = []
    for _ in range(10):
        sw.flush_l2(); sw.timer_start(); sw.sweep(); ms.append(sw.timer_stop())
    print("sweep 64x64 x 4096 %-10s: %.1f us" % (name, 1e3 * float(np.mean(ms))), flush=True)
    sw.close()
