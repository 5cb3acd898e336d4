"""Small batches of 64x64 replicas: warp per replica (2) vs half-warp kernel (4) vs byte kernel (1)."""
import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import bench
from demo1.engine import Engine
for R, E in ((1, 200000), (2, 200000), (148, 50000), (296, 50000), (592, 20000)):
    lat = bench.synthetic_lattices(R, 0)
    out = []
    for variant in (2, 4, 1):
        eng = Engine(bench.DIM, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=R)
        eng.set_replica_kernel(variant)
        eng.upload_state(lat); eng.run(1000, bench.SEED); eng.sync()
        ts = []
        for _ in range(3):
            eng.upload_state(lat); eng.timer_start(); eng.run(E, bench.SEED); ts.append(eng.timer_stop())
        eng.close()
        out.append("v%d %.3e (%.3f us/event/replica)" % (variant, R * E / (min(ts) * 1e-3), min(ts) * 1e3 / E))
    print("R %4d: %s" % (R, "   ".join(out)), flush=True)
