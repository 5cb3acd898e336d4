"""Which fused kernel for which batch size: warp per replica (variant 2) vs two replicas per warp (variant 4)."""
import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import bench
from demo1.engine import Engine
E = 5000
for R in (64, 148, 296, 592, 1184, 2368, 4096):
    lat = bench.synthetic_lattices(R, 0)
    out = []
    for variant in (2, 4):
        eng = Engine(bench.DIM, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=R)
        eng.set_replica_kernel(variant)
        eng.upload_state(lat); eng.run(1000, bench.SEED); eng.sync()
        ts = []
        for _ in range(3):
            eng.upload_state(lat); eng.timer_start(); eng.run(E, bench.SEED); ts.append(eng.timer_stop())
        out.append(R * E / (min(ts) * 1e-3))
        eng.close()
    print("R %5d: warp/replica %.3e, half-warp %.3e events/s" % (R, out[0], out[1]), flush=True)
