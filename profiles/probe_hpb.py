"""Replica-kernel tuning probe: replicas (half-warps) per block of the half-warp kernel."""
import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import bench
from demo1.engine import Engine
R, E = 16384, 5000
eng = Engine(bench.DIM, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=R)
eng.upload_state(bench.synthetic_lattices(R, 0))
for hpb in (16, 14, 12, 10, 8, 6, 4):
    os.environ["KMC_HW_REPLICAS_PER_BLOCK"] = str(hpb)
    eng.run(E, bench.SEED); eng.sync()
    ts = []
    for _ in range(3):
        eng.timer_start(); eng.run(E, bench.SEED); ts.append(eng.timer_stop())
    print("replicas/block %2d: %.2f ms -> %.3e events/s" % (hpb, min(ts), R * E / (min(ts) * 1e-3)))
