"""Half-warp replica kernel: replicas per block (env KMC_HW_REPLICAS_PER_BLOCK; 0 = automatic) and batch size."""
import os, sys
sys.path.insert(0, "kmc-dichalcogen_b200"); sys.path.insert(0, ".")
import numpy as np
import bench
from demo1.engine import Engine
E = 10000
for R, hpbs in ((16384, (0, 56, 48, 40, 32, 24, 16)), (8192, (0, 56, 32)), (4096, (0, 56, 16)), (2048, (0, 16))):
    eng = Engine(bench.DIM, bench.wse2_rates(), bench.WSE2_ORDER, n_replicas=R)
    eng.set_replica_kernel(4)
    lat = bench.synthetic_lattices(R, 0)
    for hpb in hpbs:
        if hpb: os.environ["KMC_HW_REPLICAS_PER_BLOCK"] = str(hpb)
        else: os.environ.pop("KMC_HW_REPLICAS_PER_BLOCK", None)
        eng.upload_state(lat)
        eng.run(2000, bench.SEED); eng.sync()
        ts = []
        for _ in range(3):
            eng.upload_state(lat)
            eng.timer_start(); eng.run(E, bench.SEED); ts.append(eng.timer_stop())
        print("R %5d, replicas/block %2d: %.2f ms -> %.3e events/s" % (R, hpb, min(ts), R * E / (min(ts) * 1e-3)), flush=True)
    eng.close()
